"""measurement tool: pinned-memory PCIe bandwidth of the box, one direction at a time and both at once (what bounds bench.py's e2e)"""
import time, torch
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * n / dt / 1e9
run(True, True, 2)
print("H2D only  GB/s", round(run(True, False), 1))
print("D2H only  GB/s", round(run(False, True), 1))
print("both, per direction GB/s", round(run(True, True), 1))
