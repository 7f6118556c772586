#!/usr/bin/env python
"""bench.py -- encode+decode megapixels/s of the adaptive edge-aware JPEG hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Workload (BASELINE.json configs[1], "C2"): synthetic 3840x2160 RGB frames (tests/synth.py, seeds
rank*B+i), YCbCr, quality (30,95), blocks (4,128).  One step = one pass of the hot path -- encode
(colour -> quantised coefficients + quadtree) and decode (coefficients -> RGB) -- over a batch of B
distinct frames; B*99.5 MB of input per step is larger than the 126 MB L2 (no flush needed).
Entropy coding / .ajpg framing are host-side and outside the timed region (north_star).

  value   : whole-job MP/s with inputs resident in HBM, CUDA events, max over ranks
  e2e     : the same through the host-buffer API, every host<->device copy inside the timed region, in the reference's
            own end-to-end flow: 8-bit pixels in (Image.load), coefficient / leaf / state streams to the host and back
            (where the host-side entropy coder sits), 8-bit pixels out (Image.save / get_uint8).  `e2e.f32` is the same
            with float32 host buffers on both ends (Image.data in, Image.data out: 4x the pixel bytes over PCIe)
  roofline: the dominant kernel, per-launch algorithmic bytes / its CUDA-event duration
  cpu_baseline: the CPU oracle port on the same frames, on this box's host cores

N>1: launched by torchrun, one rank per GPU; frames are sharded across ranks, no data-path
collective (the path shards by image); timing = max over ranks (all_reduce MAX).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "adaptive-edge-aware-jpeg_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

H, W = 2160, 3840
SPACE, QRANGE, BRANGE = "YCbCr", (30, 95), (4, 128)
WORKLOAD = "C2: synthetic 3840x2160 RGB (8-bit pixels / 255 as float32, what Image.load yields), YCbCr, quality 30-95, blocks 4-128"
ALG_BYTES_PER_PX = 36.0          # 12 B RGB in + 6 B coefficients out, both directions


def _oracle():
    """CPU oracle (test infrastructure) -- only the cpu_baseline / reference legs may touch it."""
    p = os.path.join(ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)
    import oracle as O
    return O


def shard_seeds(rank: int, frames_per_rank: int):
    """frames are sharded by rank: rank r owns seeds r*B .. r*B+B-1 (weak scaling, no data-path collective)"""
    return list(range(rank * frames_per_rank, (rank + 1) * frames_per_rank))


def reduce_max(t, dist):
    """the only cross-rank exchange of the benchmark: max over ranks of the timings"""
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def whole_job_mps(world: int, frames_per_rank: int, steps: int, ms: float) -> float:
    return world * frames_per_rank * (H * W / 1e6) * steps / (ms / 1e3)


def _frames_u8(seeds):
    """8-bit pixels of the synthetic frames (what an image file holds)"""
    from synth import synth
    return np.stack([(synth(H, W, seed=s) * 255).astype(np.uint8) for s in seeds])


def _frames(seeds):
    """the float32 frames the reference works on: Image.load's imread(path).astype(float32) / 255.0 (image.py:84)"""
    return _frames_u8(seeds).astype(np.float32) / 255.0


def cpu_baseline(frames: np.ndarray, threads: int):
    """encode+decode of the oracle port on `frames` with `threads` OpenMP threads -> (MP/s, seconds)."""
    O = _oracle()
    O.set_threads(threads)
    O.encode_hot(frames[0][:256, :256].copy(), SPACE, QRANGE, BRANGE)          # warm-up (tables, page-in)
    t0 = time.perf_counter()
    for f in frames:
        enc = O.encode_hot(f, SPACE, QRANGE, BRANGE)
        O.decode_hot(enc, H, W, SPACE, QRANGE, BRANGE)
    dt = time.perf_counter() - t0
    return len(frames) * H * W / 1e6 / dt, dt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Python over
    un-vendored wheels and /root/reference does not exist on the GPU box, so this arm times the
    validated CPU port (oracle/, pinned against the reference) with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O = _oracle()
    cores = os.cpu_count() or 1
    O.set_threads(cores)
    frames = _frames([0])
    for _ in range(max(args.warmup, 1)):
        O.decode_hot(O.encode_hot(frames[0], SPACE, QRANGE, BRANGE), H, W, SPACE, QRANGE, BRANGE)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.decode_hot(O.encode_hot(frames[0], SPACE, QRANGE, BRANGE), H, W, SPACE, QRANGE, BRANGE)
    dt = time.perf_counter() - t0
    v = args.steps * H * W / 1e6 / dt
    line = {"impl": "reference", "metric": "encode+decode megapixels/sec", "value": v, "unit": "MP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": 1},
            "cpu_baseline": {"value": v, "unit": "MP/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} x 1 frame 3840x2160 (seed 0), encode+decode hot path, OpenMP {cores} threads"},
            "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16, help="frames per step per GPU")
    ap.add_argument("--streams", type=int, default=2, help="CUDA streams the batch of a step is spread over (sub-batches run concurrently)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=2, help="frames in the bounded cpu_baseline sample (0 = skip)")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself the way the driver does (one rank per GPU under torchrun)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from aeaj.codec import get_codec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")                  # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    codec = get_codec(local)
    frames_u8 = _frames_u8(shard_seeds(rank, B))
    frames_np = frames_u8.astype(np.float32) / 255.0
    host_in = torch.from_numpy(frames_np).pin_memory()
    host_in8 = torch.from_numpy(frames_u8).pin_memory()
    rgb = host_in.to(f"cuda:{local}")
    mp_per_step = B * H * W / 1e6

    def step():
        # encode + decode of the batch; two half-batches on two CUDA streams (fork / join around the current stream)
        return codec.roundtrip_device(rgb, SPACE, QRANGE, BRANGE, streams=args.streams)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    out = torch.cat(out)
    launches_per_step = codec.last_launches
    # the same step with 8-bit pixels resident in HBM on both ends (18 instead of 36 compulsory B/px)
    rgb8 = host_in8.to(f"cuda:{local}")
    for _ in range(2):
        codec.roundtrip_device(rgb8, SPACE, QRANGE, BRANGE, streams=args.streams, out="u8")
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for _ in range(args.steps):
        out8 = codec.roundtrip_device(rgb8, SPACE, QRANGE, BRANGE, streams=args.streams, out="u8")
    u1.record()
    barrier()
    ms_u8 = u0.elapsed_time(u1)
    u8_check = int((torch.cat(out8).to(torch.int16) - (out * 255).to(torch.uint8).to(torch.int16)).abs().max())
    del rgb8
    torch.cuda.synchronize()
    status = codec._plan(rgb.chunk(args.streams)[0].shape[0], H, W, SPACE, BRANGE, QRANGE, instance=1000).out.status.cpu().numpy()

    # ---- e2e: host buffers in, host buffers out ------------------------------------------------
    # pinned host RGB -> H2D -> encode -> D2H (coefficients, leaves, states) -> H2D -> decode -> D2H pinned host RGB,
    # one frame per job, jobs pipelined over 8 streams so that copies in both directions overlap (PCIe-bound)
    out_ref = out.cpu()
    host_out8 = torch.empty_like(host_in8).pin_memory()
    for _ in range(2):
        codec.roundtrip_host_pipelined(host_in8, host_out8, SPACE, QRANGE, BRANGE, repeat=3)   # touches every slot's plan + staging
    barrier()
    # same pixels as the device-resident float path: (data * 255).astype(uint8) of its output
    e2e_check = int((host_out8.to(torch.int16) - (out_ref * 255).to(torch.uint8).to(torch.int16)).abs().max())
    t0 = time.perf_counter()
    h2d, d2h = codec.roundtrip_host_pipelined(host_in8, host_out8, SPACE, QRANGE, BRANGE, repeat=args.steps)
    h2d, d2h = h2d // args.steps, d2h // args.steps
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    # the float32-buffer variant (Image.data in / out)
    host_out = torch.empty_like(host_in).pin_memory()
    codec.roundtrip_host_pipelined(host_in, host_out, SPACE, QRANGE, BRANGE, repeat=2)
    barrier()
    f32_check = float((host_out - out_ref).abs().max())
    f32_steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    h2d_f, d2h_f = codec.roundtrip_host_pipelined(host_in, host_out, SPACE, QRANGE, BRANGE, repeat=f32_steps)
    barrier()
    e2e_f32_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None

    # ---- per-stage timing (separate instrumented single-stream passes over one sub-batch; CUDA events on the launching stream)
    Bs = rgb.chunk(args.streams)[0].shape[0]                        # frames per launch (sub-batch)
    sub = rgb[:Bs]
    codec.enable_timing(Bs, H, W, SPACE, BRANGE, QRANGE, True)
    stage_ms = {}
    reps = max(3, min(args.steps, 10))
    for _ in range(reps):
        enc = codec.encode(sub, SPACE, QRANGE, BRANGE)
        for k, v in codec.read_timing(Bs, H, W, SPACE, BRANGE, QRANGE).items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v / reps
        codec.decode_encoded(enc, SPACE, QRANGE, BRANGE)
        for k, v in codec.read_timing(Bs, H, W, SPACE, BRANGE, QRANGE).items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v / reps
    codec.enable_timing(Bs, H, W, SPACE, BRANGE, QRANGE, False)
    counts_np = codec._plan(Bs, H, W, SPACE, BRANGE, QRANGE).out.counts.cpu().numpy()

    if world > 1:
        t = reduce_max(torch.tensor([ms, e2e_ms, e2e_f32_ms, ms_u8], device=f"cuda:{local}", dtype=torch.float64), dist)
        ms, e2e_ms, e2e_f32_ms, ms_u8 = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = whole_job_mps(world, B, args.steps, ms)
    e2e_value = whole_job_mps(world, B, args.steps, e2e_ms)
    # roofline: dominant stage of the step
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = max(stage_ms, key=stage_ms.get)
    n_samples = float(counts_np[:, :, 2].sum())                       # coefficients = samples incl. padding, whole batch
    full_px = Bs * H * W                                              # pixels one launch processes
    alg = {  # algorithmic bytes per launch of each stage (DESIGN.md "kernels")
        "color_forward_planar": full_px * (12 + 6 + 1.5), "upsample_color_inverse": full_px * (6 + 12),
        "prefilter": full_px * 1.5 * 2, "clahe_hist": full_px * 1.5, "canny_nms": full_px * 1.5 * (1 + 0.25),
    }
    if dom.startswith("dct_quant_") or dom.startswith("dequant_idct_"):
        s = int(dom.rsplit("_", 1)[1])
        lg = int(np.log2(s))
        # samples of this size class: recount from the leaf lists
        plan = codec._plan(Bs, H, W, SPACE, BRANGE, QRANGE)
        ns = 0
        for l in range(3):
            lv = plan.out.leaves[l].cpu().numpy()
            for b in range(Bs):
                sz = lv[b, : counts_np[b, l, 0], 2]
                ns += int((sz == s).sum()) * s * s
        alg[dom] = ns * 8.0                                            # 4 B in + 4 B out per sample
    alg_bytes = alg.get(dom)
    dom_ms = stage_ms[dom]
    # measured DRAM traffic of that kernel per launch (dram__bytes_read + write from the committed `ncu --set full`
    # capture of one 8-frame launch, tools/prof_step.py); null for other sub-batch sizes / kernels not captured
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        kname = {"prefilter": "k_prefilter", "canny_nms": "k_canny_nms", "color_forward_planar": "k_color_forward_planar<0, 0, 0>",
                 "dct_quant_128": "k_dct_tc128<0>", "dequant_idct_128": "k_dct_tc128<1>", "upsample_color_inverse": "k_upsample2x_color_inverse<0>", "dct_quant_64": "k_dct_cta<64, 0>", "hysteresis": "k_hysteresis_rounds"}.get(dom)
        if Bs == 8 and kname in tr:
            traffic = tr[kname]
    except Exception:
        pass
    ncu_stats = None                                                   # what actually limits that kernel (same ncu capture)
    try:
        ncu_stats = json.load(open(os.path.join(ROOT, "profiles", "r1_kernel_stats.json"))).get(kname)
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": (alg_bytes / 1e9) / (dom_ms / 1e3) if alg_bytes else None, "peak": peak,
            "unit": "GB/s", "frac": ((alg_bytes / 1e9) / (dom_ms / 1e3) / peak) if alg_bytes else None, "traffic": traffic,
            "algorithmic_bytes": alg_bytes, "ncu": ncu_stats,
            "peak_source": peak_src, "kernel_ms": dom_ms, "kernel_share_of_step": dom_ms / sum(stage_ms.values()),
            "frames_per_launch": Bs,
            "pipeline_achieved": (ALG_BYTES_PER_PX * B * H * W * world / 1e9) / (ms / args.steps / 1e3),
            "pipeline_frac": (ALG_BYTES_PER_PX * B * H * W / 1e9) / (ms / args.steps / 1e3) / peak,
            "stage_ms": {k: round(v, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])}}

    cpu = None
    if args.cpu_frames > 0:
        cores = os.cpu_count() or 1
        v, dt = cpu_baseline(frames_np[: min(args.cpu_frames, B)], cores)
        cpu = {"value": v, "unit": "MP/s", "cores": cores, "kind": "port",
               "sample": f"{min(args.cpu_frames, B)} of this step's frames, encode+decode hot path, oracle port (C, OpenMP {cores} threads), {dt:.1f} s"}

    line = {"metric": "encode+decode megapixels/sec", "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "l2": "inputs larger than L2 (no flush)",
                       "streams_per_step": args.streams, "parallelism": f"batch-sharded x{world}", "hysteresis_rounds": int(status[0]), "hysteresis_converged": int(status[1])},
            "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
                    "api": "DeviceCodec.roundtrip_host_pipelined (pinned host uint8 pixels in/out, int32 streams to the host and back, 8 CUDA streams)",
                    "max_abs_diff_vs_device_path": e2e_check,
                    "f32": {"value": whole_job_mps(world, B, f32_steps, e2e_f32_ms), "unit": "MP/s", "steps": f32_steps,
                            "h2d_bytes_per_step": int(h2d_f // f32_steps), "d2h_bytes_per_step": int(d2h_f // f32_steps),
                            "api": "same call with float32 host buffers (Image.data in / out)", "max_abs_diff_vs_device_path": f32_check}},
            "value_u8_io": {"value": whole_job_mps(world, B, args.steps, ms_u8), "unit": "MP/s", "ms_per_step": ms_u8 / args.steps,
                            "note": "same step, uint8 pixels resident in HBM in and out (Image.load / Image.get_uint8 conversions on the device)",
                            "max_abs_diff_vs_f32_path_u8_view": u8_check},
            "gpu_launches": int(launches_per_step * args.steps), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
