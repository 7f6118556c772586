"""measurement / debugging tool: every size class 16 .. 128 on the tensor-core kernels, one at a time, against the FP32 kernels
(coefficients, decoded samples, stage times).  usage: python tools/tc_classes.py [H W]"""
import sys, torch, numpy as np
sys.path.insert(0, 'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0, 'tests')
from aeaj.codec import get_codec
from synth import synth
c = get_codec(0)
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2160, 3840)
sp, q, b = 'YCbCr', (30, 95), (4, 128)
rgb = torch.from_numpy(np.stack([synth(H, W, s) for s in range(2)])).cuda()
res = {}
for zig in (False, True):
    for mask in (0, 1, 2, 4, 8, 0xf):
        c.tensor_dct = mask
        enc = c.encode(rgb, sp, q, b, stream=zig)
        L = c.download(enc)
        dec = c.decode_encoded(enc, sp, q, b).clone()
        to = c.tensor_dct_timed_out()
        res[(zig, mask)] = (L, dec)
        if mask:
            L0, d0 = res[(zig, 0)]
            flips = mx = 0
            for k in range(2):
                for l in range(3):
                    d = np.abs(L[k][l]['coef'].astype(np.int64) - L0[k][l]['coef'].astype(np.int64))
                    flips += int((d != 0).sum()); mx = max(mx, int(d.max()))
            print(f'zigzag {int(zig)} mask {mask:#x}: timed_out {to}  coefficient diffs vs FP32 kernels {flips} (max {mx})  decode max|diff| {float((dec - d0).abs().max()):.3g}', flush=True)
sizes = np.concatenate([res[(False, 0)][0][k][l]['leaves'][:, 2] for k in range(2) for l in range(3)])
print('leaves per class', {int(s): int((sizes == s).sum()) for s in np.unique(sizes)})
for mask in (0, 0x8, 0xc, 0xe, 0xf):
    c.tensor_dct = mask
    B = 8
    big = torch.from_numpy(np.stack([synth(2160, 3840, s) for s in range(B)])).cuda() if mask == 0 else big
    args = (B, 2160, 3840, sp, b, q)
    for _ in range(2):
        enc = c.encode(big, sp, q, b); c.decode_encoded(enc, sp, q, b)
    c.enable_timing(*args, True)
    acc = {}
    N = 5
    for _ in range(N):
        enc = c.encode(big, sp, q, b)
        for k, v in c.read_timing(*args).items(): acc['E ' + k] = acc.get('E ' + k, 0) + v / N
        c.decode_encoded(enc, sp, q, b)
        for k, v in c.read_timing(*args).items(): acc['D ' + k] = acc.get('D ' + k, 0) + v / N
    c.enable_timing(*args, False)
    print(f'mask {mask:#x}:', ' '.join(f'{k.split("_")[-1] if "dct" in k else k}={v:.3f}' for k, v in acc.items() if 'dct' in k), ' dct total', round(sum(v for k, v in acc.items() if 'dct' in k), 3), 'all', round(sum(acc.values()), 3))
