"""debugging tool: which hostile leaves does the device-side guard of aeaj_decode skip?"""
import sys, torch, numpy as np
sys.path.insert(0, 'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0, 'tests')
from aeaj.codec import get_codec
from synth import synth
c = get_codec(0)
H, W, space, q, b = 96, 160, "YCbCr", (40, 80), (4, 64)
rgb = synth(H, W, seed=4)
enc = c.encode(torch.from_numpy(rgb).cuda(), space, q, b)
evil = [[0, 0, 0, 0], [4, 4, 3, 16], [0, 0, 512, 0], [-8, 0, 8, 0], [100000, 0, 8, 0], [0, 0, 8, 2 ** 30], [0, 0, 128, 0]]
print('n leaves', enc.counts[0, :, 0].tolist(), 'plan info cap_coef', [int(x) for x in c.plan_info(1, H, W, space, b, q).cap_coef])
for e in evil:
    lv = enc.leaves[0].clone()
    lv[0, 0] = torch.tensor(e, dtype=torch.int32, device=lv.device)
    c.decode([enc.coef[0], enc.coef[1], enc.coef[2]], [lv, enc.leaves[1], enc.leaves[2]], enc.counts, 1, H, W, space, q, b)
    torch.cuda.synchronize()
    print(e, 'rejected', int(c.last_decode_status[3]))
