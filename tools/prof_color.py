"""measurement tool for ncu: one encode + decode of 4 4K frames in a non-linear colour space; usage: python tools/prof_color.py [space]"""
import sys, torch, numpy as np
sys.path.insert(0, 'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0, 'tests')
from aeaj.codec import get_codec
from synth import synth
c = get_codec(0)
sp = sys.argv[1] if len(sys.argv) > 1 else 'JzAzBz'
rgb = torch.from_numpy(np.stack([synth(2160, 3840, s) for s in range(4)])).cuda()
for _ in range(2):
    enc = c.encode(rgb, sp, (30, 95), (4, 128))
    c.decode_encoded(enc, sp, (30, 95), (4, 128))
torch.cuda.synchronize()
print('ok')
