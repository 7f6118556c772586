"""measurement tool for ncu: a few encode + decode passes over B 4K frames (default 8, tensor mask 0xf); usage: python tools/prof_step.py [B] [mask]"""
import sys, torch, numpy as np
sys.path.insert(0, 'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0, 'tests')
from aeaj.codec import get_codec
from synth import synth
c = get_codec(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
c.tensor_dct = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0xf
rgb = torch.from_numpy(np.stack([synth(2160, 3840, s) for s in range(B)])).cuda()
sp, q, b = 'YCbCr', (30, 95), (4, 128)
for _ in range(3):
    enc = c.encode(rgb, sp, q, b)
    c.decode_encoded(enc, sp, q, b)
torch.cuda.synchronize()
print('ok', enc.status[:2].tolist())
