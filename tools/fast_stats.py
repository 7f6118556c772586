"""debugging tool (needs a libaeaj.so built with EXTRA=-DAEAJ_FAST_STATS): how often the table-driven transfer functions fall back"""
import sys, ctypes as C, torch, numpy as np
sys.path.insert(0, 'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0, 'tests')
from aeaj.codec import get_codec
from synth import synth
c = get_codec(0)
rgb = torch.from_numpy(np.stack([synth(1080, 1920, s) for s in range(2)])).cuda()
out = (C.c_ulonglong * 8)()
for sp in ('JzAzBz', 'ICtCp', 'OKLAB'):
    c.lib.aeaj_debug_fast_stats(None, 1)
    enc = c.encode(rgb, sp, (30, 95), (4, 128))
    dec = c.decode_encoded(enc, sp, (30, 95), (4, 128))
    torch.cuda.synchronize()
    c.lib.aeaj_debug_fast_stats(out, 0)
    v = list(out)
    print(sp, 'forward px', v[0], 'fallbacks', v[1], f'({v[1] / max(v[0], 1):.2e})', ' inverse px', v[2], 'XYZ fallbacks', v[3], 'sRGB fallbacks', v[4], f'({v[4] / max(3 * v[2], 1):.2e})')
