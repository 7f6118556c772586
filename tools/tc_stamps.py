"""debugging tool (needs a libaeaj.so built with EXTRA=-DAEAJ_TC_STAMPS): clock64 timeline of one steady-state tile of CTA 0"""
import sys, torch, numpy as np, ctypes as C
sys.path.insert(0, 'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0, 'tests')
from aeaj.codec import get_codec
from aeaj import native
from synth import synth
c = get_codec(0)
rgb = torch.from_numpy(np.stack([synth(2160, 3840, s) for s in range(8)])).cuda()
sp, q, b = 'YCbCr', (30, 95), (4, 128)
names = {2: 'P wait c0', 3: 'P arrive c0', 4: 'P wait c1', 5: 'P arrive c1', 6: 'P wait c2', 7: 'P arrive c2', 8: 'P wait c3', 9: 'P arrive c3',
         10: 'M full c0', 11: 'M full c1', 12: 'M full c2', 13: 'M full c3', 14: 'M G1 issued', 15: 'M w_ready', 16: 'M G2 issued',
         25: 'P before wait c1', 26: 'P stored c1', 27: 'P loads issued c1 (c3)', 20: 'C start', 21: 'C d1_full', 22: 'C split done', 23: 'C d2_full', 24: 'C epilogue done'}
for mask, which in ((8, 'fwd 128 (last kernel of the encode)'), (2, 'fwd 32')):
    c.tensor_dct = mask
    enc = c.encode(rgb, sp, q, b)
    torch.cuda.synchronize()
    t = (C.c_int * 32)()
    native.check(c.lib.aeaj_tensor_dct_status(c.handle, C.cast(t, C.POINTER(C.c_int))))
    v = {k: t[k] for k in names}
    t0 = min(x for x in v.values() if x)
    print(which)
    for k in sorted(v, key=lambda k: v[k]):
        print(f'   {names[k]:18s} {v[k] - t0:8d} cycles')
