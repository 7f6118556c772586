"""measurement tool: per-stage device times (ms) of the C2 workload; usage: python tools/stage_times.py [B] [tensor mask] [space]"""
import sys, torch, numpy as np
sys.path.insert(0,'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0,'tests')
from aeaj.codec import get_codec
from synth import synth
c = get_codec(0)
B = int(sys.argv[1]) if len(sys.argv)>1 else 4
c.tensor_dct = int(sys.argv[2], 0) if len(sys.argv)>2 else 0xf   # mask: bit k = class 16 << k on the tensor cores
H,W=2160,3840
rgb = torch.from_numpy(np.stack([synth(H,W,s) for s in range(B)])).cuda()
sp,q,b=(sys.argv[3] if len(sys.argv)>3 else 'YCbCr'),(30,95),(4,128)
if len(sys.argv)>4: c.lib.aeaj_set_fast_transfer(c.handle, int(sys.argv[4]))   # 1: table-driven transfer functions (opt-in)
args=(B,H,W,sp,b,q)
for _ in range(3):
    enc=c.encode(rgb,sp,q,b); c.decode_encoded(enc,sp,q,b)
c.enable_timing(*args, True)
acc={}
N=10
for _ in range(N):
    enc=c.encode(rgb,sp,q,b)
    for k,v in c.read_timing(*args).items(): acc['E '+k]=acc.get('E '+k,0)+v/N
    c.decode_encoded(enc,sp,q,b)
    for k,v in c.read_timing(*args).items(): acc['D '+k]=acc.get('D '+k,0)+v/N
c.enable_timing(*args, False)
tot=sum(acc.values())
for k,v in acc.items(): print(f'{k:24s} {v:8.4f} ms  {100*v/tot:5.1f}%')
print('sum', tot)
ev=[torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize(); ev[0].record()
for _ in range(20):
    enc=c.encode(rgb,sp,q,b); c.decode_encoded(enc,sp,q,b)
ev[1].record(); torch.cuda.synchronize()
ms=ev[0].elapsed_time(ev[1])/20
print('step ms', ms, 'MP/s', B*H*W/ms/1e3)
