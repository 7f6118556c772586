"""measurement tool: does running two half-batches on two CUDA streams hide the latency-bound stages?"""
import sys, torch, numpy as np
sys.path.insert(0,'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0,'tests')
from aeaj.codec import get_codec
from synth import synth
c = get_codec(0)
B = int(sys.argv[1]); NS = int(sys.argv[2])
H,W=2160,3840
rgb = torch.from_numpy(np.stack([(synth(H,W,s)*255).astype(np.uint8).astype(np.float32)/255.0 for s in range(B)])).cuda()
sp,q,b='YCbCr',(30,95),(4,128)
streams=[torch.cuda.Stream() for _ in range(NS)]
parts=list(rgb.chunk(NS))
def step():
    main=torch.cuda.current_stream()
    ev=torch.cuda.Event(); ev.record(main)
    for i,s in enumerate(streams):
        s.wait_event(ev)
        with torch.cuda.stream(s):
            enc=c.encode(parts[i],sp,q,b,instance=100+i); c.decode(enc.coef,enc.leaves,enc.counts,parts[i].shape[0],H,W,sp,q,b,instance=100+i)
        e2=torch.cuda.Event(); e2.record(s); main.wait_event(e2)
for _ in range(3): step()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/20
print('B',B,'streams',NS,'step ms',ms,'MP/s',B*H*W/ms/1e3)
