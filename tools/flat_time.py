"""measurement tool: prefilter time on flat / noisy inputs (histogram atomics worst cases)"""
import sys, torch, numpy as np
sys.path.insert(0,'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0,'tests')
from aeaj.codec import get_codec
c=get_codec(0); B,H,W=8,2160,3840
sp,q,b='YCbCr',(30,95),(4,128); args=(B,H,W,sp,b,q)
rng=np.random.default_rng(0)
for name,img in [('flat',np.full((B,H,W,3),0.5,np.float32)),('noise',rng.random((B,H,W,3),dtype=np.float32))]:
    rgb=torch.from_numpy(img).cuda()
    for _ in range(2): c.encode(rgb,sp,q,b)
    c.enable_timing(*args,True); t=[]
    for _ in range(5):
        c.encode(rgb,sp,q,b); t.append(c.read_timing(*args)['prefilter'])
    c.enable_timing(*args,False)
    print(name,'prefilter ms',np.mean(t))
