"""measurement tool for ncu: two device-resident steps (encode+decode) of the C2 workload; usage: python tools/prof_step.py [B]"""
import sys, torch, numpy as np
sys.path.insert(0,'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0,'tests')
from aeaj.codec import get_codec
from synth import synth
c = get_codec(0)
B = int(sys.argv[1]) if len(sys.argv)>1 else 8
H,W=2160,3840
rgb = torch.from_numpy(np.stack([(synth(H,W,s)*255).astype(np.uint8).astype(np.float32)/255.0 for s in range(B)])).cuda()
sp,q,b='YCbCr',(30,95),(4,128)
for _ in range(2):
    enc=c.encode(rgb,sp,q,b); c.decode_encoded(enc,sp,q,b)
torch.cuda.synchronize()
print('ok')
