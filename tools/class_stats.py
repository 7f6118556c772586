"""measurement tool: leaf-size census of the C2 bench frames"""
import sys, torch, numpy as np
sys.path.insert(0,'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0,'tests')
from aeaj.codec import get_codec
from synth import synth
c = get_codec(0); B=4; H,W=2160,3840
rgb = torch.from_numpy(np.stack([(synth(H,W,s)*255).astype(np.uint8) for s in range(B)])).cuda()
L = c.download(c.encode(rgb,'YCbCr',(30,95),(4,128)))
tot={}
for k in range(B):
    for l in range(3):
        s=L[k][l]['leaves'][:,2]
        for v,n in zip(*np.unique(s,return_counts=True)): tot[int(v)]=tot.get(int(v),0)+int(n)
al=sum(v*v*n for v,n in tot.items())
for v in sorted(tot): print(v, tot[v], 'leaves', v*v*tot[v]/1e6, 'Msamples', round(100*v*v*tot[v]/al,1), '%')
