"""checker-side report (uses the oracle, hence under tests/): decoded float error of the PQ spaces for both DCT paths; run from the repo root on a GPU box"""
import sys, torch, numpy as np
sys.path.insert(0,'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0,'tests'); sys.path.insert(0,'oracle')
import oracle as O
from aeaj.codec import get_codec
from synth import synth
c=get_codec(0)
for space,shape,q,b in [("ICtCp",(1024,2048),(30,95),(4,128)),("JzAzBz",(1024,1536),(30,95),(4,128)),("ICaCb",(1024,1024),(30,95),(4,128))]:
    H,W=shape
    img=synth(H,W,seed=1)
    for mode in (False,True):
        c.tensor_dct=mode
        enc=c.encode(torch.from_numpy(img[None]).cuda(),space,q,b)
        got=c.download(enc)[0]
        dec=c.decode_encoded(enc,space,q,b)[0].cpu().numpy()
        ref=O.decode_hot([dict(leaves=got[i]["leaves"][:,:3],coef=got[i]["coef"]) for i in range(3)],H,W,space,q,b)
        nan=np.all(dec==1.0,axis=-1)|np.all(ref==1.0,axis=-1)
        d=np.abs(dec-ref); d[nan]=0
        lsb=np.abs((dec*255).astype(np.uint8).astype(int)-(ref*255).astype(np.uint8).astype(int)); lsb[nan]=0
        idx=np.unravel_index(d.argmax(),d.shape)
        print(space,'tensor' if mode else 'fp32','max diff',d.max(),'at value',ref[idx],'lsb max',lsb.max(),'n>1e-5',int((d>1e-5).sum()),'nan frac',nan.mean())
