"""measurement tool: tensor-core vs FP32 128x128 DCT / IDCT (mismatch counts, phase clocks, stage times); usage: python tools/tc_check.py H W"""
import sys, torch, numpy as np
sys.path.insert(0,'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0,'tests')
from aeaj.codec import get_codec
from synth import synth
c = get_codec(0)
H,W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv)>2 else (512,768)
B = 4
rgb = torch.from_numpy(np.stack([synth(H,W,s) for s in range(B)])).cuda()
sp,q,b = 'YCbCr',(30,95),(4,128)
c.tensor_dct=False
ref = c.download(c.encode(rgb,sp,q,b))
c.tensor_dct=True
enc = c.encode(rgb,sp,q,b); torch.cuda.synchronize()
print('timed_out', c.tensor_dct_timed_out(), 'phase cycles (chunk0,chunk1,chunk2,chunk3,gemm1 wait,split,issue2,wait2,tmem->smem,quantize+store):', c.tensor_dct_phase_cycles)
got = c.download(enc)
tot=0; bad=0; mx=0; n128=0
for k in range(B):
    for l in range(3):
        a=ref[k][l]['coef'].astype(np.int64); g=got[k][l]['coef'].astype(np.int64)
        assert np.array_equal(ref[k][l]['leaves'], got[k][l]['leaves'])
        lv=ref[k][l]['leaves']; n128+=int((lv[:,2]==128).sum())
        d=np.abs(a-g); tot+=a.size; bad+=int((d!=0).sum()); mx=max(mx,int(d.max()))
print('leaves128', n128, 'coef total', tot, 'mismatch', bad, 'max', mx)
args=(B,H,W,sp,b,q)
for mode in (False, True):
    c.tensor_dct=mode
    for _ in range(3): c.encode(rgb,sp,q,b)
    c.enable_timing(*args, True); t=[]
    for _ in range(5):
        c.encode(rgb,sp,q,b); t.append(c.read_timing(*args)['dct_quant_128'])
    c.enable_timing(*args, False)
    print('tensor' if mode else 'fp32', 'dct_quant_128 ms', np.mean(t))

# zigzag stream layout through the tensor path
c.tensor_dct=False
r=c.download(c.encode(rgb,sp,q,b,stream=True))
c.tensor_dct=True
g=c.download(c.encode(rgb,sp,q,b,stream=True))
bad=0
for k in range(B):
    for l in range(3):
        bad+=int((r[k][l]['coef']!=g[k][l]['coef']).sum())
print('zigzag-layout mismatches', bad)
# inverse: tensor vs fp32 decode
c.tensor_dct=False
enc = c.encode(rgb,sp,q,b)
d0 = c.decode_encoded(enc,sp,q,b).clone()
c.tensor_dct=True
d1 = c.decode_encoded(enc,sp,q,b).clone()
print('inverse timed_out', c.tensor_dct_timed_out(), 'cycles', c.tensor_dct_phase_cycles)
diff=(d0-d1).abs()
print('decode max abs diff', float(diff.max()), 'u8 lsb diff max', int(((d0*255).to(torch.uint8).int()-(d1*255).to(torch.uint8).int()).abs().max()), 'n lsb', int(((d0*255).to(torch.uint8)!=(d1*255).to(torch.uint8)).sum()))
for mode in (False, True):
    c.tensor_dct=mode
    for _ in range(3): c.decode_encoded(enc,sp,q,b)
    c.enable_timing(*args, True); t=[]
    for _ in range(5):
        c.decode_encoded(enc,sp,q,b); t.append(c.read_timing(*args)['dequant_idct_128'])
    c.enable_timing(*args, False)
    print('tensor' if mode else 'fp32', 'dequant_idct_128 ms', np.mean(t))
