"""measurement tool: hysteresis round / dirty-tile statistics of the bench frames"""
import sys, os, torch, numpy as np
sys.path.insert(0,'adaptive-edge-aware-jpeg_b200'); sys.path.insert(0,'tests')
from aeaj.codec import get_codec
from synth import synth
c = get_codec(0)
B=4
rgb = torch.from_numpy(np.stack([synth(2160,3840,s) for s in range(B)])).cuda()
args=(B,2160,3840,'YCbCr',(4,128),(30,95))
for _ in range(3): enc = c.encode(rgb,'YCbCr',(30,95),(4,128))
torch.cuda.synchronize()
enc.status.zero_()
c.enable_timing(*args, True)
t=[]
for _ in range(5):
    enc = c.encode(rgb,'YCbCr',(30,95),(4,128)); t.append(c.read_timing(*args)['hysteresis'])
print('BPS', os.environ.get('AEAJ_HYST_BPS'), 'hyst ms', np.mean(t), 'status', (enc.status.cpu().numpy()[:20]//5).tolist())
