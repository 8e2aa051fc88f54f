import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import yfv2, synth, torch, numpy as np
import model.detector as det
from oracle import net as onet
sd = synth.make_state_dict(103); x = synth.make_images(203, 2, 352, 352)
taps={}
with torch.no_grad(): ref = onet.forward(sd, x, taps=taps)
m = det.Detector(80,3,True); m.load_state_dict(sd); m=m.cuda().eval()
xc = x.cuda()
preds = m(xc)
plan = list(m._plans.values())[0]
names = ["stem"] + ["stage2.%d"%i for i in range(4)] + ["stage3.%d"%i for i in range(8)] + ["stage4.%d"%i for i in range(4)]
for i,nm in enumerate(names):
    plan.debug_stop_after(i+1)
    plan.forward(xc)
    g = plan.debug_gather(i).cpu()
    d = (g - taps[nm]).abs()
    print(nm, tuple(g.shape), float(d.max()), float(taps[nm].abs().max()))
    if d.max() > 1e-3:
        bad = (d > 1e-3).nonzero()
        print("  bad count", bad.shape[0], "first", bad[:5].tolist(), "rows with error:", sorted(set(bad[:,2].tolist()))[:50], "cols", sorted(set(bad[:,3].tolist()))[:50], "chans", sorted(set(bad[:,1].tolist()))[:60])
        break
plan.debug_stop_after(0)
plan.forward(xc)
for wid, nm in ((17,"S2"),(18,"S3"),(19,None),(20,None)):
    g = plan.debug_gather(wid).cpu()
    if nm: d=(g-taps[nm]).abs(); print(nm, float(d.max()))
