import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import yfv2, synth, torch, numpy as np
import model.detector as det
from oracle import net as onet
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sd = synth.make_state_dict(21); x = synth.make_images(22, N, 352, 352)
taps={}
with torch.no_grad(): ref = onet.forward(sd, x, taps=taps)
m = det.Detector(80,3,True); m.load_state_dict(sd); m=m.cuda().eval()
xc = x.cuda()
preds = m(xc)
plan = list(m._plans.values())[0]
names = ["stem"] + ["stage2.%d"%i for i in range(4)] + ["stage3.%d"%i for i in range(8)] + ["stage4.%d"%i for i in range(4)]
stages = plan.stage_names
done = 0
for rep in range(2):
  done = 0
  for bi,nm in enumerate(names):
    last = max(i for i, sn in enumerate(stages) if sn.split("/")[0] == nm)
    plan.forward_range(xc, preds, done, last+1); done = last+1
    g = plan.debug_gather(bi).cpu()
    d = (g - taps[nm]).abs()
    bad = (d > 2e-5).nonzero()
    print(rep, nm, "maxerr %.2e" % float(d.max()), "nbad", bad.shape[0], "rows", sorted(set(bad[:,2].tolist()))[:12], "ch", sorted(set(bad[:,1].tolist()))[:12])
