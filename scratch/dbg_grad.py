import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np, yfv2, synth
from test_train_gpu import _oracle_grads, make_model
import utils.loss as ul
sd = synth.make_state_dict(61); x = synth.make_images(62, 8, 128, 160); targets = synth.make_targets(63, 8); cfg = synth.coco_cfg(160, 128)
_, _, g32 = _oracle_grads(sd, x, targets, cfg, torch.float32)
_, _, g64 = _oracle_grads(sd, x, targets, cfg, torch.float64)
m = make_model(sd); preds = m(x.cuda()); losses = ul.compute_loss(preds, targets.cuda(), cfg, "cuda"); losses[3].backward()
ours = {n: p.grad.cpu().double() for n, p in m.named_parameters()}
gmax = max(float(v.abs().max()) for v in g64.values())
for k in g64:
    if not (k.startswith("output") or k.startswith("fpn.cls_head_2") or k.startswith("fpn.conv1x1_2") or k.startswith("backbone.stage4.3")): continue
    d = max(float(g64[k].norm()), 1e-5*gmax*g64[k].numel()**0.5)
    print("%-48s ours %.2e  fp32ref %.2e   |g| %.3e" % (k, float((ours[k]-g64[k]).norm())/d, float((g32[k]-g64[k]).norm())/d, float(g64[k].norm())))
