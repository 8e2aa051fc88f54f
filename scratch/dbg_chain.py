import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, torch.nn.functional as F, yfv2
from model import train_ops as T
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(0)
def P(*s, scale=1.0): return (torch.randn(*s, device="cuda", generator=g)*scale).requires_grad_(True)
x0 = P(8,72,8,10)
wd1 = P(72,1,5,5, scale=0.2); g1 = P(72); b1 = P(72, scale=0.2)
wp1 = P(72,72,1,1, scale=0.12); g2 = P(72); b2 = P(72, scale=0.2)
wd2 = P(72,1,5,5, scale=0.2); g3 = P(72); b3 = P(72, scale=0.2)
wp2 = P(72,72,1,1, scale=0.12); g4 = P(72); b4 = P(72, scale=0.2)
params = [x0, wd1,g1,b1,wp1,g2,b2,wd2,g3,b3,wp2,g4,b4]
names = "x0 wd1 g1 b1 wp1 g2 b2 wd2 g3 b3 wp2 g4 b4".split()
dy = torch.randn(8,72,8,10, device="cuda", generator=g)
def ours():
    z = lambda: (torch.zeros(72, device="cuda"), torch.ones(72, device="cuda"))
    h = T.BnTrain.apply(T.DwConv.apply(x0, wd1, 1), g1, b1, *z(), True)
    h = T.BnTrain.apply(T.Conv1x1.apply(h, wp1, None), g2, b2, *z(), False)
    h = T.BnTrain.apply(T.DwConv.apply(h, wd2, 1), g3, b3, *z(), True)
    return T.BnTrain.apply(T.Conv1x1.apply(h, wp2, None), g4, b4, *z(), False)
def ref(dt):
    ps = [p.detach().to(dt).requires_grad_(True) for p in params]
    x0_, wd1_,g1_,b1_,wp1_,g2_,b2_,wd2_,g3_,b3_,wp2_,g4_,b4_ = ps
    bn = lambda t, gg, bb: F.batch_norm(t, None, None, gg, bb, True, 0.1, 1e-5)
    h = F.relu(bn(F.conv2d(x0_, wd1_, None, 1, 2, 1, 72), g1_, b1_))
    h = bn(F.conv2d(h, wp1_), g2_, b2_)
    h = F.relu(bn(F.conv2d(h, wd2_, None, 1, 2, 1, 72), g3_, b3_))
    h = bn(F.conv2d(h, wp2_), g4_, b4_)
    h.backward(dy.to(dt))
    return h.detach(), [p.grad for p in ps]
y = ours(); y.backward(dy); go = [p.grad.clone() for p in params]
y32, g32 = ref(torch.float32); y64, g64 = ref(torch.float64)
print("fwd err ours %.2e ref32 %.2e" % (float((y.double()-y64).norm()/y64.norm()), float((y32.double()-y64).norm()/y64.norm())))
for n, a, b, c in zip(names, go, g32, g64):
    print("%-4s ours %.2e  ref32 %.2e" % (n, float((a.double()-c).norm()/c.norm()), float((b.double()-c).norm()/c.norm())))
