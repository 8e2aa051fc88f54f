"""Times one training step (train-mode forward -> device loss -> backward -> flat-bucket all-reduce -> SGD) of the mirror
Detector on BASELINE config[2]'s per-GPU shard (batch 64 @352x352, ~7 boxes per image).  Informational: the training
operators of round 1 are correctness-first FFMA kernels."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import yfv2  # noqa: E402,F401
import synth  # noqa: E402
import model.detector as det  # noqa: E402
import utils.loss as ul  # noqa: E402
import train_ddp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(2)
m = det.Detector(80, 3, True).cuda().train()
bucket = train_ddp.FlatGradBucket(m.parameters())
opt = train_ddp.make_optimizer(m, 1e-3)
x = torch.rand(N, 3, 352, 352).cuda()
targets = synth.make_targets(2, N).cuda()
cfg = synth.coco_cfg()
for _ in range(2):
    train_ddp.train_step(m, bucket, opt, x, targets, cfg, ul.compute_loss)
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 5
for _ in range(K):
    losses = train_ddp.train_step(m, bucket, opt, x, targets, cfg, ul.compute_loss)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
print(json.dumps({"what": "train step (fwd+loss+bwd+SGD), 1 GPU", "batch": N, "ms_per_step": 1e3 * dt, "images_per_s": N / dt,
                  "loss": float(losses[3])}))
