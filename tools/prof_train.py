#!/usr/bin/env python
"""Where a training step's time goes: device-side kernel time (torch.profiler / CUPTI) against the wall clock of the step,
per-kernel totals of one step at batch 64 @352x352.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import yfv2  # noqa: E402,F401
import synth  # noqa: E402
import model.detector as det  # noqa: E402
import utils.loss as ul  # noqa: E402
import train_ddp  # noqa: E402

TB = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
torch.manual_seed(2)
m = det.Detector(80, 3, True).to(dev).train()
bucket = train_ddp.FlatGradBucket(m.parameters())
opt = train_ddp.make_optimizer(m, 1e-3)
cfg = synth.coco_cfg()
x = torch.rand(TB, 3, 352, 352, device=dev)
t = synth.make_targets(3, TB).to(dev)
for _ in range(3):
    train_ddp.train_step(m, bucket, opt, x, t, cfg, ul.compute_loss)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    train_ddp.train_step(m, bucket, opt, x, t, cfg, ul.compute_loss)
e1.record(); torch.cuda.synchronize()
wall_ms = e0.elapsed_time(e1) / 5
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    train_ddp.train_step(m, bucket, opt, x, t, cfg, ul.compute_loss)
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        r = rows.setdefault(ev.name[:70], [0, 0.0])
        r[0] += 1; r[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in rows.values())
top = sorted(rows.items(), key=lambda kv: -kv[1][1])[:25]
print(json.dumps({"batch": TB, "step_ms_events": wall_ms, "device_kernel_ms": tot / 1e3, "launches": sum(v[0] for v in rows.values()),
                  "top": [{"kernel": k, "n": v[0], "us": round(v[1], 1)} for k, v in top]}))
