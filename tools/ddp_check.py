#!/usr/bin/env python
"""Data-parallel gradient check on real GPUs over NCCL (run under torchrun with N >= 2 ranks):
every rank back-propagates the reference loss on its own shard through the native trainer into the flat bucket, ONE all-reduce
averages the buckets, and the result must equal the mean of the per-shard gradients that rank 0 recomputes alone (all shards, one
after the other, same weights).  Also counts the collectives the step issues.  Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import yfv2  # noqa: E402,F401
import synth  # noqa: E402
import model.detector as det  # noqa: E402
import utils.loss as ul  # noqa: E402
import train_ddp  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 8
cfg = synth.coco_cfg(160, 128)
sd = synth.make_state_dict(5)


def shard(r):
    return synth.make_images(100 + r, B, 128, 160).to(dev), synth.make_targets(200 + r, B).to(dev)


def grads_of(r):
    m = det.Detector(80, 3, True)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    bucket = train_ddp.FlatGradBucket(m.parameters())
    bucket.zero()
    x, t = shard(r)
    ul.compute_loss(m(x), t, cfg, dev)[3].backward()
    return bucket


calls = {"n": 0}
orig = dist.all_reduce


def counting(*a, **k):
    calls["n"] += 1
    return orig(*a, **k)


mine = grads_of(rank)
dist.all_reduce = counting
flat = mine.allreduce_mean().clone()
dist.all_reduce = orig
ok, err = True, 0.0
if rank == 0:
    ref = sum(grads_of(r).flat.double() for r in range(world)) / world
    err = float((flat.double() - ref).norm() / ref.norm())
    ok = err < 1e-5                     # fp32 atomics in the weight gradients: run-to-run order noise only
    print(json.dumps({"what": "rank-averaged gradients == mean of per-shard gradients", "world": world, "rel_l2_err": err, "pass": ok,
                      "collectives_in_step": calls["n"], "bucket_bytes": flat.numel() * 4}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
