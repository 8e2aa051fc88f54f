"""world_size-2 gloo test (CPU) of the N>1 training plumbing: the flat gradient bucket and its single all-reduce."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import yfv2  # noqa: F401
    import train_ddp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7))]
    bucket = train_ddp.FlatGradBucket(params)
    assert bucket.flat.numel() == 22 and params[1].grad.data_ptr() == bucket.flat[15:].data_ptr()
    loss = (params[0] * (rank + 1)).sum() + (params[1] ** 2).sum() * (rank + 1)
    loss.backward()                                   # accumulates straight into the bucket views
    flat = bucket.allreduce_mean().clone()
    expect0 = torch.full((3, 5), (1 + 2) / 2.0)
    expect1 = 2 * params[1].detach() * (1 + 2) / 2.0
    ok = torch.allclose(flat[:15].view(3, 5), expect0) and torch.allclose(flat[15:], expect1)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29517, out), nprocs=2, join=True)
    assert out[0] and out[1]
