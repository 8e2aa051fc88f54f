"""Data-parallel training step for the mirror Detector: one flat gradient bucket, ONE all-reduce
per step (NCCL over NVLink on GPUs, gloo in the CPU tests), then the reference's optimizer (torch.optim.SGD, momentum 0.949,
weight decay 5e-4, train.py:81-85).  The reference has no distributed code at all; semantics are those of torch DDP: every
rank computes the reference loss on its shard, the all-reduced gradient is the mean over ranks (SURVEY 7 hard part 6)."""
import torch
import torch.distributed as dist


class FlatGradBucket:
    """Views every parameter's .grad into one contiguous fp32 buffer (243 095 floats for 80 classes / 3 anchors)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)      # backward accumulates straight into the bucket
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)      # the single collective of the step
            self.flat.div_(dist.get_world_size(group))
        return self.flat


def make_optimizer(model, lr):
    return torch.optim.SGD(params=model.parameters(), lr=lr, momentum=0.949, weight_decay=0.0005)    # train.py:81-85


def train_step(model, bucket, optimizer, imgs, targets, cfg, compute_loss, group=None):
    """forward (train mode) -> compute_loss -> backward -> one all-reduce -> SGD step.  Returns the 4 loss tensors."""
    bucket.zero()
    preds = model(imgs)
    lbox, lobj, lcls, loss = compute_loss(preds, targets, cfg, imgs.device)
    loss.backward()
    bucket.allreduce_mean(group)
    optimizer.step()
    return lbox, lobj, lcls, loss
