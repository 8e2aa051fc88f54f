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
        self.views = []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            p.grad = v                                              # backward accumulates straight into the bucket
            self.views.append(v)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def reattach(self):
        """`optimizer.zero_grad()` defaults to set_to_none=True since torch 2.0, which drops the views into the bucket (the
        next backward then allocates fresh .grad tensors and the all-reduce would see zeros).  Re-point every .grad at its
        slice, copying a gradient that was accumulated elsewhere."""
        for p, view in zip(self.params, self.views):
            g = p.grad
            if g is view:                                           # the common case: nothing touched the views
                continue
            if g is None:
                p.grad = view
            elif g.data_ptr() != view.data_ptr():
                view.copy_(g)
                p.grad = view
            else:
                p.grad = view

    def allreduce_mean(self, group=None):
        self.reattach()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)      # the single collective of the step
            self.flat.div_(dist.get_world_size(group))
        return self.flat


def make_optimizer(model, lr):
    return torch.optim.SGD(params=model.parameters(), lr=lr, momentum=0.949, weight_decay=0.0005)    # train.py:81-85


def train_step(model, bucket, optimizer, imgs, targets, cfg, compute_loss, group=None, accumulate=False, step=True):
    """forward (train mode) -> compute_loss -> backward -> one all-reduce -> SGD step.  Returns the 4 loss tensors.

    accumulate=True keeps the gradients already in the bucket (the reference's `subdivisions`, train.py:122-124: several
    backward passes per optimizer step); step=False skips the all-reduce and the optimizer step (all but the last
    sub-batch).  Do not call `optimizer.zero_grad()` between steps: the bucket is zeroed here."""
    bucket.reattach()
    if not accumulate:
        bucket.zero()
    preds = model(imgs)
    lbox, lobj, lcls, loss = compute_loss(preds, targets, cfg, imgs.device)
    loss.backward()
    if step:
        bucket.allreduce_mean(group)
        optimizer.step()
    return lbox, lobj, lcls, loss
