"""Drop-in replacement for the reference's utils/loss.py: compute_loss(preds, targets, cfg, device).

Same signature and return value — (lbox, lobj, lcls, loss), each a 1-element tensor, `loss.backward()` sends
d(loss)/d(preds) back into whatever produced the six head tensors — but target matching, CIoU, BCE and CE and their
gradients run in libyfv2.so (csrc/k_loss.cu) instead of ~60 host-launched tensor ops.  CUDA only.
"""
import os
import sys

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import yfv2_engine  # noqa: E402

layer_index = [0, 0, 0, 1, 1, 1]


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, targets, cfg, *preds):
        losses, dpreds = yfv2_engine.compute_loss(preds, targets, cfg, want_grads=True)
        ctx.save_for_backward(*dpreds)
        ctx.set_materialize_grads(False)
        return losses[0:1], losses[1:2], losses[2:3], losses[3:4]

    @staticmethod
    def backward(ctx, g_box, g_obj, g_cls, g_total):
        # The kernel differentiates the SUM (what train.py:110 back-propagates); the three parts are only logged.
        # Back-propagating a part would silently give zeros, so it is refused instead.
        if g_box is not None or g_obj is not None or g_cls is not None:
            raise RuntimeError("yfv2 compute_loss: only the total loss (4th return value) is differentiable; "
                               "lbox / lobj / lcls are returned for logging (reference train.py:110 back-propagates the sum)")
        dpreds = ctx.saved_tensors
        if g_total is None:
            return (None, None) + tuple(None for _ in dpreds)
        return (None, None) + tuple(d * g_total for d in dpreds)


def compute_loss(preds, targets, cfg, device):
    if not preds[0].is_cuda:
        raise RuntimeError("yfv2 compute_loss runs on CUDA only (no CPU fallback)")
    return _LossFn.apply(targets, cfg, *preds)


def build_target(preds, targets, cfg, device):
    """(tcls, tbox, indices, anch) per level, as the reference's build_target (utils/loss.py:53-124) returns them."""
    _, _, ws = yfv2_engine.compute_loss(preds, targets, cfg, want_grads=False, return_workspace=True)
    tcls, tbox, indices, anch = [], [], [], []
    for lv in range(len(preds) // 3):
        idx, tb, an, tc = yfv2_engine.read_targets(ws, lv)
        tcls.append(tc.long()); tbox.append(tb); anch.append(an)
        indices.append(tuple(idx[i].long() for i in range(4)))
    return tcls, tbox, indices, anch
