"""GPU parity of the training path (csrc/k_train.cu + csrc/k_loss.cu through the C ABI): train-mode forward against the
reference's golden activations, and all 225 parameter gradients of forward -> compute_loss -> backward against the CPU
oracle's autograd.  Gradient tolerance follows SURVEY 7 hard part 7 (the reference against itself differs by 1.5e-4 per
tensor): per-tensor relative L2 <= 1e-3 with an absolute floor of 1e-5 * max|grad|."""
import os

import numpy as np
import pytest
import torch

import yfv2  # noqa: F401
import synth
from oracle import net as onet
from oracle import loss as oloss

pytestmark = pytest.mark.gpu


def make_model(sd):
    import model.detector as det
    m = det.Detector(80, 3, True)
    m.load_state_dict(sd, strict=True)
    return m.cuda().train()


def test_train_forward_against_reference_golden(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "net_small.npz")))
    m = make_model(synth.make_state_dict(11))
    preds = m(synth.make_images(12, 2, 64, 96).cuda())
    for i, p in enumerate(preds):
        np.testing.assert_allclose(p.detach().cpu().numpy(), g["train_pred%d" % i], rtol=1e-4, atol=1e-4, err_msg="pred%d" % i)
    sd = m.state_dict()
    for k in ("backbone.first_conv.1", "backbone.stage3.2.branch_main.4", "fpn.cls_head_2.block.9"):
        np.testing.assert_allclose(sd[k + ".running_mean"].cpu().numpy(), g["train_rm_" + k], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(sd[k + ".running_var"].cpu().numpy(), g["train_rv_" + k], rtol=1e-4, atol=1e-5)
        assert int(sd[k + ".num_batches_tracked"]) == 1


def test_all_parameter_gradients_against_oracle():
    import utils.loss as ul
    sd = synth.make_state_dict(61)
    x = synth.make_images(62, 4, 128, 160)
    targets = synth.make_targets(63, 4)
    cfg = synth.coco_cfg(160, 128)
    # oracle: functional forward in train mode with autograd on every float parameter
    osd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    ref_preds = onet.forward(osd, x, training=True, update_running=False)
    ref_losses = oloss.compute_loss(ref_preds, targets, cfg)
    ref_losses[3].backward()
    m = make_model(sd)
    preds = m(x.cuda())
    for p, r in zip(preds, ref_preds):
        np.testing.assert_allclose(p.detach().cpu().numpy(), r.detach().numpy(), rtol=1e-4, atol=1e-4)
    losses = ul.compute_loss(preds, targets.cuda(), cfg, "cuda")
    np.testing.assert_allclose([t.item() for t in losses], [t.item() for t in ref_losses], rtol=1e-4)
    losses[3].backward()
    worst = 0.0
    gmax = max(float(v.grad.abs().max()) for k, v in osd.items() if getattr(v, "grad", None) is not None)
    for name, p in m.named_parameters():
        ref = osd[name].grad
        assert p.grad is not None and ref is not None, name
        got = p.grad.cpu()
        err = float((got - ref).norm())
        denom = max(float(ref.norm()), 1e-5 * gmax * ref.numel() ** 0.5)
        worst = max(worst, err / denom)
        assert err / denom <= 1e-3, (name, err, denom)
    print("worst per-tensor relative L2:", worst)


def test_sgd_step_with_flat_bucket_matches_plain_autograd():
    """The flat gradient bucket (what gets all-reduced) is just a different home for .grad: one SGD step from it equals
    one SGD step from ordinary per-parameter grads."""
    import train_ddp
    import utils.loss as ul
    sd = synth.make_state_dict(71)
    x = synth.make_images(72, 2, 96, 96).cuda()
    targets = synth.make_targets(73, 2).cuda()
    cfg = synth.coco_cfg(96, 96)
    a, b = make_model(sd), make_model(sd)
    bucket = train_ddp.FlatGradBucket(a.parameters())
    assert bucket.flat.numel() == 243095
    opt_a, opt_b = train_ddp.make_optimizer(a, 1e-3), train_ddp.make_optimizer(b, 1e-3)
    train_ddp.train_step(a, bucket, opt_a, x, targets, cfg, ul.compute_loss)
    lb = ul.compute_loss(b(x), targets, cfg, "cuda")
    lb[3].backward()
    opt_b.step()
    for (n1, p1), (_, p2) in zip(a.named_parameters(), b.named_parameters()):
        torch.testing.assert_close(p1, p2, rtol=1e-5, atol=1e-7, msg=n1)
