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
    # batch 2 @64x96: the stride-32 BatchNorms normalise over 2*2*3 = 12 samples, which amplifies fp32 summation-order
    # noise of the reference itself (our statistics are accumulated in fp64) — 5e-4 here, 1e-4 in the larger case below
    for i, p in enumerate(preds):
        np.testing.assert_allclose(p.detach().cpu().numpy(), g["train_pred%d" % i], rtol=5e-4, atol=5e-4, err_msg="pred%d" % i)
    sd = m.state_dict()
    for k in ("backbone.first_conv.1", "backbone.stage3.2.branch_main.4", "fpn.cls_head_2.block.9"):
        np.testing.assert_allclose(sd[k + ".running_mean"].cpu().numpy(), g["train_rm_" + k], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(sd[k + ".running_var"].cpu().numpy(), g["train_rv_" + k], rtol=1e-4, atol=1e-5)
        assert int(sd[k + ".num_batches_tracked"]) == 1


def _oracle_grads(sd, x, targets, cfg, dtype):
    osd = {k: (v.clone().to(dtype).requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    preds = onet.forward(osd, x.to(dtype), training=True, update_running=False)
    losses = oloss.compute_loss(preds, targets, cfg)
    losses[3].backward()
    return preds, losses, {k: v.grad.double() for k, v in osd.items() if getattr(v, "grad", None) is not None}


def test_all_parameter_gradients_against_oracle():
    """Train-mode BN plus ReLU / max-pool gates make the gradients chaotic in fp32: the fp32 reference differs from an fp64
    copy of itself by several 1e-3 in global relative L2 (SURVEY 7 hard part 7).  So the yardstick is the fp64 oracle:
    our fp32 kernels must be as close to it as the reference's own fp32 path is."""
    import utils.loss as ul
    sd = synth.make_state_dict(61)
    x = synth.make_images(62, 8, 128, 160)
    targets = synth.make_targets(63, 8)
    cfg = synth.coco_cfg(160, 128)
    ref_preds, ref_losses, g32 = _oracle_grads(sd, x, targets, cfg, torch.float32)
    _, _, g64 = _oracle_grads(sd, x, targets, cfg, torch.float64)
    m = make_model(sd)
    preds = m(x.cuda())
    for p, r in zip(preds, ref_preds):      # train-mode BN over 160 samples at stride 32: fp32 reorder noise reaches ~1e-4
        np.testing.assert_allclose(p.detach().cpu().numpy(), r.detach().numpy(), rtol=3e-4, atol=3e-4)
    losses = ul.compute_loss(preds, targets.cuda(), cfg, "cuda")
    np.testing.assert_allclose([t.item() for t in losses], [t.item() for t in ref_losses], rtol=1e-4)
    losses[3].backward()
    ours = {n: p.grad.cpu().double() for n, p in m.named_parameters()}
    assert set(ours) == set(g64) and len(ours) == 225

    def rel(a, b):
        num = sum(float((a[k] - b[k]).pow(2).sum()) for k in b)
        return (num / sum(float(b[k].pow(2).sum()) for k in b)) ** 0.5

    e_ours, e_ref = rel(ours, g64), rel(g32, g64)
    print("global relative L2 vs fp64 oracle: ours %.3e, fp32 oracle %.3e" % (e_ours, e_ref))
    assert e_ours <= max(2.0 * e_ref, 1e-4), (e_ours, e_ref)
    gmax = max(float(v.abs().max()) for v in g64.values())
    worst = []
    for k in g64:
        floor = 1e-5 * gmax * g64[k].numel() ** 0.5
        d64 = max(float(g64[k].norm()), floor)
        worst.append((float((ours[k] - g64[k]).norm()) / d64, float((g32[k] - g64[k]).norm()) / d64, k))
    worst.sort(reverse=True)
    print("worst tensors (ours, fp32 oracle):", worst[:5])
    # Noise enters at different layers in different implementations (each op matches PyTorch's fp32 op to ~3e-7 in
    # isolation, tests/test_train_ops_gpu.py); per tensor we only require the same order of magnitude as the chaos floor.
    for eo, er, k in worst:
        assert eo <= max(4.0 * er, 1e-2), (k, eo, er)


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
