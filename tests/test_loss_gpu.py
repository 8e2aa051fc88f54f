"""GPU parity of the device loss (csrc/k_loss.cu through the C ABI): build_target rows bit-exact against the reference's
goldens, the four loss scalars and d(loss)/d(preds) against the reference's autograd."""
import os

import numpy as np
import pytest
import torch

import yfv2  # noqa: F401
import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,n,tseed", [("a", 2, 41), ("b", 3, 43)])
def test_loss_against_reference_golden(golden_dir, tag, n, tseed):
    import utils.loss as ul
    g = dict(np.load(os.path.join(golden_dir, "loss_cases.npz")))
    preds = [p.cuda().requires_grad_(True) for p in synth.make_head_logits(40 + n, n, 352, 352, obj_std=1.0)]
    targets = synth.make_targets(tseed, n).cuda()
    cfg = synth.coco_cfg()
    tcls, tbox, indices, anch = ul.build_target(preds, targets, cfg, "cuda")
    for L in range(2):
        assert np.array_equal(tcls[L].cpu().numpy(), g["%s_tcls%d" % (tag, L)])
        assert np.array_equal(tbox[L].cpu().numpy(), g["%s_tbox%d" % (tag, L)])            # bit-exact fp32
        assert np.array_equal(anch[L].cpu().numpy(), g["%s_anch%d" % (tag, L)])            # bit-exact fp64
        assert np.array_equal(np.stack([t.cpu().numpy() for t in indices[L]], 0), g["%s_idx%d" % (tag, L)])
    lb, lo, lc, loss = ul.compute_loss(preds, targets, cfg, "cuda")
    assert lb.shape == (1,) and loss.shape == (1,)
    np.testing.assert_allclose([lb.item(), lo.item(), lc.item(), loss.item()], g[tag + "_losses"], rtol=1e-5)
    loss.backward()
    for i, p in enumerate(preds):
        ref = g["%s_grad%d" % (tag, i)]
        np.testing.assert_allclose(p.grad.cpu().numpy(), ref, rtol=1e-4, atol=1e-6 * max(1e-3, np.abs(ref).max()), err_msg="grad%d" % i)


def test_loss_without_targets(golden_dir):
    import utils.loss as ul
    g = dict(np.load(os.path.join(golden_dir, "loss_cases.npz")))
    preds = [p.cuda() for p in synth.make_head_logits(45, 2, 352, 352)]
    out = ul.compute_loss(preds, torch.zeros(0, 6).cuda(), synth.coco_cfg(), "cuda")
    np.testing.assert_allclose([t.item() for t in out], g["empty_losses"], rtol=1e-5)


def test_loss_against_oracle_at_training_batch():
    """BASELINE config[2] per-GPU shard: batch 64, ~7 boxes per image."""
    import utils.loss as ul
    from oracle import loss as oloss
    cpu = [p.clone().requires_grad_(True) for p in synth.make_head_logits(91, 64, 352, 352, obj_std=1.0)]
    targets = synth.make_targets(92, 64)
    cfg = synth.coco_cfg()
    ref = oloss.compute_loss(cpu, targets, cfg)
    ref[3].backward()
    gpu = [p.detach().cuda().requires_grad_(True) for p in cpu]
    out = ul.compute_loss(gpu, targets.cuda(), cfg, "cuda")
    np.testing.assert_allclose([t.item() for t in out], [t.item() for t in ref], rtol=1e-5)
    out[3].backward()
    for a, b in zip(gpu, cpu):
        r = b.grad.numpy()
        np.testing.assert_allclose(a.grad.cpu().numpy(), r, rtol=1e-4, atol=1e-6 * np.abs(r).max())
