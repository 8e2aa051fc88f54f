"""GPU regression tests for behaviours the round-1 review found missing: eval after a train-mode forward must fold
the UPDATED BatchNorm statistics, module copies / pickles must work after an eval forward, `.data` edits are picked
up after invalidate_packed(), an empty class filter keeps nothing (reference utils/utils.py:266-268), and only the
total loss is differentiable."""
import copy
import io

import numpy as np
import pytest
import torch

import yfv2  # noqa: F401
import synth
from oracle import net as onet

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-4, atol=1e-4)


def _model(sd):
    import model.detector as det
    m = det.Detector(80, 3, True)
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def test_eval_after_train_forward_uses_updated_running_stats():
    sd = synth.make_state_dict(31)
    x = synth.make_images(32, 2, 64, 96)
    m = _model(sd).eval()
    m(x.cuda())                                          # packs the weights with the original statistics
    m.train()
    with torch.no_grad():
        m(x.cuda())                                      # updates running_mean / running_var through raw pointers
    m.eval()
    got = m(x.cuda())
    ref_sd = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        onet.forward(ref_sd, x, training=True, update_running=True)     # the oracle's train-mode pass updates ref_sd in place
        want = onet.forward(ref_sd, x)
    stale = onet.forward({k: v.clone() for k, v in sd.items()}, x)
    assert not np.allclose(want[2].numpy(), stale[2].numpy(), **TOL), "test needs statistics that move the output"
    for g, w in zip(got, want):
        np.testing.assert_allclose(g.cpu().numpy(), w.numpy(), **TOL)


def test_data_edit_needs_invalidate_and_is_then_seen():
    sd = synth.make_state_dict(33)
    x = synth.make_images(34, 1, 64, 64)
    m = _model(sd).eval()
    a = [p.clone() for p in m(x.cuda())]
    m.output_cls_layers.bias.data.add_(1.0)              # does not bump tensor._version
    m.invalidate_packed()
    b = m(x.cuda())
    np.testing.assert_allclose(b[2].cpu().numpy(), a[2].cpu().numpy() + 1.0, rtol=1e-5, atol=1e-5)


def test_deepcopy_and_pickle_after_forward():
    sd = synth.make_state_dict(35)
    x = synth.make_images(36, 1, 64, 64).cuda()
    m = _model(sd).eval()
    want = m(x)
    m2 = copy.deepcopy(m)                                # e.g. an EMA copy
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False)
    for other in (m2, m3):
        got = other(x)
        for g, w in zip(got, want):
            assert torch.equal(g, w)


def test_empty_class_filter_keeps_nothing():
    import utils.utils as uu
    cfg = synth.coco_cfg(96, 64)
    m = _model(synth.make_state_dict(37)).eval()
    dets = uu.handel_preds(m(synth.make_images(38, 2, 64, 96).cuda()), cfg, "cuda")
    assert sum(o.shape[0] for o in uu.non_max_suppression(dets, 0.01, 0.4)) > 0
    out = uu.non_max_suppression(dets, 0.01, 0.4, classes=[])
    assert len(out) == 2 and all(o.shape[0] == 0 for o in out)


def test_partial_loss_backward_is_refused():
    import utils.loss as ul
    cfg = synth.coco_cfg(96, 64)
    m = _model(synth.make_state_dict(39)).train()
    x = synth.make_images(40, 2, 64, 96).cuda()
    targets = synth.make_targets(41, 2).cuda()
    lbox, lobj, lcls, loss = ul.compute_loss(m(x), targets, cfg, "cuda")
    with pytest.raises(RuntimeError, match="only the total loss"):
        lbox.backward(retain_graph=True)
    loss.backward()
    assert all(p.grad is not None for p in m.parameters())


def test_bucket_survives_zero_grad_set_to_none():
    import train_ddp
    import utils.loss as ul
    cfg = synth.coco_cfg(96, 64)
    m = _model(synth.make_state_dict(43)).train()
    x = synth.make_images(44, 2, 64, 96).cuda()
    targets = synth.make_targets(45, 2).cuda()
    bucket = train_ddp.FlatGradBucket(m.parameters())
    opt = train_ddp.make_optimizer(m, 1e-3)
    train_ddp.train_step(m, bucket, opt, x, targets, cfg, ul.compute_loss)
    opt.zero_grad()                                      # set_to_none=True: detaches every .grad view
    train_ddp.train_step(m, bucket, opt, x, targets, cfg, ul.compute_loss)
    off = 0
    for p in bucket.params:
        assert p.grad.data_ptr() == bucket.flat[off:].data_ptr()
        off += p.numel()
    assert float(bucket.flat.abs().sum()) > 0


def test_head_tensors_that_are_only_4_byte_aligned():
    """The pixel-pair head kernel writes a pair as one 8-byte store when the destination allows it; caller-owned head tensors that
    are only 4-byte aligned must take the scalar path and still come out bit-identical."""
    import model.detector as det
    sd = synth.make_state_dict(3)
    m = det.Detector(80, 3, True)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = synth.make_images(4, 2, 352, 352).cuda()
    want = [p.clone() for p in m(x)]
    plan = m._plan_for(x)
    odd = []
    for p in want:                                   # same shapes, storage shifted by one float
        flat = torch.zeros(p.numel() + 3, dtype=torch.float32, device="cuda")
        view = flat[1:1 + p.numel()].view(p.shape)
        assert view.data_ptr() % 8 == 4
        odd.append(view)
    plan.forward(x, preds=tuple(odd))
    for a, b in zip(odd, want):
        assert torch.equal(a, b)
