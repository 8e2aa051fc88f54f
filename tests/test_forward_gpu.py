"""GPU parity of Detector.forward (through the C ABI) against the reference's golden outputs and the
CPU oracle.  Tolerance: north_star asks for 1e-4 on fp32 tensors; logits here are O(1)."""
import os

import numpy as np
import pytest
import torch

import yfv2  # noqa: F401
import synth
from oracle import net as onet

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-4, atol=1e-4)


def make_model(sd, classes=80, anchors=3):
    import model.detector as det
    m = det.Detector(classes, anchors, True)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def test_small_against_reference_golden(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "net_small.npz")))
    m = make_model(synth.make_state_dict(11))
    x = synth.make_images(12, 2, 64, 96)
    preds = m(x.cuda())
    torch.cuda.synchronize()
    for i, p in enumerate(preds):
        assert p.shape == g["pred%d" % i].shape
        np.testing.assert_allclose(p.cpu().numpy(), g["pred%d" % i], err_msg="pred%d" % i, **TOL)


def test_352_against_reference_golden(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "net_352.npz")))
    m = make_model(synth.make_state_dict(21))
    preds = m(synth.make_images(22, 1, 352, 352).cuda())
    for i, p in enumerate(preds):
        np.testing.assert_allclose(p.cpu().numpy(), g["pred%d" % i], err_msg="pred%d" % i, **TOL)


@pytest.mark.parametrize("n,h,w,seed", [(1, 32, 32, 1), (3, 96, 160, 2), (2, 352, 352, 3), (5, 224, 96, 4), (1, 640, 640, 5)])
def test_shapes_against_oracle(n, h, w, seed):
    sd = synth.make_state_dict(100 + seed)
    x = synth.make_images(200 + seed, n, h, w)
    with torch.no_grad():
        ref = onet.forward(sd, x)
    preds = make_model(sd)(x.cuda())
    for i, (p, r) in enumerate(zip(preds, ref)):
        np.testing.assert_allclose(p.cpu().numpy(), r.numpy(), err_msg="pred%d" % i, **TOL)


def test_other_class_and_anchor_counts():
    sd = synth.make_state_dict(7, classes=20, anchor_num=2)
    x = synth.make_images(8, 2, 128, 128)
    with torch.no_grad():
        ref = onet.forward(sd, x)
    preds = make_model(sd, 20, 2)(x.cuda())
    for p, r in zip(preds, ref):
        np.testing.assert_allclose(p.cpu().numpy(), r.numpy(), **TOL)


def test_uint8_input_fuses_the_255_division():
    sd = synth.make_state_dict(31)
    u8 = (synth.make_images(32, 2, 96, 96) * 255).to(torch.uint8)
    m = make_model(sd)
    a = m(u8.cuda())
    b = m((u8.float() / 255.0).cuda())
    with torch.no_grad():
        ref = onet.forward(sd, u8.float() / 255.0)
    for p, q, r in zip(a, b, ref):
        # the tensor-core stem takes the uint8 pixels as exact TF32 operands and folds 1/255 into the weights (two roundings of
        # the weight instead of one of the pixel): same result to fp32 round-off, and both paths within the bar of the reference
        np.testing.assert_allclose(p.cpu().numpy(), q.cpu().numpy(), rtol=5e-5, atol=5e-5)
        np.testing.assert_allclose(p.cpu().numpy(), r.numpy(), **TOL)


def test_modelzoo_known_answers(golden_dir):
    """Trained weights on the bundled images reproduce the reference's head tensors (and, through the
    device post-process, the boxes drawn in img/000139_result.png: person .87, bicycle .46, person .32)."""
    import utils.utils as uu
    g = dict(np.load(os.path.join(golden_dir, "images_modelzoo.npz")))
    w = dict(np.load(os.path.join(golden_dir, "modelzoo_weights.npz")))
    m = make_model({k: torch.from_numpy(v) for k, v in w.items()})
    cfg = synth.coco_cfg()
    for name in ("000139", "000004"):
        x = torch.from_numpy(g[name + "_u8"]).cuda()
        preds = m(x)
        for i, p in enumerate(preds):
            np.testing.assert_allclose(p.cpu().numpy(), g["%s_pred%d" % (name, i)], err_msg="%s pred%d" % (name, i), **TOL)
        rows = uu.non_max_suppression(uu.handel_preds(preds, cfg, x.device), 0.3, 0.4)[0].numpy()
        ref = g[name + "_nms_0.3_0.4_rows"]
        assert rows.shape == ref.shape
        assert np.array_equal(rows[:, 5], ref[:, 5])
        np.testing.assert_allclose(rows, ref, rtol=1e-4, atol=2e-3)
        fused = uu.detect(preds, cfg, 0.3, 0.4)[0].numpy()
        assert np.array_equal(fused, rows)
    rows = uu.detect(m(torch.from_numpy(g["000139_u8"]).cuda()), cfg, 0.3, 0.4)[0].numpy()
    assert [(int(r[5]), "%.2f" % r[4]) for r in rows] == [(0, "0.87"), (1, "0.46"), (0, "0.32")]


def test_batch_invariance_at_full_size():
    """Size-independent property at BASELINE config[1] scale: every image of a batch-256 forward equals
    the same image run alone (per-image work is independent, eval-mode BN)."""
    sd = synth.make_state_dict(41)
    m = make_model(sd)
    x = synth.make_images(42, 8, 352, 352).cuda()
    xb = x.repeat(32, 1, 1, 1)                       # 256 images
    big = m(xb)
    small = m(x)
    for p, q in zip(big, small):
        assert torch.equal(p[:8], q) and torch.equal(p[248:], q)
    one = m(x[3:4])
    for p, q in zip(one, small):
        assert torch.equal(p[0], q[3])


def test_every_fused_stage_against_reference_taps(golden_dir):
    """Stage-by-stage parity: run the forward one fused kernel at a time (yfv2_forward_range) and compare each
    block output, in the reference's logical channel order, with the activations hooked out of the real
    reference (tests/golden/net_small.npz tap_*)."""
    g = dict(np.load(os.path.join(golden_dir, "net_small.npz")))
    m = make_model(synth.make_state_dict(11))
    x = synth.make_images(12, 2, 64, 96).cuda()
    preds = m(x)
    plan = next(iter(m._plans.values()))
    names = ["stem"] + ["stage2.%d" % i for i in range(4)] + ["stage3.%d" % i for i in range(8)] + ["stage4.%d" % i for i in range(4)]
    stages = plan.stage_names                    # a block may be more than one launch ("stage4.1/pw1", "stage4.1/dwpw")
    done = 0
    for bi, name in enumerate(names):
        last = max(i for i, sn in enumerate(stages) if sn.split("/")[0] == name)
        plan.forward_range(x, preds, done, last + 1)
        done = last + 1
        got = plan.debug_gather(bi).cpu().numpy()
        np.testing.assert_allclose(got, g["tap_" + name], err_msg=name, **TOL)
    plan.forward_range(x, preds, done, len(stages))
    np.testing.assert_allclose(plan.debug_gather(18).cpu().numpy(), g["tap_S3"], **TOL)
    np.testing.assert_allclose(plan.debug_gather(17).cpu().numpy(), g["tap_S2"], **TOL)
    for i, p in enumerate(preds):
        np.testing.assert_allclose(p.cpu().numpy(), g["pred%d" % i], **TOL)


def test_export_onnx_head(golden_dir):
    """Detector(..., export_onnx=True): sigmoid(reg) | sigmoid(obj) | softmax(cls), channel-last (model/detector.py:33-44)."""
    import model.detector as det
    g = dict(np.load(os.path.join(golden_dir, "net_small.npz")))
    m = det.Detector(80, 3, True, True)
    m.load_state_dict(synth.make_state_dict(11))
    e2, e3 = m.cuda().eval()(synth.make_images(12, 2, 64, 96).cuda())
    assert e2.shape == g["export_2"].shape and e3.shape == g["export_3"].shape
    np.testing.assert_allclose(e2.cpu().numpy(), g["export_2"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(e3.cpu().numpy(), g["export_3"], rtol=1e-4, atol=1e-5)
