"""Pins the CPU oracle (oracle/) against golden vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import synth
from oracle import net as onet
from oracle import post as opost
from oracle import loss as oloss

THR = [(0.3, 0.4), (0.01, 0.4), (0.001, 0.4)]


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def split_rows(rows, counts):
    out, o = [], 0
    for c in counts:
        out.append(rows[o:o + c]); o += c
    return out


def check_post(g, prefix, dets, thr_list, exact=True):
    """exact=True: dets are bit-identical to what the reference NMS saw -> rows must be bit-equal.
    exact=False: dets came from the oracle's own decode (1-ulp sigmoid/softmax differences from the
    reference's per-image non-contiguous calls), so compare kept rows within 1e-5."""
    for ct, it in thr_list:
        counts = g["%snms_%g_%g_counts" % (prefix, ct, it)]
        ref = split_rows(g["%snms_%g_%g_rows" % (prefix, ct, it)], counts)
        for impl in ("c", "numpy"):
            got = opost.nms(dets, ct, it, impl=impl)
            assert [r.shape[0] for r in got] == list(counts), (impl, ct, it)
            for a, b in zip(got, ref):
                if exact:
                    assert np.array_equal(a.numpy(), b), (impl, ct, it)
                else:
                    assert np.array_equal(a.numpy()[:, 5], b[:, 5])
                    np.testing.assert_allclose(a.numpy(), b, rtol=1e-5, atol=1e-5)


def test_net_small_eval_taps_and_post(golden_dir):
    g = load(golden_dir, "net_small.npz")
    sd = synth.make_state_dict(11)
    x = synth.make_images(12, 2, 64, 96)
    taps = {}
    with torch.no_grad():
        preds = onet.forward(sd, x, taps=taps)
    for i, p in enumerate(preds):
        np.testing.assert_allclose(p.numpy(), g["pred%d" % i], rtol=1e-5, atol=1e-5)
    for k in ("stem", "stage2.0", "stage2.1", "stage2.3", "stage3.0", "stage3.7", "stage4.0", "stage4.3", "S2", "S3"):
        np.testing.assert_allclose(taps[k].numpy(), g["tap_" + k], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(taps["cls_2"].numpy(), g["tap_cls_head_2"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(taps["reg_3"].numpy(), g["tap_reg_head_3"], rtol=1e-5, atol=1e-5)
    # decode from the REFERENCE's preds: identical arithmetic -> tight tolerance
    ref_preds = tuple(torch.from_numpy(g["pred%d" % i]) for i in range(6))
    dets = opost.decode(ref_preds, synth.coco_cfg(96, 64))
    np.testing.assert_allclose(dets.numpy(), g["decode"], rtol=1e-6, atol=1e-6)
    check_post(g, "", torch.from_numpy(g["decode"]), THR)
    with torch.no_grad():
        e2, e3 = onet.forward_export(sd, x)
    np.testing.assert_allclose(e2.numpy(), g["export_2"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(e3.numpy(), g["export_3"], rtol=1e-5, atol=1e-5)


def test_net_small_train_mode(golden_dir):
    g = load(golden_dir, "net_small.npz")
    sd = synth.make_state_dict(11)
    x = synth.make_images(12, 2, 64, 96)
    with torch.no_grad():
        preds = onet.forward(sd, x, training=True, update_running=True)
    for i, p in enumerate(preds):
        np.testing.assert_allclose(p.numpy(), g["train_pred%d" % i], rtol=1e-4, atol=1e-4)
    for k in ("backbone.first_conv.1", "backbone.stage3.2.branch_main.4", "fpn.cls_head_2.block.9"):
        np.testing.assert_allclose(sd[k + ".running_mean"].numpy(), g["train_rm_" + k], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(sd[k + ".running_var"].numpy(), g["train_rv_" + k], rtol=1e-4, atol=1e-5)


def test_net_352(golden_dir):
    g = load(golden_dir, "net_352.npz")
    sd = synth.make_state_dict(21)
    x = synth.make_images(22, 1, 352, 352)
    with torch.no_grad():
        preds = onet.forward(sd, x)
    for i, p in enumerate(preds):
        np.testing.assert_allclose(p.numpy(), g["pred%d" % i], rtol=1e-5, atol=1e-5)
    dets = opost.decode(tuple(torch.from_numpy(g["pred%d" % i]) for i in range(6)), synth.coco_cfg())
    check_post(g, "", dets, THR, exact=False)


def test_modelzoo_known_answers(golden_dir):
    """img/000139_result.png shows person .87, bicycle .46, person .32 (README.md:31-35)."""
    g = load(golden_dir, "images_modelzoo.npz")
    w = load(golden_dir, "modelzoo_weights.npz")
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    for name in ("000139", "000004"):
        x = torch.from_numpy(g[name + "_u8"]).float() / 255.0
        with torch.no_grad():
            preds = onet.forward(sd, x)
        for i, p in enumerate(preds):
            np.testing.assert_allclose(p.numpy(), g["%s_pred%d" % (name, i)], rtol=1e-5, atol=2e-5)
        dets = opost.decode(tuple(torch.from_numpy(g["%s_pred%d" % (name, i)]) for i in range(6)), synth.coco_cfg())
        check_post(g, name + "_", dets, [(0.3, 0.4), (0.001, 0.4)], exact=False)
    rows = g["000139_nms_0.3_0.4_rows"]
    assert [(int(r[5]), "%.2f" % r[4]) for r in rows] == [(0, "0.87"), (1, "0.46"), (0, "0.32")]
    rows = g["000004_nms_0.3_0.4_rows"]
    assert ["%.2f" % r[4] for r in rows] == ["0.87", "0.85", "0.76", "0.75", "0.68", "0.60", "0.56", "0.47", "0.33"]
    assert all(int(r[5]) == 2 for r in rows)       # nine cars


@pytest.mark.parametrize("tag,kw,hw,n", [("dense", {}, (352, 352), 3), ("sparse", {"obj_mean": -6.0}, (352, 352), 3),
                                         ("dense640", {}, (640, 640), 1)])
def test_post_cases(golden_dir, tag, kw, hw, n):
    g = load(golden_dir, "post_cases.npz")
    preds = synth.make_head_logits(31, n, hw[0], hw[1], **kw)
    dets = opost.decode(preds, synth.coco_cfg(hw[1], hw[0]))
    np.testing.assert_allclose(dets.numpy()[:, ::37], g[tag + "_decode_sample"], rtol=1e-6, atol=1e-6)
    check_post(g, tag + "_", dets, THR, exact=False)


@pytest.mark.parametrize("tag", list(synth.NMS_CASES))
def test_nms_cases_bit_exact(golden_dir, tag):
    g = load(golden_dir, "nms_cases.npz")
    dets = synth.make_dets(**synth.NMS_CASES[tag])
    for ct, it in THR + [(0.25, 0.45)]:
        counts = g["%s_%g_%g_counts" % (tag, ct, it)]
        ref = split_rows(g["%s_%g_%g_rows" % (tag, ct, it)], counts)
        impls = ("c", "numpy") if dets.shape[1] <= 2000 else ("c",)
        for impl in impls:
            got, idx = opost.nms(dets, ct, it, impl=impl, return_indices=True)
            assert [r.shape[0] for r in got] == list(counts)
            for a, b, ii, d in zip(got, ref, idx, dets):
                assert np.array_equal(a.numpy(), b), (impl, ct, it)
                # the returned indices really address the source rows
                src = d.numpy()[ii]
                assert np.array_equal((src[:, 5:] * src[:, 4:5]).argmax(1).astype(np.float32), a.numpy()[:, 5])


def test_nms_class_filter(golden_dir):
    g = load(golden_dir, "nms_cases.npz")
    dets = synth.make_dets(**synth.NMS_CASES["dense"])[:1]
    for impl in ("c", "numpy"):
        got = opost.nms(dets, 0.01, 0.4, classes=[0, 5, 17], impl=impl)[0]
        assert np.array_equal(got.numpy(), g["filter_rows"])


def test_nms_matches_installed_torchvision():
    """torchvision.ops.nms is the third-party kernel the reference calls (utils/utils.py:286)."""
    tv = pytest.importorskip("torchvision")
    rs = np.random.RandomState(5)
    for trial in range(30):
        n = int(rs.randint(1, 400))
        xy = rs.rand(n, 2).astype(np.float32) * 300
        wh = rs.rand(n, 2).astype(np.float32) * 120 + 1
        boxes = np.concatenate((xy, xy + wh), 1)
        if trial % 3 == 0:                                   # exact duplicates and score ties
            boxes[n // 2:] = boxes[: n - n // 2]
        scores = rs.rand(n).astype(np.float32)
        if trial % 2 == 0:
            scores = np.round(scores * 8) / 8
        for thr in (0.4, 0.45, 0.5):
            ref = tv.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
            got = opost.greedy_nms_numpy(boxes, scores, thr)
            assert np.array_equal(ref, got)
    # iou exactly float32(0.4) must be suppressed: threshold is compared in double (SURVEY 7 hard part 4)
    b = np.array([[0, 0, 10, 10], [0, 0, 10, 4]], np.float32)
    ref = tv.ops.nms(torch.from_numpy(b), torch.tensor([0.9, 0.8]), 0.4).numpy()
    assert np.array_equal(ref, opost.greedy_nms_numpy(b, np.array([0.9, 0.8], np.float32), 0.4))
    assert list(ref) == [0]


@pytest.mark.parametrize("tag,n,tseed", [("a", 2, 41), ("b", 3, 43)])
def test_loss_cases(golden_dir, tag, n, tseed):
    g = load(golden_dir, "loss_cases.npz")
    preds = [p.clone().requires_grad_(True) for p in synth.make_head_logits(40 + n, n, 352, 352, obj_std=1.0)]
    targets = synth.make_targets(tseed, n)
    cfg = synth.coco_cfg()
    tcls, tbox, indices, anch = oloss.build_target(preds, targets, cfg)
    for L in range(2):
        assert np.array_equal(tcls[L].numpy(), g["%s_tcls%d" % (tag, L)])
        assert np.array_equal(tbox[L].numpy(), g["%s_tbox%d" % (tag, L)])
        assert np.array_equal(anch[L].numpy(), g["%s_anch%d" % (tag, L)])
        assert np.array_equal(np.stack([t.numpy() for t in indices[L]], 0), g["%s_idx%d" % (tag, L)])
    lb, lo, lc, loss = oloss.compute_loss(preds, targets, cfg)
    loss.backward()
    np.testing.assert_allclose([lb.item(), lo.item(), lc.item(), loss.item()], g[tag + "_losses"], rtol=1e-6)
    for i, p in enumerate(preds):
        np.testing.assert_allclose(p.grad.numpy(), g["%s_grad%d" % (tag, i)], rtol=1e-5, atol=1e-9)


def test_loss_no_targets(golden_dir):
    g = load(golden_dir, "loss_cases.npz")
    preds = list(synth.make_head_logits(45, 2, 352, 352))
    out = oloss.compute_loss(preds, torch.zeros(0, 6), synth.coco_cfg())
    np.testing.assert_allclose([t.item() for t in out], g["empty_losses"], rtol=1e-6)
