"""GPU parity of decode / NMS / fused decode+NMS (through the C ABI).
NMS: bit-exact against the reference's golden rows and the oracle's indices on identical inputs.
decode: fp32 tolerance 1e-4 (GPU expf vs the CPU's vectorised exp differ in the last ulp)."""
import os

import numpy as np
import pytest
import torch

import yfv2  # noqa: F401
import synth
import yfv2_engine as eng
from oracle import post as opost

pytestmark = pytest.mark.gpu
THR = [(0.3, 0.4), (0.01, 0.4), (0.001, 0.4), (0.25, 0.45)]


def split_rows(rows, counts):
    out, o = [], 0
    for c in counts:
        out.append(rows[o:o + c]); o += c
    return out


@pytest.mark.parametrize("tag", list(synth.NMS_CASES))
def test_nms_bit_exact_against_reference_golden(golden_dir, tag):
    g = dict(np.load(os.path.join(golden_dir, "nms_cases.npz")))
    dets = synth.make_dets(**synth.NMS_CASES[tag])
    d = dets.cuda()
    for ct, it in THR:
        out, counts, idx = eng.nms(d, ct, it)
        out, counts, idx = out.cpu().numpy(), counts.cpu().numpy(), idx.cpu().numpy()
        ref_counts = g["%s_%g_%g_counts" % (tag, ct, it)]
        assert list(counts) == list(ref_counts), (ct, it)
        ref = split_rows(g["%s_%g_%g_rows" % (tag, ct, it)], ref_counts)
        _, oidx = opost.nms(dets, ct, it, return_indices=True)
        for i, c in enumerate(counts):
            assert np.array_equal(out[i, :c], ref[i]), (tag, ct, it, i)        # bit-exact rows
            assert np.array_equal(idx[i, :c], oidx[i]), (tag, ct, it, i)       # bit-exact indices
            assert np.all(out[i, c:] == 0) and np.all(idx[i, c:] == -1)


def test_nms_class_filter(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "nms_cases.npz")))
    dets = synth.make_dets(**synth.NMS_CASES["dense"])[:1].cuda()
    out, counts, _ = eng.nms(dets, 0.01, 0.4, classes=[0, 5, 17])
    assert np.array_equal(out[0, : int(counts[0])].cpu().numpy(), g["filter_rows"])


def test_nms_empty_and_single():
    d = synth.make_dets(61, 2, 50).cuda()
    out, counts, idx = eng.nms(d, 2.0, 0.4)                 # nothing passes
    assert counts.tolist() == [0, 0] and float(out.abs().sum()) == 0.0
    d1 = synth.make_dets(62, 1, 1).cuda()
    d1[0, 0, 4] = 0.9; d1[0, 0, 5:] = 0.0; d1[0, 0, 7] = 0.8
    out, counts, idx = eng.nms(d1, 0.3, 0.4)
    assert counts.tolist() == [1] and int(idx[0, 0]) == 0 and float(out[0, 0, 5]) == 2.0


def test_decode_against_reference_golden(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "net_small.npz")))
    preds = [torch.from_numpy(g["pred%d" % i]).cuda() for i in range(6)]
    d = eng.decode(preds, synth.coco_cfg(96, 64)).cpu().numpy()
    assert d.shape == g["decode"].shape
    np.testing.assert_allclose(d, g["decode"], rtol=1e-4, atol=1e-6)
    g2 = dict(np.load(os.path.join(golden_dir, "post_cases.npz")))
    for tag, kw, hw, n in (("dense", {}, (352, 352), 3), ("sparse", {"obj_mean": -6.0}, (352, 352), 3), ("dense640", {}, (640, 640), 1)):
        preds = [p.cuda() for p in synth.make_head_logits(31, n, hw[0], hw[1], **kw)]
        d = eng.decode(preds, synth.coco_cfg(hw[1], hw[0])).cpu().numpy()
        np.testing.assert_allclose(d[:, ::37], g2[tag + "_decode_sample"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("kw,hw,n", [({}, (352, 352), 4), ({"obj_mean": -6.0}, (352, 352), 4), ({}, (640, 640), 2),
                                     ({"classes": 20, "anchor_num": 2}, (96, 160), 3)])
def test_decode_and_chain_against_oracle(kw, hw, n):
    cpu = synth.make_head_logits(71, n, hw[0], hw[1], **kw)
    cfg = synth.coco_cfg(hw[1], hw[0], kw.get("classes", 80))
    if "anchor_num" in kw:
        cfg["anchor_num"] = kw["anchor_num"]; cfg["anchors"] = cfg["anchors"][:8]
    ref = opost.decode(cpu, cfg)
    gpu = [p.cuda() for p in cpu]
    d = eng.decode(gpu, cfg)
    np.testing.assert_allclose(d.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=1e-6)
    for ct, it in THR[:3]:
        # stage-wise NMS on the GPU's own decode must equal the oracle NMS on that same tensor, bit for bit
        out, counts, idx = eng.nms(d, ct, it)
        rows, oidx = opost.nms(d.cpu(), ct, it, return_indices=True)
        f_out, f_counts, f_idx = eng.decode_nms(gpu, cfg, ct, it, want_idx=True)
        assert torch.equal(counts, f_counts) and torch.equal(out, f_out) and torch.equal(idx, f_idx)   # fused == unfused
        for i in range(n):
            c = int(counts[i])
            assert c == rows[i].shape[0]
            assert np.array_equal(out[i, :c].cpu().numpy(), rows[i].numpy())
            assert np.array_equal(idx[i, :c].cpu().numpy(), oidx[i])


def test_nms_microbench_properties_at_scale():
    """BASELINE config[4] shape (many images, 1815 candidates, conf 0.001): size-independent properties —
    (1) every image of a replicated batch gives the identical result, (2) NMS is idempotent: feeding the
    kept boxes back (as one-hot candidates) keeps all of them in the same order, (3) cap of 300 holds."""
    base = synth.make_dets(81, 4, 1815).cuda()
    big = base.repeat(128, 1, 1)                        # 512 images
    out, counts, idx = eng.nms(big, 0.001, 0.4)
    assert int(counts.max()) <= 300
    o4, c4, i4 = eng.nms(base, 0.001, 0.4)
    assert torch.equal(out.view(128, 4, 300, 6), o4.unsqueeze(0).expand(128, -1, -1, -1))
    assert torch.equal(counts.view(128, 4), c4.unsqueeze(0).expand(128, -1))
    # idempotence
    n = int(c4[0])
    kept = o4[0, :n]
    again = torch.zeros(1, n, 85, device="cuda")
    again[0, :, 0] = (kept[:, 0] + kept[:, 2]) / 2; again[0, :, 1] = (kept[:, 1] + kept[:, 3]) / 2
    again[0, :, 2] = kept[:, 2] - kept[:, 0]; again[0, :, 3] = kept[:, 3] - kept[:, 1]
    again[0, :, 4] = 1.0
    again[0, torch.arange(n), 5 + kept[:, 5].long()] = kept[:, 4]
    o2, c2, i2 = eng.nms(again, 0.001, 0.4)
    assert int(c2[0]) == n and torch.equal(i2[0, :n].long().cpu(), torch.arange(n))


def test_nms_boxes_wider_than_the_class_offset():
    """Boxes outside (-max_wh/2, max_wh/2) can overlap across classes after the class offset of utils/utils.py:283; the kernel
    then must not skip pairs on their class ids (the by-class shortcut is only taken for images whose boxes all fit).  One image
    of each kind in the same batch, bit-exact against the oracle."""
    d = synth.make_dets(91, 2, 400)
    d[1, :, 2] = d[1, :, 2] * 40.0 + 3000.0            # widths of several thousand pixels: neighbouring classes intersect
    d[1, :, 0] = d[1, :, 0] * 3.0
    for ct, it in THR[:3]:
        out, counts, idx = eng.nms(d.cuda(), ct, it)
        rows, oidx = opost.nms(d, ct, it, return_indices=True)
        for i in range(2):
            c = int(counts[i])
            assert c == rows[i].shape[0], (ct, it, i)
            assert np.array_equal(out[i, :c].cpu().numpy(), rows[i].numpy()), (ct, it, i)
            assert np.array_equal(idx[i, :c].cpu().numpy(), oidx[i]), (ct, it, i)
