"""GPU: yfv2_ncnn_post (the ncnn sample's decode + per-class NMS, sample/ncnn/src/yolo-fastestv2.cpp:58-183) through the C ABI,
bit-exact against the reference's own C++ (tests/golden/ncnn_post.npz) and against the oracle restatement on fresh seeds."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import synth  # noqa: E402
from test_ncnn_post_cpu import cases  # noqa: E402

pytestmark = pytest.mark.gpu


def _eng():
    import yfv2  # noqa: F401
    import yfv2_engine
    return yfv2_engine


@pytest.mark.parametrize("c", list(cases()), ids=lambda c: c["name"])
def test_matches_reference_cpp_goldens(c):
    eng = _eng()
    o2 = torch.from_numpy(c["out2"])[None].cuda(); o3 = torch.from_numpy(c["out3"])[None].cuda()
    (b, s, k), = eng.ncnn_post(o2, o3, c["A"], float(c["thr"]), float(c["nms"]), (c["sw"], c["sh"]), anchors=c["anchors"].tolist())
    assert len(s) == len(c["scores"])
    assert np.array_equal(b, c["boxes"]) and np.array_equal(k, c["cates"])
    assert np.array_equal(s.view(np.uint32), c["scores"].view(np.uint32))


def test_batch_from_the_network_matches_the_oracle():
    """Whole deploy path on the device: forward -> export_onnx head -> ncnn post-process, image by image against the oracle run on
    the SAME export tensors (so the comparison is exact), incl. truncation at max_out."""
    eng = _eng()
    from oracle import ncnn_post as onp
    import model.detector as det
    sd = synth.make_state_dict(9)
    m = det.Detector(80, 3, True, True)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = synth.make_images(10, 5, 160, 224).cuda()
    o2, o3 = m(x)
    assert o2.shape == (5, 10, 14, 95) and o3.shape == (5, 5, 7, 95)
    res = eng.ncnn_post(o2, o3, 3, 0.0005, 0.25, (448, 300))
    o2c, o3c = o2.cpu().numpy(), o3.cpu().numpy()
    sw, sh = np.float32(448) / np.float32(224), np.float32(300) / np.float32(160)
    total = 0
    for i in range(5):
        b, s, k = onp.ncnn_post(o2c[i], o3c[i], 3, 80, 224, 160, onp.ANCHORS_COCO, np.float32(0.0005), np.float32(0.25), sw, sh)
        assert len(s) == len(res[i][1]) > 0
        assert np.array_equal(b, res[i][0]) and np.array_equal(k, res[i][2]) and np.array_equal(s, res[i][1])
        total += len(s)
    cut = eng.ncnn_post(o2, o3, 3, 0.0005, 0.25, (448, 300), max_out=7)
    for i in range(5):
        assert np.array_equal(cut[i][0], res[i][0][:7]) and np.array_equal(cut[i][1], res[i][1][:7])
    assert total > 50


def test_dense_same_class_logits_match_the_oracle():
    """Crowded case: three dominant classes at 352x352, 1815 candidates, hundreds of same-class overlaps per image."""
    eng = _eng()
    from oracle import ncnn_post as onp
    p = list(synth.make_head_logits(77, 3, 352, 352))
    for i in (2, 5):
        p[i][:, :3] += 6.0
    o2, o3 = eng.export_heads([t.cuda() for t in p])
    res = eng.ncnn_post(o2, o3, 3, 0.001, 0.25)
    o2c, o3c = o2.cpu().numpy(), o3.cpu().numpy()
    for i in range(3):
        b, s, k = onp.ncnn_post(o2c[i], o3c[i], 3, 80, 352, 352, onp.ANCHORS_COCO, np.float32(0.001), np.float32(0.25), np.float32(1), np.float32(1))
        assert 100 < len(s) == len(res[i][1])
        assert np.array_equal(b, res[i][0]) and np.array_equal(k, res[i][2]) and np.array_equal(s, res[i][1])


def test_rejects_cpu_tensors():
    eng = _eng()
    with pytest.raises(RuntimeError):
        eng.ncnn_post(torch.zeros(1, 2, 2, 95), torch.zeros(1, 1, 1, 95), 3)
