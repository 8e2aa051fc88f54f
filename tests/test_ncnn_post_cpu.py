"""CPU: oracle/ncnn_post.c (restatement of the ncnn sample's decode + per-class NMS, sample/ncnn/src/yolo-fastestv2.cpp:58-183)
against tests/golden/ncnn_post.npz, which holds outputs of the reference's own C++ compiled in place (make_golden_ncnn.py)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    g = np.load(os.path.join(HERE, "golden", "ncnn_post.npz"))
    for name in g["names"]:
        name = str(name)
        A, C, iw, ih, sw, sh = [int(v) for v in g[name + "_params"]]
        thr, nms = [float(v) for v in g[name + "_fparams"]]
        yield dict(name=name, out2=g[name + "_out2"], out3=g[name + "_out3"], A=A, C=C, iw=iw, ih=ih, sw=sw, sh=sh, thr=np.float32(thr),
                   nms=np.float32(nms), anchors=g[name + "_anchors"], boxes=g[name + "_boxes"], scores=g[name + "_scores"], cates=g[name + "_cates"])


@pytest.mark.parametrize("c", list(cases()), ids=lambda c: c["name"])
def test_oracle_matches_reference_cpp(c):
    from oracle import ncnn_post as onp
    scale_w, scale_h = np.float32(c["sw"]) / np.float32(c["iw"]), np.float32(c["sh"]) / np.float32(c["ih"])      # .cpp:189-190 (float division)
    b, s, k = onp.ncnn_post(c["out2"], c["out3"], c["A"], c["C"], c["iw"], c["ih"], c["anchors"], c["thr"], c["nms"], scale_w, scale_h)
    assert len(s) == len(c["scores"]) > 0
    assert np.array_equal(b, c["boxes"]) and np.array_equal(k, c["cates"])
    assert np.array_equal(s.view(np.uint32), c["scores"].view(np.uint32))


def test_known_answers_of_the_bundled_image():
    # img/000139_result.png: person .87, bicycle .46 -- the deploy path finds the same three objects as test.py
    c = [c for c in cases() if c["name"] == "000139"][0]
    assert [int(v) for v in c["cates"]] == [0, 1, 0]
    assert [round(float(v), 2) for v in c["scores"]] == [0.87, 0.46, 0.32]


@pytest.mark.skipif(not os.path.exists("/root/reference/sample/ncnn/src/yolo-fastestv2.cpp"), reason="reference checkout not mounted")
def test_oracle_matches_live_reference_on_fresh_seeds():
    """Where /root/reference is mounted: build oracle/_ref and compare on seeds that are not in the fixture."""
    import ctypes
    import subprocess
    import sys
    import torch
    sys.path.insert(0, HERE)
    import synth
    from oracle import ncnn_post as onp
    root = os.path.dirname(HERE)
    subprocess.run(["make", "-C", os.path.join(root, "oracle"), "ref"], check=True, capture_output=True)
    ref = ctypes.CDLL(os.path.join(root, "oracle", "_ref", "libncnn_post_ref.so"))
    ref.ncnn_ref_detect.restype = ctypes.c_int
    ref.ncnn_ref_detect.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    ref.ncnn_ref_configure.restype = None
    ref.ncnn_ref_configure.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
    for seed in range(40, 46):
        p = list(synth.make_head_logits(seed, 1, 160, 224, obj_mean=-0.5))
        for i in (2, 5):
            p[i][:, :2] += 5.0
        outs = []
        for lv in range(2):
            o = torch.cat((p[3 * lv].sigmoid(), p[3 * lv + 1].sigmoid(), torch.softmax(p[3 * lv + 2], 1)), 1).permute(0, 2, 3, 1)
            outs.append(np.ascontiguousarray(o[0].numpy()))
        anc = onp.ANCHORS_COCO
        thr, nms = np.float32(0.01), np.float32(0.3)
        ref.ncnn_ref_configure(3, 80, 224, 160, nms, anc.ctypes.data)
        cap = 3 * (10 * 14 + 5 * 7)
        boxes = np.zeros((cap, 4), np.int32); scores = np.zeros(cap, np.float32); cates = np.zeros(cap, np.int32)
        n = ref.ncnn_ref_detect(outs[0].ctypes.data, 10, 14, outs[1].ctypes.data, 5, 7, 95, 448, 320, thr, cap, boxes.ctypes.data,
                                scores.ctypes.data, cates.ctypes.data)
        b, s, k = onp.ncnn_post(outs[0], outs[1], 3, 80, 224, 160, anc, thr, nms, np.float32(2.0), np.float32(2.0))
        assert n == len(s) and np.array_equal(b, boxes[:n]) and np.array_equal(k, cates[:n]) and np.array_equal(s, scores[:n])
