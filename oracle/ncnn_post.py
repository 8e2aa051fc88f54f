"""ctypes front of oracle/ncnn_post.c — the CPU restatement of the ncnn sample's decode + per-class NMS
(sample/ncnn/src/yolo-fastestv2.cpp:58-183).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes
import os

import numpy as np

_lib = None
ANCHORS_COCO = np.array([12.64, 19.39, 37.88, 51.48, 55.71, 138.31, 126.91, 78.23, 131.57, 214.55, 279.92, 258.87], np.float32)   # .cpp:34-35


def _load():
    global _lib
    if _lib is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.run(["make", "-C", here], check=True, capture_output=True)
        _lib = ctypes.CDLL(path)
        _lib.oracle_ncnn_post.restype = ctypes.c_int
        _lib.oracle_ncnn_post.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_float,
                                          ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def ncnn_post(out2, out3, A, C, in_w, in_h, anchors, thresh, nms_thresh, scale_w, scale_h):
    """out2 / out3: [h, w, 5A+C] float32 of ONE image.  Returns (boxes int32 [n,4], scores float32 [n], cates int32 [n])."""
    out2 = np.ascontiguousarray(out2, np.float32); out3 = np.ascontiguousarray(out3, np.float32)
    anchors = np.ascontiguousarray(anchors, np.float32)
    cap = A * (out2.shape[0] * out2.shape[1] + out3.shape[0] * out3.shape[1])
    boxes = np.zeros((max(cap, 1), 4), np.int32); scores = np.zeros(max(cap, 1), np.float32); cates = np.zeros(max(cap, 1), np.int32)
    n = _load().oracle_ncnn_post(out2.ctypes.data, out2.shape[0], out2.shape[1], out3.ctypes.data, out3.shape[0], out3.shape[1], A, C,
                                 in_w, in_h, anchors.ctypes.data, thresh, nms_thresh, scale_w, scale_h, cap, boxes.ctypes.data,
                                 scores.ctypes.data, cates.ctypes.data)
    assert n >= 0
    return boxes[:n].copy(), scores[:n].copy(), cates[:n].copy()
