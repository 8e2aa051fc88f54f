#!/usr/bin/env python
"""BASELINE configs[0]: ONE 352x352 image, Detector.forward + decode + NMS (COCO 80 classes; modelzoo weights from
tests/golden/modelzoo_weights.npz, img/000139.jpg pre-resized from tests/golden/images_modelzoo.npz), conf 0.3 / iou 0.4 as
test.py:44-45.  Latency of (a) the eager path through the public mirror API, (b) the resident plan + fused decode/NMS
launched directly, (c) the same launches replayed as a CUDA graph; beside it the oracle port of the reference's CPU path on
this box's host cores.  Prints one JSON line."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import yfv2  # noqa: E402,F401
import synth  # noqa: E402
import yfv2_engine as eng  # noqa: E402
import model.detector as det  # noqa: E402
import utils.utils as uu  # noqa: E402
from oracle import net as onet, post as opost  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
w = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "modelzoo_weights.npz")).items()}
img_u8 = torch.from_numpy(np.load(os.path.join(G, "images_modelzoo.npz"))["000139_u8"])
cfg = synth.coco_cfg()
dev = torch.device("cuda", 0)
m = det.Detector(80, 3, True)
m.load_state_dict(w)
m = m.to(dev).eval()
x = (img_u8.to(dev).float() / 255.0)
CONF, IOU = 0.3, 0.4


def wall(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1e3 * ts[len(ts) // 2]


def eager():
    preds = m(x)
    return uu.detect(preds, cfg, CONF, IOU)


rows = eager()[0]
assert [(int(r[5]), "%.2f" % r[4]) for r in rows.numpy()] == [(0, "0.87"), (1, "0.46"), (0, "0.32")]
ms_eager = wall(eager, 50)

plan = m._plan_for(x)
preds = plan.alloc_preds()
out = torch.empty((1, eng.MAX_DET, 6), dtype=torch.float32, device=dev)
counts = torch.empty((1,), dtype=torch.int32, device=dev)
anchors = eng.anchors_array(cfg)
L = eng.lib()


def direct():
    plan.forward(x, preds)
    rc = L.yfv2_decode_nms(eng._ptr_array(preds), 1, 352, 352, 3, 80, anchors, ctypes.c_float(CONF), ctypes.c_double(IOU), None, 0,
                           eng.MAX_DET, ctypes.c_float(eng.MAX_WH), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(counts.data_ptr()),
                           None, None, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    assert rc == 0, L.yfv2_last_error()


ms_direct = wall(direct, 100)
graph_ms, graph_err = None, None
try:
    s = torch.cuda.Stream(dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        direct()
    torch.cuda.current_stream(dev).wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        direct()
    gr.replay()
    torch.cuda.synchronize()
    k = int(counts[0])
    assert torch.equal(out[0, :k].cpu(), rows), "graph replay differs from the eager result"
    graph_ms = wall(gr.replay, 200)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100):
        gr.replay()
    b.record()
    b.synchronize()
    graph_dev_ms = a.elapsed_time(b) / 100
except Exception as e:       # report, never hide
    graph_err, graph_dev_ms = repr(e), None

# the reference's CPU path (oracle port) on this box
sd = {k: v.clone() for k, v in w.items()}
xc = x.cpu()
best = None
for th in (1, 4, 8, 16, 32):
    if th > (os.cpu_count() or 1):
        break
    torch.set_num_threads(th)
    for _ in range(2):
        with torch.no_grad():
            p = onet.forward(sd, xc)
        opost.nms(opost.decode(p, cfg), CONF, IOU)
    ts = []
    for _ in range(7):
        t0 = time.perf_counter()
        with torch.no_grad():
            p = onet.forward(sd, xc)
        opost.nms(opost.decode(p, cfg), CONF, IOU)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    if best is None or ts[3] < best[0]:
        best = (ts[3], th)
print(json.dumps({"what": "configs[0]: one 352x352 image, forward + decode + NMS(0.3, 0.4), modelzoo weights, img/000139",
                  "detections": [(int(r[5]), round(float(r[4]), 2)) for r in rows.numpy()],
                  "eager_mirror_api_ms": round(ms_eager, 4), "direct_launches_ms": round(ms_direct, 4),
                  "cuda_graph_wall_ms": None if graph_ms is None else round(graph_ms, 4),
                  "cuda_graph_device_ms": None if graph_dev_ms is None else round(graph_dev_ms, 4), "cuda_graph_error": graph_err,
                  "launches": plan.forward_launches + 1,
                  "cpu_oracle_ms": round(1e3 * best[0], 3), "cpu_threads": best[1], "cpu_cores_visible": os.cpu_count(),
                  "reference_published": "22.8 ms forward on 8 vCPU (BASELINE.md 2, survey container)"}))
