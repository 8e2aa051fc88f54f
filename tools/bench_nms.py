#!/usr/bin/env python
"""BASELINE configs[4]: per-image NMS microbenchmark, 22x22x3 + 11x11x3 anchors (1815 candidates, 80 classes), 10 000
images, conf 0.001 / iou 0.4.  Candidate sets (seeded logits pushed through the decode):
  dense   reg ~ N(0,1), obj ~ 2N(0,1), cls ~ 2N(0,1)      99.6 % of the candidates pass, the 300 cap is hit on every image
  sparse  same with obj ~ N(-6,2)                          trained-model-like sparsity
  ties    every logit exactly 0: all 1815 scores equal   (the degenerate near-tie regime of randomly initialised heads)
Times yfv2_nms on the resident [N,1815,85] tensor (6.2 GB > L2) and yfv2_decode_nms straight from the logits (CUDA events,
3 warm-ups), and the oracle's NMS (C port of the reference loop + torchvision greedy NMS) on a bounded sample in chunks of
64 images on the host cores; checks rows bit-exact on that sample.  Prints one JSON line."""
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
from oracle import post as opost  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
SAMPLE = int(sys.argv[2]) if len(sys.argv) > 2 else 256
CONF, IOU = 0.001, 0.4
cfg = synth.coco_cfg()
dev = torch.device("cuda", 0)


def logits(kind, n, seed=4):
    g = torch.Generator(device=dev).manual_seed(seed)
    out = []
    for h in (22, 11):
        if kind == "ties":
            out += [torch.zeros((n, 12, h, h), device=dev), torch.zeros((n, 3, h, h), device=dev), torch.zeros((n, 80, h, h), device=dev)]
            continue
        reg = torch.randn((n, 12, h, h), generator=g, device=dev)
        obj = torch.randn((n, 3, h, h), generator=g, device=dev) * 2
        if kind == "sparse":
            obj = obj - 6.0
        cls = torch.randn((n, 80, h, h), generator=g, device=dev) * 2
        out += [reg, obj, cls]
    return out


def time_gpu(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


res = {"what": "configs[4] NMS microbench", "images": N, "candidates": 1815, "classes": 80, "conf": CONF, "iou": IOU, "sets": {}}
for kind in ("dense", "sparse", "ties"):
    preds = logits(kind, N)
    dets = torch.cat([eng.decode([p[i:i + 1000] for p in preds], cfg) for i in range(0, N, 1000)], 0)     # [N,1815,85]
    pass_rate = float((dets[:2000, :, 4:5] * dets[:2000, :, 5:]).amax(2).gt(CONF).float().mean())
    out, counts, _ = eng.nms(dets, CONF, IOU, want_idx=False)
    ms_nms = time_gpu(lambda: eng.nms(dets, CONF, IOU, want_idx=False))
    ms_fused = time_gpu(lambda: eng.decode_nms(preds, cfg, CONF, IOU))
    fo, fc, _ = eng.decode_nms(preds, cfg, CONF, IOU)
    assert torch.equal(fc, counts) and torch.equal(fo, out), "fused decode+NMS differs from decode -> NMS"
    # oracle on a bounded sample, chunks of 64 (the reference aborts its loop after 1 s, utils/utils.py:292-294)
    sample = dets[:SAMPLE].cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    opost.nms(sample[:64], CONF, IOU)
    t0 = time.perf_counter()
    want = []
    for i in range(0, SAMPLE, 64):
        want += opost.nms(sample[i:i + 64], CONF, IOU)
    cpu_s = time.perf_counter() - t0
    for i, w in enumerate(want):
        k = int(counts[i])
        assert k == w.shape[0] and np.array_equal(out[i, :k].cpu().numpy(), w.numpy()), "%s image %d differs from the oracle" % (kind, i)
    bytes_in = dets.numel() * 4
    res["sets"][kind] = {"pass_rate": round(pass_rate, 4), "kept_mean": float(counts.float().mean()),
                         "nms_ms": round(ms_nms, 3), "nms_img_per_s": round(N / (ms_nms / 1e3)), "nms_GBps": round(bytes_in / (ms_nms / 1e3) / 1e9, 1),
                         "fused_decode_nms_ms": round(ms_fused, 3), "fused_img_per_s": round(N / (ms_fused / 1e3)),
                         "oracle_img_per_s": round(SAMPLE / cpu_s, 1), "oracle_sample": "%d images in chunks of 64, C NMS, %d host threads visible" % (SAMPLE, os.cpu_count() or 1),
                         "bit_exact_on_sample": True}
    del dets, preds, out, counts
    torch.cuda.empty_cache()
print(json.dumps(res))
