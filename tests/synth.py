"""Deterministic synthetic inputs shared by the golden generator and the tests.

Everything is drawn from numpy's legacy RandomState (bit-stable across numpy versions), so the
golden files only need to store the REFERENCE OUTPUTS; the inputs are regenerated here.
"""
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
COCO_ANCHORS = [12.64, 19.39, 37.88, 51.48, 55.71, 138.31, 126.91, 78.23, 131.57, 214.55, 279.92, 258.87]


def coco_cfg(width=352, height=352, classes=80):
    """The dict utils.utils.load_datafile returns for data/coco.data (hot-path keys only)."""
    return {"model_name": "coco", "classes": classes, "width": width, "height": height,
            "anchor_num": 3, "anchors": list(COCO_ANCHORS)}


def key_table():
    with open(os.path.join(HERE, "golden", "statedict_keys.json")) as f:
        return json.load(f)


def make_state_dict(seed, classes=80, anchor_num=3):
    """Random but well-conditioned weights for all 444 reference keys (non-trivial BN stats)."""
    rs = np.random.RandomState(seed)
    sd = {}
    for name, shape, dtype in key_table():
        shape = list(shape)
        if name.startswith("output_") and shape:
            if "reg" in name: shape[0] = 4 * anchor_num
            elif "obj" in name: shape[0] = anchor_num
            else: shape[0] = classes
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros((), dtype=torch.int64)
        elif name.endswith("running_var"):
            sd[name] = torch.from_numpy(rs.uniform(0.5, 1.5, shape).astype(np.float32))
        elif name.endswith("running_mean"):
            sd[name] = torch.from_numpy((0.2 * rs.randn(*shape)).astype(np.float32))
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = torch.from_numpy((rs.randn(*shape) * np.sqrt(1.0 / fan_in)).astype(np.float32))
        elif name.endswith(".weight"):      # BN gamma
            sd[name] = torch.from_numpy(rs.uniform(0.5, 1.2, shape).astype(np.float32))
        else:                               # BN beta / conv bias
            sd[name] = torch.from_numpy((0.2 * rs.randn(*shape)).astype(np.float32))
    return sd


def make_images(seed, n, h, w):
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.rand(n, 3, h, w).astype(np.float32))


def make_head_logits(seed, n, h, w, classes=80, anchor_num=3, obj_mean=0.0, obj_std=2.0):
    """Six head tensors with SURVEY 8(d) config[4] statistics (dense set by default)."""
    rs = np.random.RandomState(seed)
    out = []
    for s in (16, 32):
        hh, ww = h // s, w // s
        out.append(torch.from_numpy(rs.randn(n, 4 * anchor_num, hh, ww).astype(np.float32)))
        out.append(torch.from_numpy((obj_mean + obj_std * rs.randn(n, anchor_num, hh, ww)).astype(np.float32)))
        out.append(torch.from_numpy((2.0 * rs.randn(n, classes, hh, ww)).astype(np.float32)))
    return tuple(out)


def make_targets(seed, n, max_boxes=13, classes=80):
    """Rows (img_idx, cls, cx, cy, w, h), normalised; SURVEY 8(d) config[2] distribution."""
    rs = np.random.RandomState(seed)
    rows = []
    for i in range(n):
        for _ in range(rs.randint(1, max_boxes + 1)):
            rows.append([i, rs.randint(0, classes), rs.rand(), rs.rand(),
                         rs.uniform(0.02, 0.52), rs.uniform(0.02, 0.52)])
    return torch.tensor(rows, dtype=torch.float32).reshape(-1, 6)


def make_dets(seed, n, m, classes=80, side=352.0, quantize=False, dup=False, obj_pow=3.0):
    """Decoded-candidate tensors [n, m, 5+classes] built ONLY from RandomState draws and fp32
    multiplies (no transcendental, no reductions), so they are bit-reproducible anywhere and
    the reference NMS output on them can be pinned bit-exactly."""
    rs = np.random.RandomState(seed)
    d = np.empty((n, m, 5 + classes), np.float32)
    d[..., 0:2] = (rs.rand(n, m, 2) * side).astype(np.float32)
    d[..., 2:4] = (rs.rand(n, m, 2) * (side / 3) + 2).astype(np.float32)
    o = rs.rand(n, m).astype(np.float32)
    d[..., 4] = o * o * o if obj_pow == 3.0 else o
    c = rs.rand(n, m, classes).astype(np.float32)
    c = c * c; c = c * c; c = c * c                      # ^8: a few classes dominate
    d[..., 5:] = c
    if quantize:                                         # many exact score ties
        d[..., 4] = np.round(d[..., 4] * 16) / 16
        d[..., 5:] = np.round(d[..., 5:] * 8) / 8
    if dup:                                              # exact duplicate candidates
        d[:, m // 2:] = d[:, : m - m // 2]
    return torch.from_numpy(d)


NMS_CASES = {
    "dense":    dict(seed=51, n=3, m=1815),
    "ties":     dict(seed=52, n=2, m=1815, quantize=True),
    "dups":     dict(seed=53, n=2, m=1200, dup=True),
    "big640":   dict(seed=54, n=1, m=6000, side=640.0),
    "tiny":     dict(seed=55, n=4, m=7),
    "lowobj":   dict(seed=56, n=2, m=1815, obj_pow=3.0, classes=80, side=352.0),
    "c20":      dict(seed=57, n=2, m=1500, classes=20),
}


def make_eval_case(seed, n_img, max_det=40, max_gt=9, classes=6, size=352):
    """Seeded NMS-like outputs and pixel-xyxy targets for the evaluation bookkeeping tests: per image a list of boxes by
    descending confidence (some are jittered copies of ground-truth boxes, some have a wrong label, some are noise), and
    targets rows (image, class, x1, y1, x2, y2).  Returns (list of [n_i,6] float32 arrays, targets [nt,6] float32)."""
    rs = np.random.RandomState(seed)
    outs, tg = [], []
    for i in range(n_img):
        ng = rs.randint(0, max_gt + 1)
        gt = np.zeros((ng, 6), np.float32)
        gt[:, 0] = i
        gt[:, 1] = rs.randint(0, classes, ng)
        c = rs.rand(ng, 2) * size
        wh = 8 + rs.rand(ng, 2) * size * 0.4
        gt[:, 2:4], gt[:, 4:6] = c - wh / 2, c + wh / 2
        tg.append(gt)
        nd = rs.randint(0, max_det + 1)
        d = np.zeros((nd, 6), np.float32)
        for j in range(nd):
            kind = rs.rand()
            if ng and kind < 0.6:                      # a detection of some ground-truth box, jittered
                g = gt[rs.randint(ng)]
                d[j, :4] = g[2:6] + rs.randn(4) * (3 if kind < 0.4 else 40)
                d[j, 5] = g[1] if rs.rand() < 0.8 else rs.randint(0, classes)
            else:
                c2 = rs.rand(2) * size
                wh2 = 8 + rs.rand(2) * size * 0.4
                d[j, :2], d[j, 2:4] = c2 - wh2 / 2, c2 + wh2 / 2
                d[j, 5] = rs.randint(0, classes + 2)
        d[:, 4] = np.sort(rs.rand(nd).astype(np.float32))[::-1]
        outs.append(d)
    return outs, np.concatenate(tg, 0) if tg else np.zeros((0, 6), np.float32)
