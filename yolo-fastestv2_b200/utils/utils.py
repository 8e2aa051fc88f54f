"""Drop-in replacement for the hot-path functions of the reference's utils/utils.py:
load_datafile (:13-65), handel_preds (:303-358), non_max_suppression (:232-296).

handel_preds / non_max_suppression keep the reference's return types (a CPU [N,M,5+C] tensor; a list
of CPU [n_i,6] tensors) but do all arithmetic in libyfv2.so kernels; one device->host copy per batch
replaces the reference's per-image Python loops.  detect() is the fused fast path.
"""
import os
import sys

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import yfv2_engine  # noqa: E402

_LIST_KEYS = ("anchors", "steps")
_STR_KEYS = ("model_name", "val", "train", "names", "pre_weights")
_INT_KEYS = ("epochs", "batch_size", "classes", "width", "height", "anchor_num", "subdivisions")
_FLOAT_KEYS = ("learning_rate",)


def load_datafile(data_path):
    """Parses the reference's `.data` format: `key=value` lines, blank lines and `[section]` lines skipped,
    values typed by key (reference utils/utils.py:38-42); unknown keys are reported and ignored."""
    assert os.path.exists(data_path), "data file not found: %s" % data_path
    cfg = {k: None for k in ("model_name", "epochs", "steps", "batch_size", "subdivisions", "learning_rate", "pre_weights",
                             "classes", "width", "height", "anchor_num", "anchors", "val", "train", "names")}
    with open(data_path, "r") as f:
        for line in f.readlines():
            if line == "\n" or line[0] == "[":
                continue
            data = line.strip().split("=")
            key = data[0]
            if key not in cfg:
                print("%s: unknown configuration item %s" % (data_path, data))
            elif key in _INT_KEYS:
                cfg[key] = int(data[1])
            elif key in _STR_KEYS:
                cfg[key] = data[1]
            elif key in _FLOAT_KEYS:
                cfg[key] = float(data[1])
            elif key in _LIST_KEYS:
                cfg[key] = [float(x) for x in data[1].split(",")]
    return cfg


def handel_preds(preds, cfg, device):
    """[N, sum(h*w*A), 5+C] fp32 CPU tensor, as the reference returns (utils/utils.py:328,350-358)."""
    return yfv2_engine.decode(preds, cfg).cpu()


def _to_list(out, counts):
    out, counts = out.cpu(), counts.cpu().tolist()
    return [out[i, :c].clone() for i, c in enumerate(counts)]


def non_max_suppression(prediction, conf_thres=0.3, iou_thres=0.45, classes=None):
    """list (len N) of CPU fp32 [n_i,6] tensors (x1,y1,x2,y2,conf,cls), descending conf, at most 300 each."""
    dev = prediction.device if prediction.is_cuda else torch.device("cuda", torch.cuda.current_device())
    if not torch.cuda.is_available():
        raise RuntimeError("yfv2 non_max_suppression needs a CUDA device (no CPU fallback)")
    out, counts, _ = yfv2_engine.nms(prediction.to(dev), conf_thres, iou_thres, classes, want_idx=False)
    return _to_list(out, counts)


def detect(preds, cfg, conf_thres=0.3, iou_thres=0.45, classes=None):
    """Fused handel_preds + non_max_suppression: same list-of-[n_i,6] result without the candidate tensor."""
    out, counts, _ = yfv2_engine.decode_nms(preds, cfg, conf_thres, iou_thres, classes)
    return _to_list(out, counts)


# ---- evaluation bookkeeping (reference utils/utils.py:67-230, 361-395) -----------------------------------------------------
# Same functions, signatures and return values as the reference.  The per-box Python loop of get_batch_statistics runs as
# one kernel (yfv2_batch_statistics: same greedy order, same fp32 IoU, bit-identical flags); evaluation() keeps detections on
# the device from the forward to the true-positive flags (one device->host copy per batch).
import numpy as np  # noqa: E402


def xywh2xyxy(x):
    """[cx, cy, w, h] rows -> [x1, y1, x2, y2] (utils/utils.py:67-74); tensor or ndarray."""
    y = torch.zeros_like(x) if isinstance(x, torch.Tensor) else np.zeros_like(x)
    half_w, half_h = x[:, 2] / 2, x[:, 3] / 2
    y[:, 0], y[:, 1] = x[:, 0] - half_w, x[:, 1] - half_h
    y[:, 2], y[:, 3] = x[:, 0] + half_w, x[:, 1] + half_h
    return y


def bbox_iou(box1, box2, x1y1x2y2=True):
    """IoU with the reference's +1 pixel convention (utils/utils.py:76-108); row-wise / broadcast over [n,4] tensors."""
    if not x1y1x2y2:
        box1, box2 = xywh2xyxy(box1), xywh2xyxy(box2)
    ax1, ay1, ax2, ay2 = box1[:, 0], box1[:, 1], box1[:, 2], box1[:, 3]
    bx1, by1, bx2, by2 = box2[:, 0], box2[:, 1], box2[:, 2], box2[:, 3]
    iw = torch.clamp(torch.min(ax2, bx2) - torch.max(ax1, bx1) + 1, min=0)
    ih = torch.clamp(torch.min(ay2, by2) - torch.max(ay1, by1) + 1, min=0)
    inter = iw * ih
    area1 = (ax2 - ax1 + 1) * (ay2 - ay1 + 1)
    area2 = (bx2 - bx1 + 1) * (by2 - by1 + 1)
    return inter / (area1 + area2 - inter + 1e-16)


def compute_ap(recall, precision):
    """Area under the precision envelope at the points where recall changes (utils/utils.py:110-137)."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([0.0], precision, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]                 # running maximum from the right = the envelope
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def ap_per_class(tp, conf, pred_cls, target_cls):
    """(mean precision, mean recall, mean AP, mean F1) over the classes present in the targets (utils/utils.py:139-182)."""
    tp, conf, pred_cls, target_cls = np.asarray(tp), np.asarray(conf), np.asarray(pred_cls), np.asarray(target_cls)
    order = np.argsort(-conf)
    tp, pred_cls = tp[order], pred_cls[order]
    ap, p, r = [], [], []
    for c in np.unique(target_cls):
        sel = pred_cls == c
        n_gt, n_p = (target_cls == c).sum(), sel.sum()
        if n_p == 0 and n_gt == 0:
            continue
        if n_p == 0 or n_gt == 0:
            ap.append(0); r.append(0); p.append(0)
            continue
        tpc = tp[sel].cumsum()
        fpc = (1 - tp[sel]).cumsum()
        recall_curve = tpc / (n_gt + 1e-16)
        precision_curve = tpc / (tpc + fpc)
        r.append(recall_curve[-1]); p.append(precision_curve[-1])
        ap.append(compute_ap(recall_curve, precision_curve))
    p, r, ap = np.array(p), np.array(r), np.array(ap)
    f1 = 2 * p * r / (p + r + 1e-16)
    return np.mean(p), np.mean(r), np.mean(ap), np.mean(f1)


def _batch_statistics_device(out, counts, targets, iou_threshold):
    tp = yfv2_engine.batch_statistics(out, counts, targets, iou_threshold).cpu().numpy().astype(np.float64)
    out_c, cnt = out.cpu(), counts.cpu().tolist()
    return [[tp[i, :c], out_c[i, :c, 4], out_c[i, :c, -1]] for i, c in enumerate(cnt)]


def get_batch_statistics(outputs, targets, iou_threshold, device):
    """Per sample [true_positives (float64 ndarray), pred_scores, pred_labels] (utils/utils.py:184-230).
    outputs: the list non_max_suppression returns; targets: [nt,6] rows (image, class, x1, y1, x2, y2) in pixels."""
    if not torch.cuda.is_available():
        raise RuntimeError("yfv2 get_batch_statistics needs a CUDA device (no CPU fallback)")
    dev = targets.device if targets.is_cuda else torch.device("cuda", torch.cuda.current_device())
    keep = [i for i, o in enumerate(outputs) if o is not None]
    n, max_det = len(outputs), max([o.shape[0] for o in outputs if o is not None] + [1])
    out = torch.zeros((n, max_det, 6), dtype=torch.float32)
    counts = torch.zeros((n,), dtype=torch.int32)
    for i in keep:
        out[i, :outputs[i].shape[0]] = outputs[i]
        counts[i] = outputs[i].shape[0]
    stats = _batch_statistics_device(out.to(dev), counts.to(dev), targets, iou_threshold)
    return [stats[i] for i in keep]


def evaluation(val_dataloader, cfg, model, device, conf_thres=0.01, nms_thresh=0.4, iou_thres=0.5):
    """(precision, recall, AP, F1) of `model` over `val_dataloader` (utils/utils.py:361-395): AP pass at conf 0.01, P/R pass
    at conf 0.3, NMS IoU 0.4, match IoU 0.5.  Like the reference it rescales `targets` in place to pixel xyxy."""
    try:
        from tqdm import tqdm
    except ImportError:                                            # progress bar only
        def tqdm(it):
            return it
    labels, sample_metrics = [], []
    for imgs, targets in tqdm(val_dataloader):
        imgs = imgs.to(device)                                     # uint8 goes straight in: the /255 is fused into the stem's load
        if imgs.dtype != torch.uint8:
            imgs = imgs.float() / 255.0
        targets = targets.to(device)
        labels += targets[:, 1].tolist()
        targets[:, 2:] = xywh2xyxy(targets[:, 2:])
        targets[:, 2:] *= torch.tensor([cfg["width"], cfg["height"], cfg["width"], cfg["height"]]).to(device)
        with torch.no_grad():
            preds = model(imgs)
            out, counts, _ = yfv2_engine.decode_nms(preds, cfg, conf_thres, nms_thresh)
        sample_metrics += _batch_statistics_device(out, counts, targets, iou_thres)
    if len(sample_metrics) == 0:
        print("---- No detections over whole validation set ----")
        return None
    true_positives, pred_scores, pred_labels = [np.concatenate(x, 0) for x in list(zip(*sample_metrics))]
    return ap_per_class(true_positives, pred_scores, pred_labels, labels)


# ---- anything else the reference's utils/utils.py defines is taken from the reference checkout when one is on sys.path --------
def _overlay_reference():
    import importlib.util
    import warnings
    here = os.path.dirname(os.path.abspath(__file__))
    for p in list(sys.path):
        cand = os.path.join(os.path.abspath(p or "."), "utils", "utils.py")
        if os.path.isfile(cand) and os.path.dirname(cand) != here:
            try:
                spec = importlib.util.spec_from_file_location("_yfv2_reference_utils", cand)
                ref = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(ref)
            except Exception as e:     # say so: a half-importable reference must not silently change what `utils.utils` offers
                warnings.warn("yfv2: could not import the reference's %s (%s: %s); only the mirror's own functions are available"
                              % (cand, type(e).__name__, e))
                return None
            for name in dir(ref):
                if not name.startswith("_") and name not in globals():
                    globals()[name] = getattr(ref, name)
            return ref
    return None


_reference_utils = _overlay_reference()
