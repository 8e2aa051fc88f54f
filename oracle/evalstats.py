"""TEST INFRASTRUCTURE (see oracle/__init__.py): CPU restatement of the reference's evaluation bookkeeping,
get_batch_statistics (utils/utils.py:184-230) and bbox_iou (utils/utils.py:76-108), in plain numpy loops.
Pinned to the real reference by tests/golden/eval_cases.npz (tests/golden/make_golden_eval.py)."""
import numpy as np


def bbox_iou_np(box, boxes):
    """IoU of one xyxy box against [n,4] boxes, +1 pixel convention, every operation rounded to fp32 like torch's."""
    f = np.float32
    ix1, iy1 = np.maximum(box[0], boxes[:, 0]), np.maximum(box[1], boxes[:, 1])
    ix2, iy2 = np.minimum(box[2], boxes[:, 2]), np.minimum(box[3], boxes[:, 3])
    iw = np.maximum((ix2 - ix1).astype(f) + f(1), f(0)).astype(f)
    ih = np.maximum((iy2 - iy1).astype(f) + f(1), f(0)).astype(f)
    inter = (iw * ih).astype(f)
    a1 = f((f(box[2] - box[0]) + f(1)) * (f(box[3] - box[1]) + f(1)))
    a2 = (((boxes[:, 2] - boxes[:, 0]).astype(f) + f(1)) * ((boxes[:, 3] - boxes[:, 1]).astype(f) + f(1))).astype(f)
    den = (((a1 + a2).astype(f) - inter).astype(f) + f(1e-16)).astype(f)
    return (inter / den).astype(f)


def get_batch_statistics(outputs, targets, iou_threshold):
    """outputs: list of [n_i,6] float32 arrays (x1,y1,x2,y2,conf,cls) by descending conf; targets [nt,6] (img,cls,xyxy).
    Returns per image the float64 true-positive flags."""
    res = []
    for i, out in enumerate(outputs):
        tp = np.zeros(out.shape[0])
        ann = targets[targets[:, 0] == i][:, 1:]
        if len(ann):
            claimed = []
            for pi in range(out.shape[0]):
                if len(claimed) == len(ann):                       # every annotation matched: stop (utils.py:214-215)
                    break
                if out[pi, 5] not in ann[:, 0]:                    # label not among the image's targets (:218-219)
                    continue
                iou = bbox_iou_np(out[pi, :4].astype(np.float32), ann[:, 1:].astype(np.float32))
                bi = int(np.argmax(iou))                           # ALL annotations compete, first maximum (:221)
                if iou[bi] >= np.float32(iou_threshold) and bi not in claimed:
                    tp[pi] = 1
                    claimed.append(bi)
        res.append(tp)
    return res
