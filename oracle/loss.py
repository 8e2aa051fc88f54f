"""CPU restatement of the reference training loss (utils/loss.py).

build_target()  <- utils/loss.py:53-124
ciou()          <- utils/loss.py:8-51 with x1y1x2y2=False, CIoU=True (the only mode compute_loss uses)
compute_loss()  <- utils/loss.py:130-208

dtype flow kept as in the reference: anchors are float64 (np.array of python floats,
loss.py:59-60), so the anchor-ratio test, the predicted wh and the whole CIoU run in
float64, while gather/sigmoid/BCE/CE run in float32.  The integer clamp bound that the
reference passes as a float tensor (loss.py:119; accepted by its pinned torch 1.9, rejected
by torch >= 1.10) is taken as the integer h-1 / w-1.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

OFFSETS = ((0.0, 0.0), (0.5, 0.0), (0.0, 0.5), (-0.5, 0.0), (0.0, -0.5))   # loss.py:67-71 times g=0.5
BALANCE = (1.0, 0.4)                                                        # loss.py:131


def build_target(preds, targets, cfg):
    """Returns per level: tcls[m] int64, tbox[m,4] fp32, (b,a,gj,gi) int64, anch[m,2] fp64.
    Row order: offset-type major, then anchor, then target (loss.py:93-105)."""
    A, nt = cfg["anchor_num"], targets.shape[0]
    levels = len(preds) // 3
    anchors = torch.from_numpy(np.array(cfg["anchors"], dtype=np.float64).reshape(levels, A, 2))
    at = torch.arange(A).float().view(A, 1).repeat(1, nt)
    t7 = torch.cat((targets.float().repeat(A, 1, 1), at[:, :, None]), 2)     # [A,nt,7]
    off = torch.tensor(OFFSETS, dtype=torch.float32)
    tcls, tbox, indices, anch = [], [], [], []
    for L in range(levels):
        _, _, h, w = preds[3 * L].shape
        assert cfg["width"] / w == cfg["height"] / h                           # loss.py:78
        stride = cfg["width"] / w
        a_cfg = anchors[L] / stride                                            # fp64 [A,2]
        gain = torch.tensor([1, 1, w, h, w, h, 1], dtype=torch.float32)
        gt = t7 * gain
        if nt:
            r = gt[:, :, 4:6] / a_cfg[:, None]                                 # fp64
            sel = torch.max(r, 1.0 / r).max(2)[0] < 2                          # loss.py:94
            t = gt[sel]
            gxy = t[:, 2:4]
            gxi = gain[[2, 3]] - gxy
            j, k = ((gxy % 1.0 < 0.5) & (gxy > 1.0)).T
            l, m = ((gxi % 1.0 < 0.5) & (gxi > 1.0)).T
            mask = torch.stack((torch.ones_like(j), j, k, l, m))               # [5,m]
            t = t.repeat((5, 1, 1))[mask]
            offsets = (torch.zeros_like(gxy)[None] + off[:, None])[mask]
        else:
            t = t7[0]
            offsets = 0
        b, c = t[:, :2].long().T
        gxy, gwh = t[:, 2:4], t[:, 4:6]
        gij = (gxy - offsets).long()
        gij[:, 0].clamp_(0, w - 1)                                             # loss.py:119 (in place, before tbox)
        gij[:, 1].clamp_(0, h - 1)
        a = t[:, 6].long()
        indices.append((b, a, gij[:, 1].clone(), gij[:, 0].clone()))
        tbox.append(torch.cat((gxy - gij, gwh), 1))
        anch.append(a_cfg[a])
        tcls.append(c)
    return tcls, tbox, indices, anch


def ciou(pbox, tbox):
    """bbox_iou(pbox.t(), tbox, x1y1x2y2=False, CIoU=True) (loss.py:8-51).  pbox,tbox [m,4] xywh."""
    b1, b2 = pbox.t(), tbox.t()
    b1_x1, b1_x2 = b1[0] - b1[2] / 2, b1[0] + b1[2] / 2
    b1_y1, b1_y2 = b1[1] - b1[3] / 2, b1[1] + b1[3] / 2
    b2_x1, b2_x2 = b2[0] - b2[2] / 2, b2[0] + b2[2] / 2
    b2_y1, b2_y2 = b2[1] - b2[3] / 2, b2[1] + b2[3] / 2
    inter = (torch.min(b1_x2, b2_x2) - torch.max(b1_x1, b2_x1)).clamp(0) * \
            (torch.min(b1_y2, b2_y2) - torch.max(b1_y1, b2_y1)).clamp(0)
    w1, h1 = b1_x2 - b1_x1, b1_y2 - b1_y1
    w2, h2 = b2_x2 - b2_x1, b2_y2 - b2_y1
    union = (w1 * h1 + 1e-16) + w2 * h2 - inter
    iou = inter / union
    cw = torch.max(b1_x2, b2_x2) - torch.min(b1_x1, b2_x1)
    ch = torch.max(b1_y2, b2_y2) - torch.min(b1_y1, b2_y1)
    c2 = cw ** 2 + ch ** 2 + 1e-16
    rho2 = ((b2_x1 + b2_x2) - (b1_x1 + b1_x2)) ** 2 / 4 + ((b2_y1 + b2_y2) - (b1_y1 + b1_y2)) ** 2 / 4
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)
    with torch.no_grad():
        alpha = v / (1 - iou + v)
    return iou - (rho2 / c2 + v * alpha)


def compute_loss(preds, targets, cfg):
    """Returns (lbox, lobj, lcls, loss), each shape [1] fp32, differentiable w.r.t. preds."""
    A, C = cfg["anchor_num"], cfg["classes"]
    lcls, lbox, lobj = torch.zeros(1), torch.zeros(1), torch.zeros(1)
    tcls, tbox, indices, anchors = build_target(preds, targets, cfg)
    for L in range(len(preds) // 3):
        reg, obj, cls = preds[3 * L], preds[3 * L + 1], preds[3 * L + 2]
        N, _, h, w = reg.shape
        b, a, gj, gi = indices[L]
        nb = b.shape[0]
        reg5 = reg.reshape(N, A, -1, h, w).permute(0, 1, 3, 4, 2)             # [N,A,h,w,4]
        obj4 = obj.reshape(N, A, -1, h, w).permute(0, 1, 3, 4, 2)[..., 0]     # [N,A,h,w]
        tobj = torch.zeros_like(obj4)
        if nb:
            ps = reg5[b, a, gj, gi]
            pxy = ps[:, :2].sigmoid() * 2.0 - 0.5
            pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * anchors[L]                # fp64
            pbox = torch.cat((pxy, pwh), 1)
            lbox = lbox + (1.0 - ciou(pbox, tbox[L])).mean().float()          # loss.py:163 (+= into fp32)
            tobj[b, a, gj, gi] = 1.0                                          # loss.py:177
        lobj = lobj + F.binary_cross_entropy_with_logits(obj4, tobj) * BALANCE[L]
        if nb and C > 1:
            pc = cls.permute(0, 2, 3, 1)[b, gj, gi]                           # [m,C]
            lcls = lcls + F.cross_entropy(pc, tcls[L]) / C                    # loss.py:198
    lbox = lbox * 3.2
    lobj = lobj * 64
    lcls = lcls * 32
    return lbox, lobj, lcls, lbox + lobj + lcls
