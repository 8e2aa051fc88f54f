"""Generates tests/golden/* by running the REAL reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py [--reference /root/reference]

Shims: a `torchsummary` stub (imported at module import time by
model/backbone/shufflenetv2.py:3, not installed) and an int() cast for the float-tensor
clamp bound at utils/loss.py:119 (torch >= 1.10 rejects it; the pinned torch 1.9 accepted it).
Inputs come from tests/synth.py (seeded numpy), so only reference OUTPUTS are stored.
"""
import argparse
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def install_shims():
    ts = types.ModuleType("torchsummary")
    ts.summary = lambda *a, **k: None
    sys.modules["torchsummary"] = ts
    orig = torch.Tensor.clamp_

    def clamp_(self, min=None, max=None):
        if isinstance(max, torch.Tensor) and not self.is_floating_point():
            max = int(max)
        if isinstance(min, torch.Tensor) and not self.is_floating_point():
            min = int(min)
        return orig(self, min, max)
    torch.Tensor.clamp_ = clamp_


def np_(t):
    return t.detach().cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    install_shims()
    sys.path.insert(0, args.reference)
    os.chdir(args.reference)
    torch.manual_seed(0)
    import model.detector as rdet
    import utils.utils as rutils
    import utils.loss as rloss
    import torchvision
    import synth

    dev = torch.device("cpu")
    meta = {"torch": torch.__version__, "torchvision": torchvision.__version__, "numpy": np.__version__}

    # 1. state_dict key table -------------------------------------------------------------
    m = rdet.Detector(80, 3, True)
    keys = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    with open(os.path.join(HERE, "statedict_keys.json"), "w") as f:
        json.dump(keys, f)

    def ref_model(sd, classes=80):
        mm = rdet.Detector(classes, 3, True)
        mm.load_state_dict(sd, strict=True)
        return mm

    def run_post(preds, cfg, thr_list):
        dets = rutils.handel_preds(preds, cfg, dev)
        out = {"decode": np_(dets)}
        for (ct, it) in thr_list:
            # the reference aborts after 1 s of wall clock: feed one image at a time
            rows = [rutils.non_max_suppression(dets[i:i + 1].clone(), ct, it)[0] for i in range(dets.shape[0])]
            out["nms_%g_%g_counts" % (ct, it)] = np.array([r.shape[0] for r in rows], np.int64)
            out["nms_%g_%g_rows" % (ct, it)] = np_(torch.cat(rows, 0)) if rows else np.zeros((0, 6), np.float32)
        return out

    THR = [(0.3, 0.4), (0.01, 0.4), (0.001, 0.4)]

    # 2. small net: every stage tapped, eval + train mode -------------------------------------
    sd = synth.make_state_dict(11)
    x = synth.make_images(12, 2, 64, 96)
    mm = ref_model(sd).eval()
    taps = {}
    hooks = []
    def tap(name, mod):
        hooks.append(mod.register_forward_hook(lambda _m, _i, o, n=name: taps.__setitem__(n, np_(o))))
    tap("stem", mm.backbone.maxpool)
    for s in (2, 3, 4):
        for bi, blk in enumerate(getattr(mm.backbone, "stage%d" % s)):
            tap("stage%d.%d" % (s, bi), blk)
    tap("S2", mm.fpn.conv1x1_2); tap("S3", mm.fpn.conv1x1_3)
    for n in ("cls_head_2", "reg_head_2", "cls_head_3", "reg_head_3"):
        tap(n, getattr(mm.fpn, n))
    with torch.no_grad():
        preds = mm(x)
    for h_ in hooks: h_.remove()
    cfg = synth.coco_cfg(96, 64)
    out = {"pred%d" % i: np_(p) for i, p in enumerate(preds)}
    out.update({"tap_" + k: v for k, v in taps.items()})
    out.update(run_post(preds, cfg, THR))
    mt = ref_model(sd).train()
    pt = mt(x)
    out.update({"train_pred%d" % i: np_(p) for i, p in enumerate(pt)})
    sdt = mt.state_dict()
    for k in ("backbone.first_conv.1", "backbone.stage3.2.branch_main.4", "fpn.cls_head_2.block.9"):
        out["train_rm_" + k] = np_(sdt[k + ".running_mean"]); out["train_rv_" + k] = np_(sdt[k + ".running_var"])
    mex = rdet.Detector(80, 3, True, True); mex.load_state_dict(sd); mex.eval()
    with torch.no_grad():
        e2, e3 = mex(x)
    out["export_2"], out["export_3"] = np_(e2), np_(e3)
    np.savez_compressed(os.path.join(HERE, "net_small.npz"), **out)

    # 3. one 352x352 image, random weights ------------------------------------------------------
    sd = synth.make_state_dict(21)
    x = synth.make_images(22, 1, 352, 352)
    mm = ref_model(sd).eval()
    with torch.no_grad():
        preds = mm(x)
    out = {"pred%d" % i: np_(p) for i, p in enumerate(preds)}
    post = run_post(preds, synth.coco_cfg(), THR); post.pop("decode")
    out.update(post)
    np.savez_compressed(os.path.join(HERE, "net_352.npz"), **out)

    # 4. trained modelzoo weights on two bundled images (known answers: img/*_result.png) --------
    import cv2
    wsd = torch.load("modelzoo/coco2017-0.241078ap-model.pth", map_location="cpu")
    np.savez_compressed(os.path.join(HERE, "modelzoo_weights.npz"), **{k: np_(v) for k, v in wsd.items()})
    mm = ref_model(wsd).eval()
    cfg = rutils.load_datafile("data/coco.data")
    out = {}
    for name in ("000139", "000004"):
        ori = cv2.imread("img/%s.jpg" % name)
        res = cv2.resize(ori, (cfg["width"], cfg["height"]), interpolation=cv2.INTER_LINEAR)     # test.py:34-38
        img_u8 = torch.from_numpy(res.reshape(1, cfg["height"], cfg["width"], 3).transpose(0, 3, 1, 2).copy())
        img = img_u8.float() / 255.0
        with torch.no_grad():
            preds = mm(img)
        out[name + "_u8"] = np_(img_u8)
        for i, p in enumerate(preds):
            out["%s_pred%d" % (name, i)] = np_(p)
        post = run_post(preds, cfg, [(0.3, 0.4), (0.001, 0.4)]); post.pop("decode")
        out.update({name + "_" + k: v for k, v in post.items()})
    np.savez_compressed(os.path.join(HERE, "images_modelzoo.npz"), **out)

    # 5. decode + NMS on synthetic head logits (dense / sparse / 640) ---------------------------
    out = {}
    for tag, kw, hw in (("dense", dict(), (352, 352)), ("sparse", dict(obj_mean=-6.0), (352, 352)),
                        ("dense640", dict(), (640, 640))):
        n = 3 if hw[0] == 352 else 1
        preds = synth.make_head_logits(31, n, hw[0], hw[1], **kw)
        post = run_post(preds, synth.coco_cfg(hw[1], hw[0]), THR)
        dec = post.pop("decode")
        out[tag + "_decode_sample"] = dec[:, ::37].copy()          # every 37th row: pins decode at full size
        out.update({tag + "_" + k: v for k, v in post.items()})
    # class whitelist branch (utils.py:271-272)
    preds = synth.make_head_logits(31, 1, 352, 352)
    dets = rutils.handel_preds(preds, synth.coco_cfg(), dev)
    r = rutils.non_max_suppression(dets.clone(), 0.01, 0.4, classes=[0, 5, 17])[0]
    out["filter_rows"] = np_(r)
    np.savez_compressed(os.path.join(HERE, "post_cases.npz"), **out)

    # 5b. NMS alone on bit-reproducible synthetic candidates (tests/synth.py:make_dets) -> exact pin
    out = {}
    for tag, kw in synth.NMS_CASES.items():
        dets = synth.make_dets(**kw)
        for (ct, it) in THR + [(0.25, 0.45)]:
            rows = [rutils.non_max_suppression(dets[i:i + 1].clone(), ct, it)[0] for i in range(dets.shape[0])]
            out["%s_%g_%g_counts" % (tag, ct, it)] = np.array([r.shape[0] for r in rows], np.int64)
            out["%s_%g_%g_rows" % (tag, ct, it)] = np_(torch.cat(rows, 0))
    dets = synth.make_dets(**synth.NMS_CASES["dense"])
    r = rutils.non_max_suppression(dets[:1].clone(), 0.01, 0.4, classes=[0, 5, 17])[0]
    out["filter_rows"] = np_(r)
    np.savez_compressed(os.path.join(HERE, "nms_cases.npz"), **out)

    # 6. loss: build_target tuples, the 4 scalars and d(loss)/d(preds) --------------------------
    out = {}
    for tag, n, tseed in (("a", 2, 41), ("b", 3, 43)):
        preds = [p.clone().requires_grad_(True) for p in synth.make_head_logits(40 + n, n, 352, 352, obj_std=1.0)]
        targets = synth.make_targets(tseed, n)
        cfg = synth.coco_cfg()
        tcls, tbox, indices, anch = rloss.build_target(preds, targets.clone(), cfg, dev)
        for L in range(2):
            out["%s_tcls%d" % (tag, L)] = np_(tcls[L]); out["%s_tbox%d" % (tag, L)] = np_(tbox[L])
            out["%s_anch%d" % (tag, L)] = np_(anch[L])
            out["%s_idx%d" % (tag, L)] = np.stack([np_(t) for t in indices[L]], 0)
        lb, lo, lc, loss = rloss.compute_loss(preds, targets.clone(), cfg, dev)
        loss.backward()
        out[tag + "_losses"] = np.array([lb.item(), lo.item(), lc.item(), loss.item()], np.float64)
        for i, p in enumerate(preds):
            out["%s_grad%d" % (tag, i)] = np_(p.grad)
    preds = [p.clone().requires_grad_(True) for p in synth.make_head_logits(45, 2, 352, 352)]
    lb, lo, lc, loss = rloss.compute_loss(preds, torch.zeros(0, 6), synth.coco_cfg(), dev)   # nt == 0
    out["empty_losses"] = np.array([lb.item(), lo.item(), lc.item(), loss.item()], np.float64)
    np.savez_compressed(os.path.join(HERE, "loss_cases.npz"), **out)

    with open(os.path.join(HERE, "META.json"), "w") as f:
        json.dump(meta, f)
    print("golden written to", HERE)


if __name__ == "__main__":
    main()
