"""Generates tests/golden/eval_cases.npz by running the REAL reference's get_batch_statistics / ap_per_class / compute_ap
(imported from /root/reference; build container only) on tests/synth.make_eval_case inputs.

    python tests/golden/make_golden_eval.py [--reference /root/reference]
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
CASES = [(1, 4), (2, 8), (3, 16)]          # (seed, images)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    import synth
    sys.path.insert(0, args.reference)
    import utils.utils as rutils            # the reference's (needs cv2, tqdm, torchvision: present in the build container)
    out = {}
    for seed, n in CASES:
        outs, tg = synth.make_eval_case(seed, n)
        stats = rutils.get_batch_statistics([torch.from_numpy(o) for o in outs], torch.from_numpy(tg), 0.5, torch.device("cpu"))
        for i, (tp, _, _) in enumerate(stats):
            out["c%d_tp%d" % (seed, i)] = np.asarray(tp, np.float64)
        tp = np.concatenate([s[0] for s in stats]); conf = np.concatenate([s[1].numpy() for s in stats])
        cls = np.concatenate([s[2].numpy() for s in stats])
        out["c%d_metrics" % seed] = np.array(rutils.ap_per_class(tp, conf, cls, tg[:, 1].tolist()), np.float64)
    rs = np.random.RandomState(9)
    rec = np.sort(rs.rand(30)); prec = np.sort(rs.rand(30))[::-1].copy()
    out["ap_rec"], out["ap_prec"], out["ap_value"] = rec, prec, np.float64(rutils.compute_ap(rec, prec))
    np.savez(os.path.join(HERE, "eval_cases.npz"), **out)
    print("wrote eval_cases.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
