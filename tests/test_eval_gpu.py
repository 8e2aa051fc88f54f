"""GPU tests of the evaluation path: yfv2_batch_statistics (through the C ABI) bit-exact against the reference's golden
flags and the oracle, the mirror's get_batch_statistics / evaluation against an evaluation assembled from oracle parts."""
import os

import numpy as np
import pytest
import torch

import yfv2  # noqa: F401
import synth
from oracle import evalstats, net as onet, post as opost

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n", [(1, 4), (2, 8), (3, 16)])
def test_batch_statistics_match_reference_golden(golden_dir, seed, n):
    import utils.utils as uu
    g = np.load(os.path.join(golden_dir, "eval_cases.npz"))
    outs, tg = synth.make_eval_case(seed, n)
    stats = uu.get_batch_statistics([torch.from_numpy(o) for o in outs], torch.from_numpy(tg).cuda(), 0.5, "cuda")
    assert len(stats) == n
    for i, (tp, conf, cls) in enumerate(stats):
        assert tp.dtype == np.float64 and np.array_equal(tp, g["c%d_tp%d" % (seed, i)]), (seed, i)
        assert np.array_equal(conf.numpy(), outs[i][:, 4]) and np.array_equal(cls.numpy(), outs[i][:, 5])


def test_batch_statistics_large_random_against_oracle():
    import utils.utils as uu
    outs, tg = synth.make_eval_case(77, 64, max_det=300, max_gt=13, classes=80)
    want = evalstats.get_batch_statistics(outs, tg, 0.5)
    got = uu.get_batch_statistics([torch.from_numpy(o) for o in outs], torch.from_numpy(tg).cuda(), 0.5, "cuda")
    for w, (tp, _, _) in zip(want, got):
        assert np.array_equal(tp, w)
    none_t = uu.get_batch_statistics([torch.from_numpy(o) for o in outs[:3]], torch.zeros((0, 6)).cuda(), 0.5, "cuda")
    assert all(float(tp.sum()) == 0 for tp, _, _ in none_t)


def test_evaluation_matches_oracle_pipeline():
    """utils.utils.evaluation over a small synthetic loader: forward -> decode -> NMS -> statistics -> AP, against the same
    pipeline assembled from the oracle's parts (NMS rows are bit-exact, so the statistics must be identical)."""
    import model.detector as det
    import utils.utils as uu
    sd = synth.make_state_dict(81)
    cfg = synth.coco_cfg(96, 64)
    m = det.Detector(80, 3, True)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    batches = []
    for b in range(2):
        imgs = (synth.make_images(82 + b, 3, 64, 96) * 255).to(torch.uint8)
        t = synth.make_targets(90 + b, 3)
        batches.append((imgs, t))
    got = uu.evaluation([(i.clone(), t.clone()) for i, t in batches], cfg, m, "cuda", conf_thres=0.01)
    labels, tps, confs, clss = [], [], [], []
    for imgs, t in batches:
        t = t.clone()
        labels += t[:, 1].tolist()
        t[:, 2:] = uu.xywh2xyxy(t[:, 2:])
        t[:, 2:] *= torch.tensor([cfg["width"], cfg["height"], cfg["width"], cfg["height"]])
        preds = m(imgs.cuda())                                     # NMS is bit-exact only on identical decoded input
        dets = uu.handel_preds(preds, cfg, "cuda")
        rows = opost.nms(dets, 0.01, 0.4)
        tp = evalstats.get_batch_statistics([r.numpy() for r in rows], t.numpy(), 0.5)
        tps += tp; confs += [r[:, 4].numpy() for r in rows]; clss += [r[:, 5].numpy() for r in rows]
    want = uu.ap_per_class(np.concatenate(tps), np.concatenate(confs), np.concatenate(clss), labels)
    np.testing.assert_array_equal(np.array(got), np.array(want))
