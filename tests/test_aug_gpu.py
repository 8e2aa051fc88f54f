"""Device contrast_and_brightness (through the C ABI) against the real cv2 goldens and the oracle: bit-exact bytes."""
import os

import numpy as np
import pytest
import torch

import yfv2  # noqa: F401
import yfv2_engine as eng
from oracle import aug as oaug

pytestmark = pytest.mark.gpu


def test_bit_exact_against_cv2_golden(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "aug_cases.npz")))
    ramp = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    n = len(g["alpha"])
    batch = torch.from_numpy(np.stack([ramp] * n)).cuda()
    got = eng.contrast_and_brightness(batch, g["alpha"], g["beta"]).cpu().numpy()
    assert np.array_equal(got, g["ramp_dst"])
    img = torch.from_numpy(g["img"]).cuda()                          # 37x53x3 images: exercises the unaligned byte path
    got = eng.contrast_and_brightness(img, g["alpha"][:4], g["beta"][:4]).cpu().numpy()
    assert np.array_equal(got, g["img_dst"])


def test_batch_in_place_and_against_oracle():
    rs = np.random.RandomState(5)
    x = rs.randint(0, 256, (8, 3, 352, 352)).astype(np.uint8)         # the training batch layout (train.py:101), 16-byte path
    a, b = rs.uniform(0.25, 1.75, 8), rs.uniform(0.25, 1.75, 8)
    d = torch.from_numpy(x).cuda()
    eng.contrast_and_brightness(d, a, b, out=d)
    want = np.stack([oaug.contrast_and_brightness(x[i], a[i], b[i]) for i in range(8)])
    assert np.array_equal(d.cpu().numpy(), want)
