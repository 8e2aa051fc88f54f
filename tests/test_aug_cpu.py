"""The augmentation restatement against outputs of the real cv2.addWeighted (tests/golden/aug_cases.npz)."""
import os

import numpy as np

from oracle import aug as oaug


def test_oracle_matches_cv2_addweighted(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "aug_cases.npz")))
    ramp = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    for i, (a, b) in enumerate(zip(g["alpha"], g["beta"])):
        assert np.array_equal(oaug.contrast_and_brightness(ramp, a, b), g["ramp_dst"][i]), (i, a, b)
    for i in range(4):
        assert np.array_equal(oaug.contrast_and_brightness(g["img"][i], g["alpha"][i], g["beta"][i]), g["img_dst"][i])
