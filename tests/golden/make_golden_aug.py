"""Generates tests/golden/aug_cases.npz with the REAL cv2.addWeighted call of the reference's contrast_and_brightness
(utils/datasets.py:10-16; OpenCV is present in the build container only) on seeded inputs.

    python tests/golden/make_golden_aug.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rs = np.random.RandomState(17)
    out = {}
    # every byte value under 64 (alpha, beta) pairs from the reference's range, plus pairs that land on .5 boundaries
    ramp = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    pairs = [(rs.uniform(0.25, 1.75), rs.uniform(0.25, 1.75)) for _ in range(60)] + [(0.5, 0.5), (1.5, 0.5), (0.25, 0.25), (1.75, 1.75)]
    out["alpha"] = np.array([p[0] for p in pairs], np.float64)
    out["beta"] = np.array([p[1] for p in pairs], np.float64)
    out["ramp_dst"] = np.stack([cv2.addWeighted(ramp, a, np.zeros_like(ramp), 1 - a, b) for a, b in pairs])
    img = rs.randint(0, 256, (4, 37, 53, 3)).astype(np.uint8)          # odd sizes: the byte tail of the device kernel
    out["img"] = img
    out["img_dst"] = np.stack([cv2.addWeighted(img[i], pairs[i][0], np.zeros_like(img[i]), 1 - pairs[i][0], pairs[i][1]) for i in range(4)])
    np.savez_compressed(os.path.join(HERE, "aug_cases.npz"), **out)
    print("wrote aug_cases.npz (cv2 %s)" % cv2.__version__)


if __name__ == "__main__":
    main()
