"""CPU restatement (test infrastructure) of the reference's only active augmentation, utils/datasets.py:10-16:

    dst = cv2.addWeighted(img, alpha, blank, 1 - alpha, beta)          # blank = zeros

OpenCV is a third-party dependency of the reference (not vendored in /root/reference; 4.13.0 in the build container).  Its
published semantics for 8-bit inputs: the weighted sum is evaluated in fp32 and converted with saturate_cast<uchar>(cvRound(.)),
cvRound rounding half to even.  With a zero second image: saturate(rint(fl32(fl32(x * alpha) + beta))).  Pinned to outputs of the
real cv2.addWeighted in tests/golden/aug_cases.npz (tests/golden/make_golden_aug.py)."""
import numpy as np


def contrast_and_brightness(img, alpha, beta):
    x = np.asarray(img, dtype=np.uint8).astype(np.float32)
    t = x * np.float32(alpha) + np.float32(beta)                      # two fp32 roundings, like OpenCV's 8u path
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)
