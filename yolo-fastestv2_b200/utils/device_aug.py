"""Batch-level, on-device form of the reference's img_aug (utils/datasets.py:63-68 -> contrast_and_brightness, :10-16).

The reference augments each uint8 HWC image on the host inside TensorDataset.__getitem__; with the input pipeline moved to the
GPU (uint8 batch copied first, `/255` fused into the stem) the same byte arithmetic runs on the whole uint8 batch in one
kernel of libyfv2.so (csrc/k_aug.cu), bit-identical to cv2.addWeighted.  (alpha, beta) are drawn per image from the same
random.uniform(0.25, 1.75) calls, in the same order as the reference would draw them image by image."""
import random

import torch

import yfv2_engine


def img_aug_batch(imgs_u8, rng=random, out=None):
    """imgs_u8: CUDA uint8 [N, ...] (any per-image layout).  Returns the augmented batch (new tensor unless `out` is given)."""
    n = imgs_u8.shape[0]
    ab = [(rng.uniform(0.25, 1.75), rng.uniform(0.25, 1.75)) for _ in range(n)]      # alpha first, then beta (datasets.py:11-12)
    alpha = torch.tensor([a for a, _ in ab], dtype=torch.float32)
    beta = torch.tensor([b for _, b in ab], dtype=torch.float32)
    return yfv2_engine.contrast_and_brightness(imgs_u8, alpha, beta, out=out)
