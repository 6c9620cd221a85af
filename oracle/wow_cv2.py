"""TEST INFRASTRUCTURE ONLY — the reference post-process call sequence, restated over cv2.

The reference's post-process *is* a sequence of OpenCV calls (wow_sr.py:190-207,
farm_sr.py:66-69,79-86,94-106 as called at farm_sr.py:170-178).  This port issues the same calls
with the same constants; it is the CPU baseline timed by bench.py (kind="port") and the pinned
truth the numpy restatement (postproc_np.py) and the CUDA kernels are checked against.
"""
from __future__ import annotations

import cv2
import numpy as np


def post_process(img: np.ndarray, clip=2.5, grid=8, sigma=1.2, alpha=1.4, beta=-0.4,
                 hue_lo=35, hue_hi=85, sat=1.2) -> np.ndarray:
    lab = cv2.cvtColor(img, cv2.COLOR_RGB2LAB)
    lab[:, :, 0] = cv2.createCLAHE(clipLimit=clip, tileGridSize=(grid, grid)).apply(lab[:, :, 0])
    enhanced = cv2.cvtColor(lab, cv2.COLOR_LAB2RGB)
    blurred = cv2.GaussianBlur(enhanced, (0, 0), sigma)
    sharp = cv2.addWeighted(enhanced, alpha, blurred, beta, 0)
    hsv = cv2.cvtColor(sharp, cv2.COLOR_RGB2HSV).astype(np.float32)
    mask = (hsv[:, :, 0] > hue_lo) & (hsv[:, :, 0] < hue_hi)
    hsv[:, :, 1] = np.where(mask, np.clip(hsv[:, :, 1] * sat, 0, 255), hsv[:, :, 1])
    return cv2.cvtColor(hsv.astype(np.uint8), cv2.COLOR_HSV2RGB)


def enhance_for_crops(img):
    """wow_sr._enhance_for_crops (wow_sr.py:187-209)."""
    return post_process(img)


def farm_post(img):
    """farm_sr.apply_farm_sr steps 2-4 (farm_sr.py:170-178): CLAHE 2.5/8, unsharp 1.2/1.5, boost 1.3."""
    return post_process(img, clip=2.5, grid=8, sigma=1.5, alpha=2.2, beta=-1.2, sat=1.3)
