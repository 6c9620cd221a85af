"""TEST INFRASTRUCTURE ONLY — CPU restatement of the HSV vegetation mask, compute_green_mask_hsv
(server/app/vector_extraction.py:222-270), after the raster has been read and normalised (:237-249).

Two statements of the same arithmetic: ``green_mask_cv2`` issues the reference's own OpenCV calls (:252-270);
``green_mask_np`` restates them over the pinned integer RGB->HSV of postproc_np — cv2.inRange is
``lo <= x <= hi`` on every channel.  Pinned by tests/golden/green_mask_*.npz, produced by the unmodified reference
function through a stub ``rasterio.open`` (tests/golden/make_golden.py)."""
from __future__ import annotations

import numpy as np

from . import postproc_np

BROWN = ((10, 20, 40), (35, 200, 200))  # vector_extraction.py:262-263


def ranges(hue=(35, 85), sat_min=30, val_min=30):
    """The reference's two boxes: green from ExtractionConfig (:57-59, :255-257) and the fixed brown one."""
    return [((hue[0], sat_min, val_min), (hue[1], 255, 255)), BROWN]


def normalise_rgb(rgb: np.ndarray) -> np.ndarray:
    """:245-249"""
    if rgb.max() > 255:
        return (rgb / rgb.max() * 255).astype(np.uint8)
    return rgb.astype(np.uint8)


def green_mask_cv2(rgb: np.ndarray, rng=None) -> np.ndarray:
    import cv2
    hsv = cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV)
    combined = None
    for lo, hi in (rng or ranges()):
        m = cv2.inRange(hsv, np.array(lo), np.array(hi))
        combined = m if combined is None else cv2.bitwise_or(combined, m)
    return (combined > 0).astype(np.float32)


def green_mask_np(rgb: np.ndarray, rng=None) -> np.ndarray:
    hsv = postproc_np.rgb2hsv_u8(rgb).astype(np.int64)
    out = np.zeros(rgb.shape[:2], dtype=bool)
    for lo, hi in (rng or ranges()):
        m = np.ones(rgb.shape[:2], dtype=bool)
        for c in range(3):
            m &= (hsv[..., c] >= lo[c]) & (hsv[..., c] <= hi[c])
        out |= m
    return out.astype(np.float32)
