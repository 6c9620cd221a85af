"""CPU: the HSV vegetation-mask oracle (oracle/green_mask.py) against the goldens produced by the unmodified reference
(compute_green_mask_hsv, vector_extraction.py:222-270), against the reference itself when the tree is present, and an
exhaustive check of the numpy restatement against cv2 over all 2^24 colours."""
import os

import numpy as np
import pytest

from oracle import green_mask as G
from oracle import refload

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_goldens_from_reference():
    g = np.load(os.path.join(GOLD, "green_mask_u8_90x121.npz"))
    img = g["img"]
    for fn in (G.green_mask_cv2, G.green_mask_np):
        assert np.array_equal(fn(img), g["mask_default"])
        assert np.array_equal(fn(img, G.ranges(tuple(g["cfg2_hue"]), int(g["cfg2_sat"]), int(g["cfg2_val"]))), g["mask_cfg2"])
    assert 0.2 < g["mask_default"].mean() < 0.8            # the fixture exercises both outcomes
    g = np.load(os.path.join(GOLD, "green_mask_u16_64x80.npz"))
    n = G.normalise_rgb(g["raster"])
    assert n.dtype == np.uint8 and n.max() == 255
    assert np.array_equal(G.green_mask_np(n), g["mask_default"])
    assert np.array_equal(G.green_mask_cv2(n), g["mask_default"])


def test_exhaustive_all_colours_np_equals_cv2():
    """Every RGB triple, in 16 slabs of 2^20 colours: restated RGB->HSV + inclusive range test == cv2.cvtColor + cv2.inRange."""
    g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for r0 in range(0, 256, 16):
        img = np.empty((16 * 256, 256, 3), dtype=np.uint8)
        for i in range(16):
            img[i * 256:(i + 1) * 256, :, 0] = r0 + i
            img[i * 256:(i + 1) * 256, :, 1] = g
            img[i * 256:(i + 1) * 256, :, 2] = b
        assert np.array_equal(G.green_mask_np(img), G.green_mask_cv2(img)), r0


def test_mirror_config_matches_reference_defaults(ws):
    ve = ws.app.vector_extraction
    cfg = ve.ExtractionConfig()
    assert ve.hsv_ranges(cfg) == G.ranges()
    assert (cfg.hsv_green_hue_range, cfg.hsv_saturation_min, cfg.hsv_value_min) == ((35, 85), 30, 30)   # vector_extraction.py:57-59
    raster = np.random.default_rng(0).integers(0, 5000, (8, 9, 3)).astype(np.uint16)
    assert np.array_equal(ve.normalise_rgb(raster), G.normalise_rgb(raster))
    assert np.array_equal(ve.normalise_rgb(raster // 32), G.normalise_rgb(raster // 32))                  # max <= 255: plain cast


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_oracle_equals_reference_function():
    store = {}
    ve = refload.load_vector_extraction(lambda p: store[str(p)])
    rng = np.random.default_rng(77)
    img = rng.integers(0, 256, (53, 67, 3), dtype=np.uint8)
    store["x"] = [img[..., c] for c in range(3)]
    cfg = ve.ExtractionConfig(hsv_green_hue_range=(40, 70), hsv_saturation_min=60, hsv_value_min=10)
    want = ve.compute_green_mask_hsv("x", cfg)
    assert np.array_equal(G.green_mask_np(img, G.ranges((40, 70), 60, 10)), want)
    assert want.dtype == np.float32 and set(np.unique(want)) <= {0.0, 1.0}
