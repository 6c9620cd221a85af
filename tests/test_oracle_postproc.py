"""CPU: pins the numpy post-process oracle against cv2 (the third-party arithmetic the reference calls,
wow_sr.py:190-207), against the committed reference goldens, and — when /root/reference is present —
against the unmodified reference functions."""
import os

import cv2
import numpy as np
import pytest

from oracle import postproc_np as P
from oracle import refload, wow_cv2
from tests.conftest import image_like

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _all_rgb():
    full = np.arange(1 << 24, dtype=np.uint32)
    return np.stack([(full >> 16) & 255, (full >> 8) & 255, full & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)


def test_table_checksums():
    gam, cbrt = P.lab_tables()
    y, ify, ab, ig = P.lab2rgb_tables()
    sd, hd = P.hsv_tables()
    assert (gam[1], gam[10], gam[128], gam[255], int(gam.sum())) == (1, 6, 440, 2040, 162416)
    assert (cbrt[0], cbrt[2040], cbrt[3071], int(cbrt.sum())) == (4520, 32768, 37555, 86529539)
    assert (y[20], y[21], y[255], int(y.sum())) == (142, 149, 16384, 1219054)
    assert (ify[0], ify[255], int(ify.sum())) == (2260, 16384, 2386418)
    assert (ab[0], ab[-1], int(ab.sum())) == (-1335, 88231, 626487776)
    assert (ig[1], ig[13], ig[100], ig[2048], ig[4095], int(ig.sum())) == (1, 10, 43, 188, 255, 720284)
    assert (int(sd.sum()), int(hd.sum())) == (6392676, 752077)
    assert list(P.gaussian_kernel_u8(1.2)) == [0, 4, 21, 60, 86, 60, 21, 4, 0]
    assert list(P.gaussian_kernel_u8(1.5)) == [0, 2, 9, 28, 55, 68, 55, 28, 9, 2, 0]


@pytest.mark.parametrize("name,fn,code", [("rgb2lab", P.rgb2lab_u8, cv2.COLOR_RGB2LAB),
                                          ("lab2rgb", P.lab2rgb_u8, cv2.COLOR_LAB2RGB),
                                          ("rgb2hsv", P.rgb2hsv_u8, cv2.COLOR_RGB2HSV)])
def test_colour_conversions_exhaustive(name, fn, code):
    img = _all_rgb()
    assert np.array_equal(fn(img), cv2.cvtColor(img, code))


def test_hsv2rgb_exhaustive_body_and_tail():
    h = np.arange(180 * 256 * 256, dtype=np.uint32)
    hsv = np.stack([h // 65536, (h >> 8) & 255, h & 255], -1).astype(np.uint8)
    body = hsv.reshape(180 * 16, 4096, 3)            # width % 32 == 0: SIMD body only (truncation)
    assert np.array_equal(P.hsv2rgb_u8(body), cv2.cvtColor(body, cv2.COLOR_HSV2RGB))
    n = (len(hsv) // 31) * 31
    tail = hsv[:n].reshape(-1, 31, 3)                # width 31: every pixel goes through the scalar tail (rounding)
    assert np.array_equal(P.hsv2rgb_u8(tail), cv2.cvtColor(tail, cv2.COLOR_HSV2RGB))


@pytest.mark.parametrize("shape", [(64, 64), (512, 512), (517, 1003), (300, 200)])
def test_clahe_blur_addweighted(shape):
    img = image_like(*shape)
    L = cv2.cvtColor(img, cv2.COLOR_RGB2LAB)[..., 0].copy()
    for clip in (2.5, 3.0):
        assert np.array_equal(P.clahe_u8(L, clip, 8), cv2.createCLAHE(clipLimit=clip, tileGridSize=(8, 8)).apply(L))
    for s in (1.2, 1.5, 1.0):
        assert np.array_equal(P.gaussian_blur_u8(img, s), cv2.GaussianBlur(img, (0, 0), s))
    a = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 256, 1)
    for al, be in ((1.4, -0.4), (2.2, -1.2)):
        assert np.array_equal(P.add_weighted_u8(a, al, a.T.copy(), be), cv2.addWeighted(a, al, a.T.copy(), be, 0))


@pytest.mark.parametrize("shape", [(96, 128), (517, 1003), (640, 620)])
def test_full_chain_vs_cv2_port(shape):
    img = image_like(*shape, seed=5)
    assert np.array_equal(P.enhance_for_crops(img), wow_cv2.enhance_for_crops(img))
    assert np.array_equal(P.farm_post(img), wow_cv2.farm_post(img))


def test_goldens_from_reference():
    g = np.load(os.path.join(GOLD, "post_wow_96x128.npz"))
    assert np.array_equal(P.enhance_for_crops(g["img"]), g["out"])
    assert np.array_equal(wow_cv2.enhance_for_crops(g["img"]), g["out"])
    g = np.load(os.path.join(GOLD, "post_farm_101x77.npz"))
    assert np.array_equal(P.farm_post(g["img"]), g["out"])
    assert np.array_equal(wow_cv2.farm_post(g["img"]), g["out"])


@pytest.mark.skipif(not refload.available(), reason="reference tree not present")
def test_against_unmodified_reference():
    _, wow, farm = refload.load()
    img = image_like(203, 310, seed=9)
    assert np.array_equal(P.enhance_for_crops(img), wow._enhance_for_crops(img))
    ref = farm.enhance_vegetation(farm.apply_unsharp_mask(farm.enhance_local_contrast(img, 2.5, 8), 1.2, 1.5))
    assert np.array_equal(P.farm_post(img), ref)
    # the trio individually, with the reference's default arguments
    assert np.array_equal(P.enhance_local_contrast(img, 3.0, 8), farm.enhance_local_contrast(img))
    assert np.array_equal(P.apply_unsharp_mask(img, 1.5, 1.0), farm.apply_unsharp_mask(img))
    assert np.array_equal(P.enhance_vegetation(img), farm.enhance_vegetation(img))
