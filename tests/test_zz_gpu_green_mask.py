"""GPU: wowsr_green_mask through the C ABI and the vector_extraction mirror vs the oracle and the reference goldens
(bit-exact bar: the mask is 0 / 1).  Sorted last on purpose: this kernel was added after the last hardware run of the
round, so a problem here cannot mask the parity tests of the main path."""
import os

import numpy as np
import pytest

from oracle import green_mask as G
from tests.conftest import image_like

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_goldens_from_reference(ws, handle):
    ve = ws.app.vector_extraction
    g = np.load(os.path.join(GOLD, "green_mask_u8_90x121.npz"))
    got = ve.green_mask_hsv_array(g["img"])
    assert got.dtype == np.float32 and np.array_equal(got, g["mask_default"])
    cfg = ve.ExtractionConfig(hsv_green_hue_range=tuple(int(v) for v in g["cfg2_hue"]), hsv_saturation_min=int(g["cfg2_sat"]),
                              hsv_value_min=int(g["cfg2_val"]))
    assert np.array_equal(ve.green_mask_hsv_array(g["img"], cfg), g["mask_cfg2"])
    g = np.load(os.path.join(GOLD, "green_mask_u16_64x80.npz"))
    assert np.array_equal(ve.green_mask_hsv_array(ve.normalise_rgb(g["raster"])), g["mask_default"])


@pytest.mark.parametrize("shape", [(1, 1), (7, 5), (64, 64), (517, 1003), (300, 201), (2048, 2048)])
def test_matches_oracle_ragged_sizes(ws, handle, shape):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    if shape[0] >= 64:
        img[: shape[0] // 2] = image_like(shape[0] // 2, shape[1], seed=5)
    assert np.array_equal(handle.green_mask_host(img, G.ranges()), G.green_mask_cv2(img))


def test_exhaustive_all_colours(ws, handle):
    """All 2^24 RGB triples in one 4096 x 4096 image."""
    v = np.arange(1 << 24, dtype=np.uint32).reshape(4096, 4096)
    img = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=-1).astype(np.uint8)
    got = handle.green_mask_host(img, G.ranges())
    assert np.array_equal(got, G.green_mask_cv2(img))
    assert np.array_equal(handle.green_mask_host(img, G.ranges((0, 179), 0, 0)[:1]), np.ones((4096, 4096), np.float32))


def test_device_resident_and_file_entry_points(ws, handle, tmp_path):
    import cv2
    import torch
    ve = ws.app.vector_extraction
    img = np.random.default_rng(3).integers(0, 256, (130, 257, 3), dtype=np.uint8)
    d = torch.from_numpy(img).cuda()
    m = ve.green_mask_hsv_cuda(d)
    torch.cuda.synchronize()
    assert np.array_equal(m.cpu().numpy(), G.green_mask_cv2(img))
    p = tmp_path / "rgb.png"
    cv2.imwrite(str(p), np.ascontiguousarray(img[:, :, ::-1]))
    assert np.array_equal(ve.compute_green_mask_hsv(p), G.green_mask_cv2(img))


def test_argument_errors(ws, handle):
    img = np.zeros((4, 4, 3), np.uint8)
    with pytest.raises(ws.WowsrError):
        handle.green_mask_host(img, [((0, 0, 0), (1, 1, 1))] * 5)          # more than 4 ranges
    with pytest.raises(ValueError):
        handle.green_mask_host(img, [((0, 0, 0), (300, 1, 1))])
    with pytest.raises(ValueError):
        handle.green_mask_host(np.zeros((4, 4), np.uint8), G.ranges())
