"""GPU: post-process kernels through the C ABI vs the oracle (bit-exact bar)."""
import os

import numpy as np
import pytest

from oracle import postproc_np as P
from oracle import wow_cv2
from tests.conftest import image_like

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True, params=["march", "tile"])
def post_kernel(request, ws, handle):
    """Every test of this module runs on both compiled pass-B kernels: the strip-march kernel (default) and the
    round-1 tile kernel (option post_kernel=0)."""
    v = 1 if request.param == "march" else 0
    handles = [handle, ws._lib.default_handle(0)]
    for h in handles:
        h.set_option("post_kernel", v)
    yield request.param
    for h in handles:
        h.set_option("post_kernel", 1)


def test_goldens_from_reference(ws, handle):
    g = np.load(os.path.join(GOLD, "post_wow_96x128.npz"))
    assert np.array_equal(ws.app.wow_sr._enhance_for_crops(g["img"]), g["out"])
    g = np.load(os.path.join(GOLD, "post_farm_101x77.npz"))
    assert np.array_equal(ws.app.farm_sr.farm_post(g["img"]), g["out"])
    img = g["img"]
    f = ws.app.farm_sr
    step = f.enhance_vegetation(f.apply_unsharp_mask(f.enhance_local_contrast(img, clip_limit=2.5, grid_size=8), strength=1.2, radius=1.5))
    assert np.array_equal(step, g["out"])          # the trio called one by one == fused pass == reference


@pytest.mark.parametrize("shape", [(64, 64), (512, 512), (517, 1003), (300, 200), (1104, 1104), (9, 40), (2304, 1728)])
def test_full_chain_bit_exact(ws, handle, shape):
    img = image_like(*shape, seed=shape[0])
    assert np.array_equal(handle.post_process_host(img, ws._lib.post_params("wow")), wow_cv2.enhance_for_crops(img))
    assert np.array_equal(handle.post_process_host(img, ws._lib.post_params("farm")), wow_cv2.farm_post(img))


def test_farm_trio_defaults(ws, handle):
    img = image_like(203, 310, seed=9)
    f = ws.app.farm_sr
    assert np.array_equal(f.enhance_local_contrast(img), P.enhance_local_contrast(img, 3.0, 8))
    assert np.array_equal(f.apply_unsharp_mask(img), P.apply_unsharp_mask(img, 1.5, 1.0))
    assert np.array_equal(f.enhance_vegetation(img), P.enhance_vegetation(img))


@pytest.mark.parametrize("shape", [(512, 512), (517, 1003), (4096, 4096)])
def test_clahe_hist_and_luts_bit_exact(ws, handle, shape):
    import torch
    img = image_like(*shape, seed=7)
    H, W = shape
    d = torch.from_numpy(img).cuda()
    tw, th, pw, ph = ws._lib.clahe_geometry(H, W, 8)
    hist = torch.zeros(64 * 256, dtype=torch.int32, device="cuda")
    luts = torch.zeros(64 * 256, dtype=torch.uint8, device="cuda")
    im = ws._lib.Image(d.data_ptr(), W * 3, W, H, 0, H)
    L = P.rgb2l_u8(img)
    ref_hist = P.clahe_hist(L, 8)
    for variant in (0, 1, 2):            # every compiled histogram update: one-bin shortcut, match_any grouping, plain atomics (default)
        hist.zero_()
        handle.set_option("hist_match", variant)
        try:
            handle.clahe_hist(im, 8, 0, ph, hist.data_ptr())
        finally:
            handle.set_option("hist_match", 2)
        torch.cuda.synchronize()
        assert np.array_equal(hist.cpu().numpy().astype(np.uint32).reshape(8, 8, 256), ref_hist), f"hist_match={variant}"
    handle.clahe_luts(hist.data_ptr(), 8, tw, th, 2.5, luts.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(luts.cpu().numpy().reshape(8, 8, 256), P.clahe_luts(ref_hist, tw * th, 2.5))
    assert int(hist.sum()) == pw * ph


def test_band_split_equals_whole(ws, handle):
    """Multi-GPU decomposition on one GPU: two bands (+halo) with summed histograms == whole image."""
    import torch
    H, W = 1100, 900
    img = image_like(H, W, seed=21)
    p = ws._lib.post_params("farm")
    want = wow_cv2.farm_post(img)
    tw, th, pw, ph = ws._lib.clahe_geometry(H, W, 8)
    d = torch.from_numpy(img).cuda()
    hist = torch.zeros(64 * 256, dtype=torch.int32, device="cuda")
    luts = torch.zeros(64 * 256, dtype=torch.uint8, device="cuda")
    split, r = 537, 4
    bands = [(0, split), (split, H)]
    for (a, b) in bands:                                   # each "rank" sees only its own rows
        sub = d[a:b].contiguous()
        im = ws._lib.Image(sub.data_ptr(), W * 3, W, H, a, b - a)
        handle.clahe_hist(im, 8, a, b if b < H else ph, hist.data_ptr())
    handle.clahe_luts(hist.data_ptr(), 8, tw, th, 2.5, luts.data_ptr())
    out = torch.zeros_like(d)
    for (a, b) in bands:
        lo, hi = max(a - r, 0), min(b + r, H)               # halo rows from the neighbour
        sub = d[lo:hi].contiguous()
        src = ws._lib.Image(sub.data_ptr(), W * 3, W, H, lo, hi - lo)
        dst_t = out[a:b]
        dst = ws._lib.Image(dst_t.data_ptr(), W * 3, W, H, a, b - a)
        handle.post_apply(src, luts.data_ptr(), p, a, b, dst)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)


def test_idempotent_launch_and_error_paths(ws, handle):
    img = image_like(64, 64)
    a = handle.post_process_host(img, ws._lib.post_params("wow"))
    b = handle.post_process_host(img, ws._lib.post_params("wow"))
    assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        handle.post_process_host(img[..., 0], ws._lib.post_params("wow"))
    with pytest.raises(ws.WowsrError):
        handle.post_process_host(img, ws._lib.post_params("wow", sigma=9.0))


@pytest.mark.parametrize("shape", [(8, 8), (9, 9), (1, 40), (40, 1), (33, 31)])
def test_tiny_and_degenerate_shapes(ws, handle, shape):
    """Smaller than the CLAHE grid / blur radius: reflect-101 folds several times, tiles are 1-2 pixels."""
    img = image_like(max(shape[0], 16), max(shape[1], 16), seed=shape[0] + 3)[:shape[0], :shape[1]].copy()
    try:
        want = wow_cv2.enhance_for_crops(img)
    except Exception:
        pytest.skip("cv2 rejects this shape")
    assert np.array_equal(handle.post_process_host(img, ws._lib.post_params("wow")), want)


def test_constant_and_extreme_images(ws, handle):
    for val in (0, 255, 128):
        img = np.full((128, 96, 3), val, np.uint8)
        assert np.array_equal(handle.post_process_host(img, ws._lib.post_params("wow")), wow_cv2.enhance_for_crops(img))
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (257, 263, 3), dtype=np.uint8)          # white noise: every LUT bin populated
    assert np.array_equal(handle.post_process_host(img, ws._lib.post_params("farm")), wow_cv2.farm_post(img))


def test_full_size_properties_4096(ws, handle):
    """BASELINE config 4 size (4096x4096): histogram mass conservation, determinism, and agreement with cv2."""
    import torch
    img = np.tile(image_like(1024, 1024, seed=3), (4, 4, 1))
    d = torch.from_numpy(img).cuda()
    hist = torch.zeros(64 * 256, dtype=torch.int32, device="cuda")
    handle.clahe_hist(ws._lib.Image(d.data_ptr(), 4096 * 3, 4096, 4096, 0, 4096), 8, 0, 4096, hist.data_ptr())
    torch.cuda.synchronize()
    assert int(hist.sum()) == 4096 * 4096 and (hist.view(64, 256).sum(1) == 512 * 512).all()
    a = handle.post_process_host(img, ws._lib.post_params("wow"))
    assert np.array_equal(a, handle.post_process_host(img, ws._lib.post_params("wow")))
    assert np.array_equal(a, wow_cv2.enhance_for_crops(img))


def test_clahe_tiles_larger_than_2_pow_24_and_full_chain_stripe(ws, handle):
    """cfg5's regime: CLAHE tiles of more than 2^24 pixels (32800 = 8 x 4100 -> 16.81 M per tile, clipLimit
    > 10^5), where cv2's float arithmetic on the clip limit and the uint32 -> float LUT scaling leave the range in which every
    integer is exact.  Histograms and LUTs bit-exact against the restatement, the CLAHE stage bit-exact against cv2 on the whole
    1.08 Gpix image, and the full chain (CLAHE -> unsharp -> vegetation boost) bit-exact on a stripe assembled from pinned cv2
    pieces: cv2 CLAHE on the whole L plane, the rest on the stripe with its blur halo."""
    import cv2
    import torch
    S = 32800
    small = image_like(S // 4, S // 4, seed=11)
    img = np.ascontiguousarray(np.repeat(np.repeat(small, 4, 0), 4, 1))
    del small
    d = torch.from_numpy(img).cuda()
    tw, th, pw, ph = ws._lib.clahe_geometry(S, S, 8)
    assert tw * th > 1 << 24 and (pw, ph) == (S, S)
    hist = torch.zeros(64 * 256, dtype=torch.int32, device="cuda")
    luts = torch.zeros(64 * 256, dtype=torch.uint8, device="cuda")
    im = ws._lib.Image(d.data_ptr(), S * 3, S, S, 0, S)
    handle.clahe_hist(im, 8, 0, ph, hist.data_ptr())
    handle.clahe_luts(hist.data_ptr(), 8, tw, th, 2.5, luts.data_ptr())
    torch.cuda.synchronize()
    lab = cv2.cvtColor(img, cv2.COLOR_RGB2LAB)
    L = np.ascontiguousarray(lab[:, :, 0])
    ref_hist = np.stack([np.stack([np.bincount(L[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw].ravel(), minlength=256)
                                   for tx in range(8)]) for ty in range(8)]).astype(np.uint32)
    assert np.array_equal(hist.cpu().numpy().astype(np.uint32).reshape(8, 8, 256), ref_hist)
    ref_luts = P.clahe_luts(ref_hist, tw * th, 2.5)
    assert np.array_equal(luts.cpu().numpy().reshape(8, 8, 256), ref_luts)
    # CLAHE stage on the whole image against cv2 itself
    lab[:, :, 0] = cv2.createCLAHE(clipLimit=2.5, tileGridSize=(8, 8)).apply(L)
    del L
    enh = cv2.cvtColor(lab, cv2.COLOR_LAB2RGB)
    del lab
    out = torch.empty_like(d)
    p_clahe = ws._lib.post_params("wow", stages=ws._lib.STAGE_CLAHE)
    handle.post_process_dev(d.data_ptr(), out.data_ptr(), S, S, p_clahe)
    torch.cuda.synchronize()
    for y in range(0, S, 4100):   # compare in slabs: keeps the host copies small
        assert np.array_equal(out[y:y + 4100].cpu().numpy(), enh[y:y + 4100]), y
    # full chain on a stripe that crosses a CLAHE tile boundary (rows 4100 * 4 = 16400): the remaining stages are local
    # (blur radius 3 after zero taps), so cv2 on the stripe plus a 16-row halo reproduces the whole-image result
    handle.post_process_dev(d.data_ptr(), out.data_ptr(), S, S, ws._lib.post_params("wow"))
    torch.cuda.synchronize()
    y0, y1, hl = 16200, 16600, 16
    e = enh[y0 - hl:y1 + hl]
    sharp = cv2.addWeighted(e, 1.4, cv2.GaussianBlur(e, (0, 0), 1.2), -0.4, 0)
    hsv = cv2.cvtColor(sharp, cv2.COLOR_RGB2HSV).astype(np.float32)
    mask = (hsv[:, :, 0] > 35) & (hsv[:, :, 0] < 85)
    hsv[:, :, 1] = np.where(mask, np.clip(hsv[:, :, 1] * 1.2, 0, 255), hsv[:, :, 1])
    ref = cv2.cvtColor(hsv.astype(np.uint8), cv2.COLOR_HSV2RGB)[hl:-hl]
    assert np.array_equal(out[y0:y1].cpu().numpy(), ref)


@pytest.mark.parametrize("nt,seg", [(64, 8), (64, 1), (96, 37), (256, 256), (128, 5)])
def test_march_kernel_any_strip_and_segment_geometry(ws, handle, post_kernel, nt, seg):
    """Strip-march kernel with forced strip widths (threads per CTA) and segment heights: many strips and segments on a
    small image, segments shorter than the blur radius, a partial last strip, one strip wider than the image."""
    if post_kernel != "march":
        pytest.skip("geometry options belong to the strip-march kernel")
    img = image_like(333, 1001, seed=5)
    handle.set_option("post_nt", nt)
    handle.set_option("post_seg", seg)
    try:
        assert np.array_equal(handle.post_process_host(img, ws._lib.post_params("wow")), wow_cv2.enhance_for_crops(img))
        assert np.array_equal(handle.post_process_host(img, ws._lib.post_params("farm")), wow_cv2.farm_post(img))
        big = ws._lib.post_params("farm", sigma=2.0)                       # radius 6 -> the 8-column halo instantiation
        assert np.array_equal(handle.post_process_host(img, big),
                              wow_cv2.post_process(img, clip=2.5, grid=8, sigma=2.0, alpha=2.2, beta=-1.2, sat=1.3))
    finally:
        handle.set_option("post_nt", 0)
        handle.set_option("post_seg", 0)


def test_unaligned_pitch_and_stage_subsets(ws, handle):
    """Device entry point on a packed image whose row size is not a multiple of 4 (byte loads / stores instead of 32-bit
    ones), and every subset of the three stages."""
    import torch
    H, W = 203, 301
    img = image_like(H, W, seed=13)
    d = torch.from_numpy(img).cuda()
    out = torch.zeros_like(d)
    handle.post_process_dev(d.data_ptr(), out.data_ptr(), H, W, ws._lib.post_params("wow"))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), wow_cv2.enhance_for_crops(img))
    import cv2
    for stages in range(1, 8):
        want = img
        if stages & 1:
            lab = cv2.cvtColor(want, cv2.COLOR_RGB2LAB)
            lab[:, :, 0] = cv2.createCLAHE(clipLimit=2.5, tileGridSize=(8, 8)).apply(lab[:, :, 0])
            want = cv2.cvtColor(lab, cv2.COLOR_LAB2RGB)
        if stages & 2:
            want = cv2.addWeighted(want, 1.4, cv2.GaussianBlur(want, (0, 0), 1.2), -0.4, 0)
        if stages & 4:
            hsv = cv2.cvtColor(want, cv2.COLOR_RGB2HSV).astype(np.float32)
            mask = (hsv[:, :, 0] > 35) & (hsv[:, :, 0] < 85)
            hsv[:, :, 1] = np.where(mask, np.clip(hsv[:, :, 1] * 1.2, 0, 255), hsv[:, :, 1])
            want = cv2.cvtColor(hsv.astype(np.uint8), cv2.COLOR_HSV2RGB)
        got = handle.post_process_host(img, ws._lib.post_params("wow", stages=stages))
        assert np.array_equal(got, want), f"stages={stages}"


@pytest.mark.parametrize("grid", [2, 4, 16])
def test_other_clahe_grids(ws, handle, grid):
    """CLAHE grids other than 8: up to 8 the LUTs are staged in shared memory, above they are read from global memory
    (enhance_local_contrast takes grid_size as a parameter, farm_sr.py:61-69)."""
    img = image_like(517, 403, seed=grid)
    got = handle.post_process_host(img, ws._lib.post_params("wow", grid=grid))
    assert np.array_equal(got, wow_cv2.post_process(img, grid=grid))
