"""TEST INFRASTRUCTURE ONLY — numpy restatement of the WOW / farm post-process arithmetic.

This module is the CPU oracle for the post-process half of the hot path.  It is imported only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg, never by the
product package.

What it restates (reference call sites; the arithmetic itself lives in OpenCV, which is a
third-party dependency that is *not* under /root/reference — ``opencv-contrib-python>=4.8.0``,
``server/requirements.txt:29``; pinned here by execution against cv2 4.13.0, see
``tests/test_oracle_postproc.py`` and ``tests/golden/make_golden.py``):

* ``wow_sr._enhance_for_crops``      server/app/wow_sr.py:187-209
* ``farm_sr.enhance_local_contrast`` server/app/farm_sr.py:74-88
* ``farm_sr.apply_unsharp_mask``     server/app/farm_sr.py:61-71
* ``farm_sr.enhance_vegetation``     server/app/farm_sr.py:91-108

Every function is integer arithmetic or exactly specified fp32, following SURVEY.md Appendix A.
"""
from __future__ import annotations

import functools

import numpy as np

f32 = np.float32


def _rint(x):
    return np.rint(x)


# ----------------------------------------------------------------------------------------------
# tables (built once; float32 arithmetic where cv2 uses float)
# ----------------------------------------------------------------------------------------------

@functools.lru_cache(maxsize=None)
def lab_tables():
    """Tables for COLOR_RGB2LAB 8u (cv2 ``RGB2Lab_b``): gamma LUT (x2040) and cube-root LUT (x32768)."""
    i = np.arange(256, dtype=np.float32) / f32(255.0)
    lin = np.where(i <= f32(0.04045), i / f32(12.92),
                   np.power((i + f32(0.055)) / f32(1.055), f32(2.4), dtype=np.float32)).astype(np.float32)
    gam = _rint(lin * f32(2040.0)).astype(np.int32)
    j = np.arange(3072, dtype=np.float32) / f32(2040.0)
    cb = np.where(j < f32(216.0 / 24389.0),
                  j * f32(841.0 / 108.0) + f32(16.0 / 116.0),
                  np.cbrt(j, dtype=np.float32)).astype(np.float32)
    cbrt = _rint(cb * f32(32768.0)).astype(np.int32)
    return gam, cbrt


@functools.lru_cache(maxsize=None)
def lab2rgb_tables():
    """Tables for COLOR_LAB2RGB 8u (cv2 ``Lab2RGBinteger``): y[L], ify[L], ab2xz[], inverse gamma."""
    BASE = 16384
    y = np.zeros(256, dtype=np.int64)
    ify = np.zeros(256, dtype=np.int64)
    for l in range(256):
        if l <= 20:
            # y = L / 903.3 ; ify = 16/116 + 7.787 * y, scaled by BASE
            y[l] = int(np.rint(np.float32(l * BASE * 180.0 / (17.0 * 29 ** 3))))
            ify[l] = int(np.rint(np.float32(BASE * (16.0 / 116.0 + 5.0 * l / (3.0 * 17.0 * 29.0)))))
        else:
            fy = np.float32(l * 100.0 * BASE / (255.0 * 116.0) + 16.0 * BASE / 116.0)
            ify[l] = int(np.rint(fy))
            y[l] = int(np.rint(np.float32(fy) * np.float32(fy) * np.float32(fy) / np.float32(BASE * BASE)))
    n = 8145 + 28719
    ab2xz = np.zeros(n, dtype=np.int64)
    for idx in range(n):
        i = idx - 8145
        if i <= 3390:
            # C integer division (truncate toward zero)
            v = i * 108
            q = abs(v) // 841
            v = q if v >= 0 else -q
            ab2xz[idx] = v - 290
        else:
            ab2xz[idx] = ((i * i) // BASE * i) // BASE
    k = np.arange(4096, dtype=np.float32) / f32(4096.0)
    ig = np.where(k <= f32(0.0031308), k * f32(12.92),
                  f32(1.055) * np.power(k, f32(1.0 / 2.4), dtype=np.float32) - f32(0.055)).astype(np.float32)
    invgam = _rint(f32(255.0) * ig).astype(np.int32)
    return y, ify, ab2xz, invgam


@functools.lru_cache(maxsize=None)
def hsv_tables():
    """sdiv/hdiv tables for COLOR_RGB2HSV 8u with hrange=180 (cv2 ``RGB2HSV_b``)."""
    sdiv = np.zeros(256, dtype=np.int64)
    hdiv = np.zeros(256, dtype=np.int64)
    for i in range(1, 256):
        sdiv[i] = int(np.rint((255 << 12) / float(i)))
        hdiv[i] = int(np.rint((180 << 12) / (6.0 * i)))
    return sdiv, hdiv


def gaussian_kernel_u8(sigma: float):
    """8-bit fixed-point Gaussian taps used by cv2.GaussianBlur(ksize=(0,0)) on CV_8U (App. A.4)."""
    ksize = int(np.rint(sigma * 6 + 1)) | 1
    # cv2.getGaussianKernel in float64
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    k /= k.sum()
    taps = np.zeros(ksize, dtype=np.int64)
    err = 0.0
    half = ksize // 2
    for i in range(half):
        adj = k[i] * 256.0 + err
        v = np.rint(adj)
        err = adj - v
        taps[i] = int(v)
        taps[ksize - 1 - i] = int(v)
    taps[half] = 256 - 2 * int(taps[:half].sum())
    return taps


def sboost_table(k: float):
    """S' = trunc(min(f32(S) * f32(k), 255)) (wow_sr.py:203-205, farm_sr.py:100-104)."""
    s = np.arange(256, dtype=np.float32)
    return np.minimum(s * f32(k), f32(255.0)).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# colour conversions
# ----------------------------------------------------------------------------------------------

def _ds(x, n):
    return (x + (1 << (n - 1))) >> n


def rgb2lab_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, COLOR_RGB2LAB) for uint8 HxWx3 (Appendix A.1)."""
    gam, cbrt = lab_tables()
    R = gam[img[..., 0]].astype(np.int64)
    G = gam[img[..., 1]].astype(np.int64)
    B = gam[img[..., 2]].astype(np.int64)
    fX = cbrt[_ds(1777 * R + 1541 * G + 778 * B, 12)].astype(np.int64)
    fY = cbrt[_ds(871 * R + 2929 * G + 296 * B, 12)].astype(np.int64)
    fZ = cbrt[_ds(73 * R + 448 * G + 3575 * B, 12)].astype(np.int64)
    L = _ds(296 * fY - 1336934, 15)
    a = _ds(500 * (fX - fY) + 128 * 32768, 15)
    b = _ds(200 * (fY - fZ) + 128 * 32768, 15)
    out = np.stack([L, a, b], axis=-1)
    return np.clip(out, 0, 255).astype(np.uint8)


def rgb2l_u8(img: np.ndarray) -> np.ndarray:
    """Only the L plane of RGB->Lab (what CLAHE pass A needs)."""
    gam, cbrt = lab_tables()
    R = gam[img[..., 0]].astype(np.int64)
    G = gam[img[..., 1]].astype(np.int64)
    B = gam[img[..., 2]].astype(np.int64)
    fY = cbrt[_ds(871 * R + 2929 * G + 296 * B, 12)].astype(np.int64)
    return np.clip(_ds(296 * fY - 1336934, 15), 0, 255).astype(np.uint8)


def lab2rgb_u8(lab: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(lab, COLOR_LAB2RGB) for uint8 HxWx3 (Appendix A.3)."""
    y_t, ify_t, ab2xz, invgam = lab2rgb_tables()
    L = lab[..., 0].astype(np.int64)
    a = lab[..., 1].astype(np.int64)
    b = lab[..., 2].astype(np.int64)
    y = y_t[L]
    ify = ify_t[L]
    adiv = ((5 * a * 53687 + 128) >> 13) - 4194
    bdiv = ((b * 41943 + 16) >> 9) - 10485 + 1
    X = ab2xz[ify + adiv + 8145]
    Z = ab2xz[ify - bdiv + 8145]
    r = _ds(12615 * X - 6296 * y - 2223 * Z, 14)
    g = _ds(-3773 * X + 7684 * y + 185 * Z, 14)
    bl = _ds(217 * X - 836 * y + 4715 * Z, 14)
    out = np.stack([r, g, bl], axis=-1)
    out = np.clip(out, 0, 4095)
    return invgam[out].astype(np.uint8)


def rgb2hsv_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, COLOR_RGB2HSV) for uint8 (H in [0,180)) (Appendix A.6)."""
    sdiv, hdiv = hsv_tables()
    r = img[..., 0].astype(np.int64)
    g = img[..., 1].astype(np.int64)
    b = img[..., 2].astype(np.int64)
    v = np.maximum(np.maximum(r, g), b)
    mn = np.minimum(np.minimum(r, g), b)
    diff = v - mn
    s = (diff * sdiv[v] + 2048) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * hdiv[diff] + 2048) >> 12
    h = np.where(h < 0, h + 180, h)
    return np.stack([h, s, v], axis=-1).astype(np.uint8)


def hsv2rgb_u8(hsv: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(hsv, COLOR_HSV2RGB) for uint8 HxWx3, as computed by the AVX2/FMA3 cv2 build.

    App. A.6 plus one amendment found while pinning this oracle (tests/test_oracle_postproc.py):
    cv2 converts each row in 32-pixel SIMD groups; the last ``W % 32`` pixels of every row go
    through its scalar tail, which evaluates the same FMA-contracted expressions but quantises
    with round-half-even (``saturate_cast``) instead of the SIMD body's truncation.  Verified
    exhaustively over H<180 x S x V for both the body and the tail.
    """
    H = hsv[..., 0].astype(np.float32)
    S = hsv[..., 1].astype(np.float32)
    V = hsv[..., 2].astype(np.float32)
    s = S * f32(1.0 / 255.0)
    v = V * f32(1.0 / 255.0)
    h6 = H * f32(6.0 / 180.0)
    sec = np.floor(h6)
    f = h6 - sec
    sec = sec.astype(np.int64)
    # sec can be >= 6 only for H >= 180, which RGB2HSV never produces; cv2 wraps it.
    sec = np.where(sec >= 6, sec - 6, sec)
    # fma(-s, f, 1): single rounding -> evaluate in float64 (exact product of two f32, one add) then round
    s64 = s.astype(np.float64)
    f64 = f.astype(np.float64)
    one_minus_f = (f32(1.0) - f).astype(np.float32)
    t0 = v
    t1 = v * (f32(1.0) - s)
    t2 = v * (1.0 - s64 * f64).astype(np.float32)
    t3 = v * (1.0 - s64 * one_minus_f.astype(np.float64)).astype(np.float32)
    tab = np.stack([t0, t1, t2, t3], axis=-1)
    # sector table gives (b, g, r) indices into t
    sector = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]], dtype=np.int64)
    idx = sector[sec]                                   # ... x 3 (b,g,r)
    bgr = np.take_along_axis(tab, idx, axis=-1) * f32(255.0)
    out = bgr.astype(np.int64)                          # SIMD body: truncation
    if hsv.ndim == 3:
        W = hsv.shape[1]
        t = W - W % 32
        if t < W:                                       # scalar tail: round half even
            out[:, t:] = np.rint(bgr[:, t:]).astype(np.int64)
    out = np.clip(out, 0, 255)
    return np.stack([out[..., 2], out[..., 1], out[..., 0]], axis=-1).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# CLAHE
# ----------------------------------------------------------------------------------------------

def _reflect101(i, n):
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * n - 2 - i, i)


def clahe_geometry(H: int, W: int, grid: int = 8):
    """(tile_w, tile_h, padded_W, padded_H) — cv2 pads right/bottom when not divisible (App. A.2)."""
    if W % grid == 0 and H % grid == 0:
        return W // grid, H // grid, W, H
    pw = W + (grid - W % grid)
    ph = H + (grid - H % grid)
    return pw // grid, ph // grid, pw, ph


def clahe_hist(L: np.ndarray, grid: int = 8) -> np.ndarray:
    """Per-tile 256-bin histograms of the (reflect-101 padded) L plane: uint32 [grid, grid, 256]."""
    H, W = L.shape
    tw, th, pw, ph = clahe_geometry(H, W, grid)
    if (pw, ph) != (W, H):
        ys = _reflect101(np.arange(ph), H)
        xs = _reflect101(np.arange(pw), W)
        Lp = L[np.ix_(ys, xs)]
    else:
        Lp = L
    hist = np.zeros((grid, grid, 256), dtype=np.uint32)
    for ty in range(grid):
        for tx in range(grid):
            t = Lp[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            hist[ty, tx] = np.bincount(t.ravel(), minlength=256).astype(np.uint32)
    return hist


def clahe_luts(hist: np.ndarray, tile_area: int, clip: float) -> np.ndarray:
    """Clip / redistribute / cumsum -> uint8 LUTs [grid, grid, 256] (App. A.2)."""
    gy, gx, _ = hist.shape
    clip_limit = max(int(clip * tile_area / 256), 1)
    lut_scale = f32(255.0) / f32(tile_area)
    luts = np.zeros((gy, gx, 256), dtype=np.uint8)
    for ty in range(gy):
        for tx in range(gx):
            h = hist[ty, tx].astype(np.int64)
            clipped = int(np.maximum(h - clip_limit, 0).sum())
            h = np.minimum(h, clip_limit)
            batch = clipped // 256
            resid = clipped - 256 * batch
            h += batch
            if resid:
                step = max(256 // resid, 1)
                k = 0
                while k < 256 and resid > 0:
                    h[k] += 1
                    k += step
                    resid -= 1
            cs = np.cumsum(h)
            v = np.rint(cs.astype(np.float32) * lut_scale)
            luts[ty, tx] = np.clip(v, 0, 255).astype(np.uint8)
    return luts


def clahe_interp(L: np.ndarray, luts: np.ndarray, tw: int, th: int) -> np.ndarray:
    """Bilinear LUT interpolation in fp32 with separately rounded multiplies/adds (App. A.2)."""
    H, W = L.shape
    gy, gx, _ = luts.shape
    inv_tw = f32(1.0) / f32(tw)
    inv_th = f32(1.0) / f32(th)
    x = np.arange(W, dtype=np.float32)
    txf = x * inv_tw - f32(0.5)
    tx1 = np.floor(txf).astype(np.int64)
    xa = (txf - tx1.astype(np.float32)).astype(np.float32)
    xa1 = (f32(1.0) - xa).astype(np.float32)
    tx2 = np.minimum(tx1 + 1, gx - 1)
    tx1 = np.maximum(tx1, 0)
    y = np.arange(H, dtype=np.float32)
    tyf = y * inv_th - f32(0.5)
    ty1 = np.floor(tyf).astype(np.int64)
    ya = (tyf - ty1.astype(np.float32)).astype(np.float32)
    ya1 = (f32(1.0) - ya).astype(np.float32)
    ty2 = np.minimum(ty1 + 1, gy - 1)
    ty1 = np.maximum(ty1, 0)
    out = np.empty((H, W), dtype=np.uint8)
    lf = luts.astype(np.float32)
    # row blocks to bound memory
    for y0 in range(0, H, 256):
        y1 = min(y0 + 256, H)
        v = L[y0:y1].astype(np.int64)
        a = ty1[y0:y1, None]
        b = ty2[y0:y1, None]
        p00 = lf[a, tx1[None, :], v]
        p01 = lf[a, tx2[None, :], v]
        p10 = lf[b, tx1[None, :], v]
        p11 = lf[b, tx2[None, :], v]
        top = (p00 * xa1[None, :]).astype(np.float32) + (p01 * xa[None, :]).astype(np.float32)
        bot = (p10 * xa1[None, :]).astype(np.float32) + (p11 * xa[None, :]).astype(np.float32)
        res = (top * ya1[y0:y1, None]).astype(np.float32) + (bot * ya[y0:y1, None]).astype(np.float32)
        out[y0:y1] = np.clip(np.rint(res), 0, 255).astype(np.uint8)
    return out


def clahe_u8(L: np.ndarray, clip: float = 2.5, grid: int = 8) -> np.ndarray:
    """cv2.createCLAHE(clip, (grid, grid)).apply(L) (App. A.2)."""
    H, W = L.shape
    tw, th, _, _ = clahe_geometry(H, W, grid)
    hist = clahe_hist(L, grid)
    luts = clahe_luts(hist, tw * th, clip)
    return clahe_interp(L, luts, tw, th)


# ----------------------------------------------------------------------------------------------
# blur / addWeighted
# ----------------------------------------------------------------------------------------------

def gaussian_blur_u8(img: np.ndarray, sigma: float) -> np.ndarray:
    """cv2.GaussianBlur(img, (0,0), sigma) for uint8 HxWxC: exact fixed-point, single rounding (App. A.4)."""
    taps = gaussian_kernel_u8(sigma)
    r = len(taps) // 2
    H, W = img.shape[:2]
    src = img.astype(np.int64)
    xs = _reflect101(np.arange(-r, W + r), W)
    ys = _reflect101(np.arange(-r, H + r), H)
    hp = src[:, xs]
    hacc = np.zeros_like(src)
    for k, t in enumerate(taps):
        if t:
            hacc += int(t) * hp[:, k:k + W]
    vp = hacc[ys]
    vacc = np.zeros_like(src)
    for k, t in enumerate(taps):
        if t:
            vacc += int(t) * vp[k:k + H]
    return ((vacc + 32768) >> 16).astype(np.uint8)


def add_weighted_u8(a: np.ndarray, alpha: float, b: np.ndarray, beta: float) -> np.ndarray:
    """cv2.addWeighted(a, alpha, b, beta, 0) for uint8 (App. A.5)."""
    t = a.astype(np.float32) * f32(alpha) + b.astype(np.float32) * f32(beta)
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# the pipelines
# ----------------------------------------------------------------------------------------------

WOW_PARAMS = dict(clip=2.5, grid=8, sigma=1.2, alpha=1.4, beta=-0.4, hue_lo=35, hue_hi=85, sat=1.2)
FARM_PARAMS = dict(clip=2.5, grid=8, sigma=1.5, alpha=2.2, beta=-1.2, hue_lo=35, hue_hi=85, sat=1.3)


def enhance_local_contrast(img, clip=3.0, grid=8):
    lab = rgb2lab_u8(img)
    lab[..., 0] = clahe_u8(lab[..., 0], clip, grid)
    return lab2rgb_u8(lab)


def apply_unsharp_mask(img, strength=1.5, radius=1.0):
    blurred = gaussian_blur_u8(img, radius)
    return add_weighted_u8(img, 1.0 + strength, blurred, -strength)


def enhance_vegetation(img, sat=1.3, hue_lo=35, hue_hi=85):
    hsv = rgb2hsv_u8(img)
    boost = sboost_table(sat)
    mask = (hsv[..., 0] > hue_lo) & (hsv[..., 0] < hue_hi)
    hsv[..., 1] = np.where(mask, boost[hsv[..., 1]], hsv[..., 1])
    return hsv2rgb_u8(hsv)


def post_process(img: np.ndarray, p: dict) -> np.ndarray:
    enhanced = enhance_local_contrast(img, p["clip"], p["grid"])
    blurred = gaussian_blur_u8(enhanced, p["sigma"])
    sharp = add_weighted_u8(enhanced, p["alpha"], blurred, p["beta"])
    return enhance_vegetation(sharp, p["sat"], p["hue_lo"], p["hue_hi"])


def enhance_for_crops(img: np.ndarray) -> np.ndarray:
    """wow_sr._enhance_for_crops (wow_sr.py:187-209)."""
    return post_process(img, WOW_PARAMS)


def farm_post(img: np.ndarray) -> np.ndarray:
    """farm_sr.apply_farm_sr steps 2-4 as called at farm_sr.py:170-178."""
    return post_process(img, FARM_PARAMS)
