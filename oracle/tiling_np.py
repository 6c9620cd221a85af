"""TEST INFRASTRUCTURE ONLY — numpy restatement of the tile-pyramid resampling (csrc/tiles.cu).

PARITY UNPINNED.  The reference produces its tiles with ``gdal2tiles.py --xyz --resampling average`` (server/app/tiling.py:
147-186); neither GDAL nor any test or golden of the reference covers it, so there is nothing to pin against.  This file
restates the DEFINITION the CUDA kernel implements (area-weighted mean of the covered source pixels in float64, rounded half
up; nearest source pixel when the mosaic is finer than the source; opaque where the mosaic pixel's centre is inside the
raster) so that the kernel is at least checked against an independent implementation of the same definition, and the
Web-Mercator tile arithmetic against the closed-form slippy-map formulas."""
from __future__ import annotations

import math

import numpy as np


def resample(src: np.ndarray, sx0: float, sy0: float, sxp: float, syp: float, OW: int, OH: int) -> np.ndarray:
    H, W = src.shape[:2]
    out = np.zeros((OH, OW, 4), np.uint8)
    for Y in range(OH):
        ay = sy0 + Y * syp
        by = ay + syp
        cy = 0.5 * (ay + by)
        if not (0.0 <= cy < H):
            continue
        for X in range(OW):
            ax = sx0 + X * sxp
            bx = ax + sxp
            cx = 0.5 * (ax + bx)
            if not (0.0 <= cx < W):
                continue
            if sxp <= 1.0 and syp <= 1.0:
                out[Y, X, :3] = src[min(max(int(math.floor(cy)), 0), H - 1), min(max(int(math.floor(cx)), 0), W - 1)]
                out[Y, X, 3] = 255
                continue
            x0, x1 = max(int(math.floor(ax)), 0), min(int(math.ceil(bx)), W)
            y0, y1 = max(int(math.floor(ay)), 0), min(int(math.ceil(by)), H)
            acc = np.zeros(3, np.float64)
            wsum = 0.0
            for y in range(y0, y1):
                wy = min(y + 1.0, by) - max(float(y), ay)
                if wy <= 0.0:
                    continue
                for x in range(x0, x1):
                    wx = min(x + 1.0, bx) - max(float(x), ax)
                    if wx <= 0.0:
                        continue
                    acc += (wx * wy) * src[y, x].astype(np.float64)
                    wsum += wx * wy
            if wsum > 0.0:
                out[Y, X, :3] = np.minimum(255, np.floor(acc / wsum + 0.5)).astype(np.uint8)
                out[Y, X, 3] = 255
    return out


def slippy_tile(lon: float, lat: float, z: int):
    """The textbook slippy-map tile of a WGS84 point (independent of app/tiling.py's mercator arithmetic)."""
    n = 1 << z
    x = int((lon + 180.0) / 360.0 * n)
    y = int((1.0 - math.asinh(math.tan(math.radians(lat))) / math.pi) / 2.0 * n)
    return x, y


def lonlat_to_3857(lon: float, lat: float):
    r = 6378137.0
    return math.radians(lon) * r, math.log(math.tan(math.pi / 4.0 + math.radians(lat) / 2.0)) * r
