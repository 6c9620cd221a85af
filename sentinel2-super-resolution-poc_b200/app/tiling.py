"""Drop-in for the tile-pyramid part of ``server/app/tiling.py``: ``generate_xyz_tiles`` (:147-186, which
shells out to ``gdal2tiles.py --xyz --resampling average --tilesize 256``) and ``create_tileset_metadata`` (:189-219), fed
from the device-resident SR image instead of a GeoTIFF on disk.

PARITY UNPINNED: GDAL is not available offline and the reference has no test for its tiles.  The Web-Mercator tile grid
(``{z}/{x}/{y}.png``, y counted from the top: ``--xyz``) is the standard one and is pinned by closed-form checks; the
resampling is a restatement of "average" (``csrc/tiles.cu``), checked against ``oracle/tiling_np.py`` only.  Reprojection
(``reproject_to_3857``, gdalwarp) and ``get_raster_info`` (gdalinfo) stay with GDAL: the caller supplies the image's
EPSG:3857 bounds."""
from __future__ import annotations

import json
import math
from pathlib import Path
from typing import Sequence

import numpy as np
import torch

from .. import _lib
from .wow_sr import _handle

ORIGIN = 20037508.342789244          # half the Web-Mercator world width in metres


def resolution(z: int, tile_size: int = 256) -> float:
    """Metres per pixel of zoom level z."""
    return 2.0 * ORIGIN / (tile_size * (1 << z))


def tile_range(bounds_3857: Sequence[float], z: int):
    """(tx0, ty0, tx1, ty1), inclusive, of the XYZ tiles (y from the top) that intersect [west, south, east, north]."""
    west, south, east, north = bounds_3857
    span = 2.0 * ORIGIN / (1 << z)
    n = (1 << z) - 1
    eps = 1e-9 * span
    tx0 = min(max(int(math.floor((west + ORIGIN) / span)), 0), n)
    tx1 = min(max(int(math.floor((east + ORIGIN - eps) / span)), 0), n)
    ty0 = min(max(int(math.floor((ORIGIN - north) / span)), 0), n)
    ty1 = min(max(int(math.floor((ORIGIN - south - eps) / span)), 0), n)
    return tx0, ty0, tx1, ty1


def mosaic_geometry(bounds_3857, W: int, H: int, z: int, tile_size: int = 256):
    """Tile range of zoom z and the source-pixel mapping of its mosaic: (tx0, ty0, tx1, ty1, sx0, sy0, sxp, syp) — mosaic
    pixel (X, Y) covers source pixels [sx0 + X sxp, ...) x [sy0 + Y syp, ...)."""
    west, south, east, north = bounds_3857
    tx0, ty0, tx1, ty1 = tile_range(bounds_3857, z)
    res = resolution(z, tile_size)
    px_w, px_h = (east - west) / W, (north - south) / H           # metres per source pixel
    span = 2.0 * ORIGIN / (1 << z)
    left, top = -ORIGIN + tx0 * span, ORIGIN - ty0 * span        # mercator corner of the mosaic
    # snapped to 1e-9 source pixels: the bounds come out of sums of tile spans, and a footprint that leaks 1e-10 pixels into a
    # neighbour would move exact .5 averages (a quarter of all 2 x 2 means) across the rounding boundary
    return (tx0, ty0, tx1, ty1, round((left - west) / px_w, 9), round((north - top) / px_h, 9), round(res / px_w, 9), round(res / px_h, 9))


def generate_xyz_tiles_cuda(rgb: torch.Tensor, bounds_3857, output_dir: Path, min_zoom: int = 10, max_zoom: int = 16, tile_size: int = 256,
                            rows_per_pass: int = 8) -> Path:
    """``rgb``: HxWx3 uint8 CUDA tensor (north up, EPSG:3857 bounds [west, south, east, north]).  Writes ``{z}/{x}/{y}.png``
    (RGBA, transparent outside the raster; fully transparent tiles are skipped, like gdal2tiles does) under ``output_dir``."""
    import cv2
    assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous() and rgb.dim() == 3 and rgb.shape[2] == 3
    H, W = int(rgb.shape[0]), int(rgb.shape[1])
    h = _handle(rgb.device.index)
    src = _lib.Image(rgb.data_ptr(), W * 3, W, H, 0, H)
    output_dir = Path(output_dir)
    stream = torch.cuda.current_stream(rgb.device).cuda_stream
    for z in range(min_zoom, max_zoom + 1):
        tx0, ty0, tx1, ty1, sx0, sy0, sxp, syp = mosaic_geometry(bounds_3857, W, H, z, tile_size)
        ntx = tx1 - tx0 + 1
        for tya in range(ty0, ty1 + 1, rows_per_pass):               # strips of tile rows bound the mosaic's memory
            tyb = min(tya + rows_per_pass, ty1 + 1)
            OW, OH = ntx * tile_size, (tyb - tya) * tile_size
            mosaic = torch.empty((OH, OW, 4), dtype=torch.uint8, device=rgb.device)
            h.tiles_resample(src, sx0, sy0 + (tya - ty0) * tile_size * syp, sxp, syp, mosaic.data_ptr(), OW * 4, OW, OH, stream=stream)
            host = h.download(mosaic)
            for ty in range(tya, tyb):
                for tx in range(tx0, tx1 + 1):
                    t = host[(ty - tya) * tile_size:(ty - tya + 1) * tile_size, (tx - tx0) * tile_size:(tx - tx0 + 1) * tile_size]
                    if not t[:, :, 3].any():
                        continue
                    d = output_dir / str(z) / str(tx)
                    d.mkdir(parents=True, exist_ok=True)
                    cv2.imwrite(str(d / f"{ty}.png"), np.ascontiguousarray(t[:, :, [2, 1, 0, 3]]))     # cv2 writes BGRA
    return output_dir


def create_tileset_metadata(tiles_dir: Path, bounds_4326: list, min_zoom: int, max_zoom: int,
                            tile_template: str = "/tiles/{z}/{x}/{y}.png") -> dict:
    """Same dictionary and ``tileset.json`` as the reference (:189-219)."""
    metadata = {
        "bounds": bounds_4326,
        "minzoom": min_zoom,
        "maxzoom": max_zoom,
        "tileTemplate": tile_template,
        "attribution": "Sentinel-2 SR via UP42",
        "format": "png",
        "tileSize": 256,
    }
    tiles_dir = Path(tiles_dir)
    tiles_dir.mkdir(parents=True, exist_ok=True)
    with open(tiles_dir / "tileset.json", "w") as f:
        json.dump(metadata, f, indent=2)
    return metadata


def process_array_to_tiles(rgb: torch.Tensor, bounds_3857, bounds_4326, tiles_dir: Path, min_zoom: int = 10, max_zoom: int = 16) -> dict:
    """The tile half of ``process_raster_to_tiles`` (:222-…) for an image that is already on the device in EPSG:3857."""
    generate_xyz_tiles_cuda(rgb, bounds_3857, tiles_dir, min_zoom, max_zoom)
    return create_tileset_metadata(tiles_dir, list(bounds_4326), min_zoom, max_zoom)
