"""XYZ tile pyramid (PARITY UNPINNED — no GDAL anywhere): the mercator tile arithmetic against the closed-form
slippy-map formulas (CPU), the tileset metadata against the reference's dictionary, and on the GPU the resampling kernel against
the numpy restatement of the same definition plus pyramid properties."""
import json
import math
import os

import numpy as np
import pytest

from oracle import tiling_np as T


def _mod():
    import importlib
    return importlib.import_module("sentinel2-super-resolution-poc_b200.app.tiling")


def test_tile_grid_matches_slippy_map_formulas():
    tiling = _mod()
    assert abs(tiling.resolution(0) - 156543.03392804097) < 1e-6            # the well-known zoom-0 resolution
    for lon, lat in ((-121.487, 36.836), (13.405, 52.52), (0.001, -0.001), (151.2, -33.87)):
        x, y = T.lonlat_to_3857(lon, lat)
        for z in (0, 3, 10, 14, 18):
            tx0, ty0, tx1, ty1 = tiling.tile_range((x, y, x + 1e-3, y + 1e-3), z)
            assert (tx0, ty0) == T.slippy_tile(lon, lat, z) and (tx1, ty1) == (tx0, ty0), (lon, lat, z)
    # an extent that spans tiles: exactly the tiles its corners fall into, y counted from the top (--xyz)
    w, s = T.lonlat_to_3857(-121.60, 36.75)
    e, n = T.lonlat_to_3857(-121.40, 36.90)
    tx0, ty0, tx1, ty1 = tiling.tile_range((w, s, e, n), 12)
    assert (tx0, ty0) == T.slippy_tile(-121.60, 36.90, 12) and (tx1, ty1) == T.slippy_tile(-121.40, 36.75, 12)
    assert ty0 <= ty1 and tx0 <= tx1
    # an extent aligned to a tile edge does not spill into the next tile
    span = 2 * tiling.ORIGIN / (1 << 5)
    assert tiling.tile_range((-tiling.ORIGIN, tiling.ORIGIN - span, -tiling.ORIGIN + span, tiling.ORIGIN), 5) == (0, 0, 0, 0)


def test_mosaic_geometry_maps_the_raster_onto_itself():
    tiling = _mod()
    # a raster that IS tile (z=3, x=2, y=5) at 256 x 256: the mosaic is that tile and maps pixel to pixel
    span = 2 * tiling.ORIGIN / 8
    b = (-tiling.ORIGIN + 2 * span, tiling.ORIGIN - 6 * span, -tiling.ORIGIN + 3 * span, tiling.ORIGIN - 5 * span)
    tx0, ty0, tx1, ty1, sx0, sy0, sxp, syp = tiling.mosaic_geometry(b, 256, 256, 3)
    assert (tx0, ty0, tx1, ty1) == (2, 5, 2, 5) and abs(sx0) < 1e-6 and abs(sy0) < 1e-6 and abs(sxp - 1) < 1e-9 and abs(syp - 1) < 1e-9
    _, _, _, _, _, _, sxp2, _ = tiling.mosaic_geometry(b, 256, 256, 2)
    assert abs(sxp2 - 2) < 1e-9                                            # one zoom level out: two source pixels per mosaic pixel


def test_tileset_metadata_matches_the_reference_dictionary(tmp_path):
    tiling = _mod()
    md = tiling.create_tileset_metadata(tmp_path / "tiles", [-121.6, 36.7, -121.4, 36.9], 10, 16)
    assert md == {"bounds": [-121.6, 36.7, -121.4, 36.9], "minzoom": 10, "maxzoom": 16, "tileTemplate": "/tiles/{z}/{x}/{y}.png",
                  "attribution": "Sentinel-2 SR via UP42", "format": "png", "tileSize": 256}      # tiling.py:209-217
    assert json.load(open(tmp_path / "tiles" / "tileset.json")) == md


@pytest.mark.gpu
@pytest.mark.parametrize("geom", [(0.0, 0.0, 1.0, 1.0), (-3.25, -2.5, 2.0, 2.0), (1.3, 0.7, 3.7, 2.9), (-10.0, -6.0, 0.5, 0.5), (5.5, 3.25, 7.0, 1.5)])
def test_resample_kernel_vs_numpy_restatement(ws, handle, geom):
    import torch
    sx0, sy0, sxp, syp = geom
    src = np.random.default_rng(5).integers(0, 256, (37, 53, 3), dtype=np.uint8)
    OW, OH = 40, 28
    d = torch.from_numpy(src).cuda()
    out = torch.full((OH, OW, 4), 7, dtype=torch.uint8, device="cuda")
    handle.tiles_resample(ws._lib.Image(d.data_ptr(), 53 * 3, 53, 37, 0, 37), sx0, sy0, sxp, syp, out.data_ptr(), OW * 4, OW, OH)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), T.resample(src, sx0, sy0, sxp, syp, OW, OH))


@pytest.mark.gpu
def test_pyramid_files_and_properties(ws, tmp_path):
    import cv2
    import torch
    tiling = _mod()
    # a 512 x 512 raster covering exactly tiles (z=11; x=330..331, y=790..791): z=11 tiles are crops, z=10 is the 2 x 2 average
    span = 2 * tiling.ORIGIN / (1 << 11)
    b = (-tiling.ORIGIN + 330 * span, tiling.ORIGIN - 792 * span, -tiling.ORIGIN + 332 * span, tiling.ORIGIN - 790 * span)
    img = np.random.default_rng(6).integers(0, 256, (512, 512, 3), dtype=np.uint8)
    out = tiling.process_array_to_tiles(torch.from_numpy(img).cuda(), b, [-122.0, 36.0, -121.6, 36.4], tmp_path / "t", 9, 12)
    assert out["minzoom"] == 9 and os.path.exists(tmp_path / "t" / "tileset.json")
    for (tx, ty) in ((330, 790), (331, 790), (330, 791), (331, 791)):
        t = cv2.imread(str(tmp_path / "t" / "11" / str(tx) / f"{ty}.png"), cv2.IMREAD_UNCHANGED)
        assert t.shape == (256, 256, 4) and (t[:, :, 3] == 255).all()
        crop = img[(ty - 790) * 256:(ty - 789) * 256, (tx - 330) * 256:(tx - 329) * 256]
        assert np.array_equal(t[:, :, [2, 1, 0]], crop)
    t10 = cv2.imread(str(tmp_path / "t" / "10" / "165" / "395.png"), cv2.IMREAD_UNCHANGED)
    want = np.floor(img.reshape(256, 2, 256, 2, 3).astype(np.float64).mean(axis=(1, 3)) + 0.5).astype(np.uint8)
    assert np.array_equal(t10[:, :, [2, 1, 0]], want) and (t10[:, :, 3] == 255).all()
    t9 = cv2.imread(str(tmp_path / "t" / "9" / "82" / "197.png"), cv2.IMREAD_UNCHANGED)      # raster = one quadrant of this tile
    assert t9.shape == (256, 256, 4) and int((t9[:, :, 3] == 255).sum()) == 128 * 128
    t12 = cv2.imread(str(tmp_path / "t" / "12" / "660" / "1580.png"), cv2.IMREAD_UNCHANGED)   # finer than the source: nearest
    assert np.array_equal(t12[:, :, [2, 1, 0]], np.repeat(np.repeat(img[:128, :128], 2, 0), 2, 1))
    assert len(list((tmp_path / "t" / "12").glob("*/*.png"))) == 16
