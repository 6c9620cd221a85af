"""IO glue of the /api/wow and /api/sr paths: raster normalisation, model residency, and the
file-in / file-out drop-ins ``apply_wow_sr`` / ``process_wow_sr`` / ``apply_farm_sr`` (wow_sr.py:27-266, farm_sr.py:110-285)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rrdbnet_ref as R


def _ref_normalise(img):
    """wow_sr.py:66-72, verbatim arithmetic."""
    if img.dtype != np.uint8:
        if img.max() > 255:
            return ((img - img.min()) / (img.max() - img.min()) * 255).astype(np.uint8)
        return img.astype(np.uint8)
    return img


@pytest.mark.parametrize("case", ["u16_wide", "u16_narrow", "i32_small", "u8", "f32_wide", "f64_wide", "u16_constant", "f32_constant"])
def test_normalise_matches_reference_arithmetic(ws, case):
    """Host logic (runs on CPU tensors too): bit-exact with the reference's numpy expression."""
    rng = np.random.default_rng(11)
    if case == "u16_wide":
        a = rng.integers(0, 10000, (37, 41, 3)).astype(np.uint16)
    elif case == "u16_narrow":
        a = rng.integers(300, 4000, (20, 33, 3)).astype(np.uint16)
    elif case == "i32_small":
        a = rng.integers(0, 256, (16, 16, 3)).astype(np.uint16)      # max <= 255: plain cast branch
    elif case == "f32_wide":
        a = (rng.random((29, 31, 3)) * 9000).astype(np.float32)      # numpy stretches float32 rasters in float32
    elif case == "f64_wide":
        a = rng.random((29, 31, 3)) * 9000
    elif case == "u16_constant":
        a = np.full((8, 9, 3), 300, np.uint16)                       # 0/0 = NaN -> astype(uint8) gives 0
    elif case == "f32_constant":
        a = np.full((8, 9, 3), 300.0, np.float32)
    else:
        a = rng.integers(0, 256, (16, 16, 3)).astype(np.uint8)
    with np.errstate(all="ignore"):
        want = _ref_normalise(a)
    t = torch.from_numpy(a.astype(np.int32) if a.dtype == np.uint16 else a)
    got = ws.app.wow_sr.normalise_to_uint8_cuda(t).numpy()
    assert got.dtype == np.uint8 and np.array_equal(got, want)


def test_normalise_with_the_cnn_sr_epsilon(ws):
    """apply_cnn_sr stretches with ``+ 1e-6`` in the denominator (cnn_super_resolution.py:308-311): the maximum maps to 254."""
    rng = np.random.default_rng(12)
    for a in (rng.integers(0, 10000, (31, 17, 3)).astype(np.uint16), rng.integers(100, 300, (9, 9, 3)).astype(np.uint16),
              rng.integers(0, 200, (9, 9, 3)).astype(np.uint16)):
        img = a
        if img.max() > 255:
            img = (img - img.min()) / (img.max() - img.min() + 1e-6) * 255
        want = img.astype(np.uint8)
        got = ws.app.wow_sr.normalise_to_uint8_cuda(torch.from_numpy(a.astype(np.int32)), eps=1e-6).numpy()
        assert np.array_equal(got, want)
        if a.max() > 255:
            assert got.max() == 254


@pytest.mark.gpu
def test_apply_wow_sr_files_metadata_and_residency(ws, tmp_path, monkeypatch):
    import cv2
    cnn, wow = ws.app.cnn_super_resolution, ws.app.wow_sr
    # the reference loads `<model_dir>/<model>.pth` (cnn_super_resolution.py:55-70): provide seeded weights there
    sd = R.calibrate_conv_last(R.random_init_state_dict(5, 6), 6)
    monkeypatch.setattr(cnn, "get_model_dir", lambda: tmp_path)
    torch.save({"params_ema": sd}, tmp_path / "realesrgan_anime.pth")
    cnn.clear_model_cache()
    rgb = np.random.default_rng(7).integers(0, 256, (45, 61, 3), dtype=np.uint8)
    src = tmp_path / "scene.png"
    cv2.imwrite(str(src), np.ascontiguousarray(rgb[:, :, ::-1]))

    out_path, meta = wow.apply_wow_sr(src, tmp_path / "out" / "scene_wow.tif", enhance_crops=True, model="realesrgan_anime")
    assert out_path == tmp_path / "out" / "scene_wow.png" and out_path.exists()
    assert set(meta) == {"input_file", "output_file", "scale", "pipeline", "stages", "enhancements", "original_size", "output_size",
                         "original_resolution_m", "effective_resolution_m", "optimized_for"}
    assert meta["scale"] == 4 and meta["original_size"] == [45, 61] and meta["output_size"] == [180, 244]
    assert meta["effective_resolution_m"] == 2.5 and len(meta["stages"]) == 2
    got = cv2.imread(str(out_path))[:, :, ::-1]
    # same result as the in-memory calls the reference makes: enhance(BGR) -> RGB -> _enhance_for_crops
    up = cnn.RealESRGAN(model_name="realesrgan_anime", tile_size=256)
    sr_rgb = np.ascontiguousarray(up.enhance(np.ascontiguousarray(rgb[:, :, ::-1]))[:, :, ::-1])
    assert np.array_equal(got, wow._enhance_for_crops(sr_rgb))
    # residency: the second construction reuses the loaded device image
    up2 = cnn.RealESRGAN(model_name="realesrgan_anime", tile_size=512)
    assert up2._h is up._h and up2.tile_size == 512
    n0 = len(cnn._MODEL_CACHE)
    _, meta2 = wow.apply_wow_sr(src, tmp_path / "out2" / "plain", enhance_crops=False, model="realesrgan_anime")
    assert len(cnn._MODEL_CACHE) == n0 and meta2["enhancements"] == [] and len(meta2["stages"]) == 1
    assert np.array_equal(cv2.imread(str(tmp_path / "out2" / "plain.png"))[:, :, ::-1], sr_rgb)

    res = wow.process_wow_sr(src, tmp_path / "job", enhance_crops=True, model="realesrgan_anime")
    assert set(res) == {"timestamp", "input", "outputs", "sr_metadata"}
    assert res["outputs"]["sr_tif"] is None and res["outputs"]["sr_png"].endswith("scene_wow_sr.png")
    assert json.load(open(tmp_path / "job" / "scene_wow_sr_metadata.json"))["sr_metadata"]["scale"] == 4
    with pytest.raises(ValueError):
        wow.apply_wow_sr(src, tmp_path / "x", model="nope")
    cnn.clear_model_cache()


@pytest.mark.gpu
def test_apply_farm_sr_files(ws, tmp_path, monkeypatch):
    import cv2
    cnn, farm = ws.app.cnn_super_resolution, ws.app.farm_sr
    blocks = cnn.MODELS["realesrgan_x4"]["blocks"]
    sd = R.calibrate_conv_last(R.random_init_state_dict(2, blocks), blocks)
    monkeypatch.setattr(cnn, "get_model_dir", lambda: tmp_path)
    torch.save(sd, tmp_path / "realesrgan_x4.pth")
    cnn.clear_model_cache()
    rgb = np.random.default_rng(8).integers(0, 256, (24, 40, 3), dtype=np.uint8)
    src = tmp_path / "field.png"
    cv2.imwrite(str(src), np.ascontiguousarray(rgb[:, :, ::-1]))
    res = farm.process_farm_sr(src, tmp_path / "job", scale=4)
    assert res["outputs"]["sr_png"].endswith("field_farm_sr_x4.png") and res["sr_metadata"]["model"] == "RealESRGAN_farm_x4"
    got = cv2.imread(res["outputs"]["sr_png"])[:, :, ::-1]
    up = cnn.RealESRGAN(scale=4, tile_size=256)
    sr_rgb = np.ascontiguousarray(up.enhance(np.ascontiguousarray(rgb[:, :, ::-1]))[:, :, ::-1])
    # the reference's three separate calls (farm_sr.py:170-178) == the fused pass
    step = farm.enhance_vegetation(farm.apply_unsharp_mask(farm.enhance_local_contrast(sr_rgb, clip_limit=2.5, grid_size=8),
                                                           strength=1.2, radius=1.5))
    assert np.array_equal(got, step)
    cnn.clear_model_cache()


def test_geotiff_branches_of_read_and_write_image(ws, tmp_path, monkeypatch):
    """rasterio is absent from this image: run the GeoTIFF branches of the IO glue (wow_sr.py:57-79, :126-164) against an
    in-memory stand-in (tests/fake_rasterio.py) so they are exercised before production."""
    from tests import fake_rasterio as FR
    FR.install(monkeypatch)
    wow = ws.app.wow_sr
    rng = np.random.default_rng(3)
    bands = [rng.integers(0, 9000, (12, 17)).astype(np.uint16) for _ in range(4)]          # 4-band Sentinel-2 style raster
    FR.put(tmp_path / "s2.tif", bands)
    img, transform, crs = wow.read_image(tmp_path / "s2.tif")
    assert img.shape == (12, 17, 3) and img.dtype == np.uint16 and crs == "EPSG:32636"
    assert all(np.array_equal(img[..., i], bands[i]) for i in range(3))                      # bands 1-3 as R, G, B
    FR.put(tmp_path / "gray.tiff", bands[:1])
    g, _, _ = wow.read_image(tmp_path / "gray.tiff")
    assert g.shape == (12, 17, 3) and all(np.array_equal(g[..., i], bands[0]) for i in range(3))   # one band replicated
    rgb = rng.integers(0, 256, (48, 68, 3), dtype=np.uint8)
    out = wow.write_image(rgb, tmp_path / "o" / "scene_wow.tif", transform, crs, 4)
    assert out == tmp_path / "o" / "scene_wow.tif" and out.exists() and (tmp_path / "o" / "scene_wow.png").exists()
    rec = FR.STORE[str(out)]
    assert rec["transform"] == FR.Affine(2.5, 0.0, 5e5, 0.0, -2.5, 4e6) and rec["crs"] == crs      # pixel size / scale (:131-138)
    assert rec["kw"]["driver"] == "GTiff" and rec["kw"]["compress"] == "lzw" and rec["kw"]["dtype"] == "uint8"
    assert all(np.array_equal(rec["bands"][i], rgb[..., i]) for i in range(3))
    import cv2
    assert np.array_equal(cv2.imread(str(tmp_path / "o" / "scene_wow.png"))[:, :, ::-1], rgb)
    # without georeferencing only the PNG is written and returned
    assert wow.write_image(rgb, tmp_path / "p" / "plain.tif", None, None, 4) == tmp_path / "p" / "plain.png"
