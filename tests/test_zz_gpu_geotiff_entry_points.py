"""GPU: the GeoTIFF branches of the file entry points (apply_wow_sr wow_sr.py:57-164, apply_farm_sr farm_sr.py:125-214,
apply_cnn_sr cnn_super_resolution.py:299-358, process_sentinel2_sr super_resolution.py:156-300) end to end: a 16-bit 4-band
raster is read, min-max stretched on the device exactly like the reference's numpy expression, super-resolved, written back
with the pixel size divided by the scale.  rasterio is not installed here: tests/fake_rasterio.py stands in for it.
Sorted last: added after the round's last hardware run."""
import numpy as np
import pytest
import torch

from oracle import rrdbnet_ref as R
from tests import fake_rasterio as FR

pytestmark = pytest.mark.gpu


def _stretch(img, eps=0.0):
    if img.max() > 255:
        return ((img - img.min()) / (img.max() - img.min() + eps) * 255).astype(np.uint8)
    return img.astype(np.uint8)


def test_geotiff_in_geotiff_out(ws, tmp_path, monkeypatch):
    from oracle import edsr_ref as E
    FR.install(monkeypatch)
    cnn, wow, farm, srm = ws.app.cnn_super_resolution, ws.app.wow_sr, ws.app.farm_sr, ws.app.super_resolution
    blocks = cnn.MODELS["realesrgan_x4"]["blocks"]
    sd = R.calibrate_conv_last(R.random_init_state_dict(6, blocks), blocks)
    esd = E.random_init_state_dict(2, 16)
    monkeypatch.setattr(cnn, "get_model_dir", lambda: tmp_path)
    monkeypatch.setattr(srm, "get_model_dir", lambda: tmp_path)
    torch.save({"params_ema": sd}, tmp_path / "realesrgan_x4.pth")
    torch.save(esd, tmp_path / "EDSR_x4.pth")
    cnn.clear_model_cache()
    rng = np.random.default_rng(17)
    bands = [rng.integers(200, 7000, (26, 38)).astype(np.uint16) for _ in range(4)]
    src = tmp_path / "s2.tif"
    FR.put(src, bands)
    raster = np.stack(bands[:3], axis=-1)
    up = cnn.RealESRGAN(scale=4, tile_size=256)
    want_t = FR.Affine(2.5, 0.0, 5e5, 0.0, -2.5, 4e6)

    def sr_rgb_of(rgb_u8):      # what every Real-ESRGAN entry point computes: RGB -> BGR -> enhance -> RGB
        return np.ascontiguousarray(up.enhance(np.ascontiguousarray(rgb_u8[:, :, ::-1]))[:, :, ::-1])

    def written(path):
        rec = FR.STORE[str(path)]
        assert rec["transform"] == want_t and rec["crs"] == "EPSG:32636" and rec["kw"]["dtype"] == "uint8"
        return np.stack(rec["bands"], axis=-1)

    # /api/wow
    out, meta = wow.apply_wow_sr(src, tmp_path / "wow" / "s2_wow.tif", enhance_crops=True)
    assert out.suffix == ".tif" and meta["original_size"] == [26, 38] and meta["output_size"] == [104, 152]
    assert np.array_equal(written(out), wow._enhance_for_crops(sr_rgb_of(_stretch(raster))))
    # farm SR
    out, meta = farm.apply_farm_sr(src, tmp_path / "farm" / "s2_farm.tif", scale=4)
    assert out.suffix == ".tif" and np.array_equal(written(out), farm.farm_post(sr_rgb_of(_stretch(raster))))
    # CLI entry point: its stretch carries the + 1e-6 (cnn_super_resolution.py:308-311)
    out, meta = cnn.apply_cnn_sr(src, tmp_path / "cli" / "s2_cnn.png", scale=4)
    assert out.suffix == ".tif" and meta["input_size"] == [38, 26] and meta["output_size"] == [152, 104]
    assert np.array_equal(written(out), sr_rgb_of(_stretch(raster, 1e-6)))
    # /api/sr (EDSR; parity of the network itself is unpinned, the glue is what is checked)
    res = srm.process_sentinel2_sr(src, tmp_path / "sr", scale=4)
    assert res["outputs"]["sr_tif"].endswith("s2_sr_x4.tif") and res["outputs"]["sr_png"].endswith("s2_sr_x4.png")
    sr, _ = srm.create_sr_model(4, "edsr", state_dict=esd)
    want = np.ascontiguousarray(sr.upsample(np.ascontiguousarray(_stretch(raster)[:, :, ::-1]))[:, :, ::-1])
    assert np.array_equal(written(res["outputs"]["sr_tif"]), want)
    import cv2
    assert np.array_equal(cv2.imread(res["outputs"]["sr_png"])[:, :, ::-1], want)
    cnn.clear_model_cache()
