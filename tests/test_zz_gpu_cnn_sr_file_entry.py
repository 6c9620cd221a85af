"""GPU: ``apply_cnn_sr`` (cnn_super_resolution.py:283-375, the file entry point behind sr_cli.py:115-124) through the mirror
module.  Sorted last: added after the round's last hardware run."""
import numpy as np
import pytest
import torch

from oracle import rrdbnet_ref as R

pytestmark = pytest.mark.gpu


def test_apply_cnn_sr_png_in_png_out(ws, tmp_path, monkeypatch):
    import cv2
    cnn = ws.app.cnn_super_resolution
    blocks = cnn.MODELS["realesrgan_x4"]["blocks"]
    sd = R.calibrate_conv_last(R.random_init_state_dict(3, blocks), blocks)
    monkeypatch.setattr(cnn, "get_model_dir", lambda: tmp_path)
    torch.save({"params": sd}, tmp_path / "realesrgan_x4.pth")
    cnn.clear_model_cache()
    bgr = np.random.default_rng(21).integers(0, 256, (33, 47, 3), dtype=np.uint8)
    src = tmp_path / "in.png"
    cv2.imwrite(str(src), bgr)
    out_path, meta = cnn.apply_cnn_sr(src, tmp_path / "o" / "result.tif", scale=4)
    assert out_path == tmp_path / "o" / "result.png" and out_path.exists()
    assert meta == {"model": "RealESRGAN_x4", "scale": 4, "input_size": [47, 33], "output_size": [188, 132],
                    "device": str(cnn.RealESRGAN(scale=4, tile_size=256).device), "original_resolution_m": 10.0,
                    "effective_resolution_m": 2.5}
    # the file holds exactly enhance(cv2.imread(input)) — the reference passes the BGR array straight through (:316-333)
    want = cnn.RealESRGAN(scale=4, tile_size=256).enhance(bgr)
    assert np.array_equal(cv2.imread(str(out_path)), want)
    ref = R.enhance(sd, bgr, blocks, 256)
    assert (np.abs(want.astype(int) - ref.astype(int)) <= 1).mean() >= 0.999
    with pytest.raises(FileNotFoundError):
        cnn.apply_cnn_sr(tmp_path / "missing.png", tmp_path / "o" / "x")
    cnn.clear_model_cache()
