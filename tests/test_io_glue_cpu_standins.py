"""CPU: the HOST logic of the file entry points (GeoTIFF and PNG branches, metadata, side files) with the device compute replaced
by oracle-backed stand-ins — the same test bodies that run against the real kernels under ``-m gpu``
(tests/test_zz_gpu_*_entry*.py) are executed here on CPU tensors, so the glue is exercised without a GPU.  Nothing in the
product imports these stand-ins; they exist only inside this test."""
import numpy as np
import pytest
import torch

from oracle import edsr_ref as E
from oracle import rrdbnet_ref as R
from oracle import wow_cv2


@pytest.fixture
def standins(ws, monkeypatch):
    cnn, wow, farm, srm = ws.app.cnn_super_resolution, ws.app.wow_sr, ws.app.farm_sr, ws.app.super_resolution

    class OracleESRGAN:
        """RealESRGAN surface over the fp32 oracle; 'device' is the CPU so that the glue's tensor traffic stays on the host."""

        def __init__(self, scale=4, device=None, tile_size=256, model_name=None, **kw):
            name = model_name or f"realesrgan_x{scale}"
            if name not in cnn.MODELS:
                raise ValueError(f"Unknown model: {name}")
            self.scale, self.tile_size, self.device, self.model_name = 4, tile_size, torch.device("cpu"), name
            sd = torch.load(cnn.get_model_dir() / f"{name}.pth", map_location="cpu")
            self.sd, self.blocks = sd.get("params_ema", sd.get("params", sd)), cnn.MODELS[name]["blocks"]

        def enhance(self, img):
            return R.enhance(self.sd, np.ascontiguousarray(img), self.blocks, self.tile_size)

        def enhance_cuda(self, x):
            return torch.from_numpy(self.enhance(x.numpy()))

    class OracleEdsr:
        def __init__(self, state_dict, num_block=16, **kw):
            self.sd, self.nb = state_dict, num_block

        def upsample(self, img):
            return E.quantise(E.forward_float(self.sd, np.ascontiguousarray(img), self.nb))

    monkeypatch.setattr(cnn, "RealESRGAN", OracleESRGAN)
    monkeypatch.setattr(cnn, "clear_model_cache", lambda: None)
    monkeypatch.setattr(wow, "enhance_for_crops_cuda", lambda t: torch.from_numpy(wow_cv2.enhance_for_crops(t.numpy())))
    monkeypatch.setattr(wow, "_enhance_for_crops", wow_cv2.enhance_for_crops)
    monkeypatch.setattr(farm, "farm_post_cuda", lambda t: torch.from_numpy(wow_cv2.farm_post(t.numpy())))
    monkeypatch.setattr(farm, "farm_post", wow_cv2.farm_post)
    monkeypatch.setattr(srm, "EdsrSuperRes", OracleEdsr)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    return ws


def test_geotiff_entry_points_host_logic(standins, tmp_path, monkeypatch):
    from tests import test_zz_gpu_geotiff_entry_points as T
    T.test_geotiff_in_geotiff_out(standins, tmp_path, monkeypatch)


def test_cnn_sr_file_entry_host_logic(standins, tmp_path, monkeypatch):
    from tests import test_zz_gpu_cnn_sr_file_entry as T
    T.test_apply_cnn_sr_png_in_png_out(standins, tmp_path, monkeypatch)


def test_edsr_file_entry_host_logic(standins, tmp_path, monkeypatch):
    from tests import test_zz_gpu_edsr_file_entry as T
    T.test_edsr_file_entry_points(standins, tmp_path, monkeypatch)
