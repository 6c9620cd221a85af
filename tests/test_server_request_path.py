"""The drop-in under the reference's request path.

* CPU (this container only — the reference tree cannot travel): the UNMODIFIED reference FastAPI app (server/app/main.py)
  with the drop-in's ``process_wow_sr`` patched into ``app.wow_sr``, driven through ``POST /api/wow`` and ``POST /api/enhance``
  with ``fastapi.testclient.TestClient``; the device calls of the drop-in are replaced by the CPU oracle (test infrastructure),
  so what is checked is the plumbing: job life cycle, result dictionary, side files, metadata keys (wow_sr.py:166-184,243-266).
* GPU: the same requests against tests/mini_wow_server.py (a restatement of those handlers) with the real drop-in on the B200;
  the written PNG must equal oracle(network) -> cv2 post-process within the network tolerance."""
import importlib
import json
import os
import sys

import cv2
import numpy as np
import pytest
import torch

from oracle import refload, rrdbnet_ref as R, wow_cv2

METADATA_KEYS = {"input_file", "output_file", "scale", "pipeline", "stages", "enhancements", "original_size", "output_size",
                 "original_resolution_m", "effective_resolution_m", "optimized_for"}          # wow_sr.py:166-184
RESULT_KEYS = {"timestamp", "input", "outputs", "sr_metadata"}                                # wow_sr.py:243-255


def _write_png(path, shape=(40, 52), seed=3):
    img = np.random.default_rng(seed).integers(0, 256, shape + (3,), dtype=np.uint8)
    cv2.imwrite(str(path), img)          # the file holds BGR = img
    return img


def _check_result(result, in_path, shape, enhance=True):
    assert set(result) == RESULT_KEYS and set(result["outputs"]) == {"sr_tif", "sr_png"}
    assert result["outputs"]["sr_tif"] is None and result["outputs"]["sr_png"].endswith("_wow_sr.png")
    md = result["sr_metadata"]
    assert set(md) == METADATA_KEYS
    assert md["scale"] == 4 and md["pipeline"] == "Real-ESRGAN x4 + Enhanced" and md["optimized_for"] == "z18_crop_visibility"
    assert md["original_size"] == list(shape) and md["output_size"] == [4 * shape[0], 4 * shape[1]]
    assert md["effective_resolution_m"] == 2.5 and md["input_file"] == str(in_path)
    assert md["enhancements"] == (["CLAHE local contrast", "Unsharp mask", "Vegetation boost"] if enhance else [])
    side = os.path.join(os.path.dirname(result["outputs"]["sr_png"]), os.path.basename(str(in_path)).rsplit(".", 1)[0] + "_wow_sr_metadata.json")
    assert json.load(open(side))["sr_metadata"]["output_size"] == md["output_size"]
    out = cv2.imread(result["outputs"]["sr_png"])
    assert out.shape == (4 * shape[0], 4 * shape[1], 3)
    return out[:, :, ::-1]               # RGB


class _OracleESRGAN:
    """Stands in for the drop-in's RealESRGAN in the CPU test: same attributes, oracle arithmetic, CPU tensors."""
    sd = {}

    def __init__(self, scale=4, device=None, tile_size=256, model_name=None, **kw):
        self.scale, self.device, self.tile_size, self.model_name = 4, torch.device("cpu"), tile_size, model_name or "realesrgan_x4"
        self.blocks = 6 if self.model_name == "realesrgan_anime" else 2     # small stand-ins: this test is about plumbing
        if self.blocks not in self.sd:
            self.sd[self.blocks] = R.random_init_state_dict(0, self.blocks)

    def enhance_cuda(self, x):
        return torch.from_numpy(R.enhance(self.sd[self.blocks], x.numpy(), self.blocks, tile_size=self.tile_size))


@pytest.mark.skipif(not refload.available(), reason="the reference tree is only present in the build container")
def test_reference_fastapi_app_with_dropin_patched_in(tmp_path, monkeypatch):
    from fastapi.testclient import TestClient
    import wowsr_b200 as ws
    monkeypatch.setenv("MAPBOX_ACCESS_TOKEN", "test")
    monkeypatch.setenv("DATA_DIR", str(tmp_path / "data"))
    refload._stub_rasterio()
    if refload.REF_SERVER not in sys.path:
        sys.path.insert(0, refload.REF_SERVER)
    for name in [m for m in sys.modules if m == "app" or m.startswith("app.")]:   # a fresh reference package (settings are cached)
        monkeypatch.delitem(sys.modules, name)
    main = importlib.import_module("app.main")
    ref_wow = importlib.import_module("app.wow_sr")
    drop = ws.app.wow_sr
    # device calls of the drop-in -> CPU oracle
    monkeypatch.setattr(ws.app.cnn_super_resolution, "RealESRGAN", _OracleESRGAN)
    monkeypatch.setattr(drop, "enhance_for_crops_cuda", lambda t: torch.from_numpy(wow_cv2.enhance_for_crops(t.numpy())))
    # the maintainer's one-line swap: app.wow_sr.process_wow_sr is the drop-in's
    monkeypatch.setattr(ref_wow, "process_wow_sr", drop.process_wow_sr)
    client = TestClient(main.app)
    src = tmp_path / "field.png"
    img = _write_png(src)
    r = client.post("/api/wow", json={"input_file": str(src), "auto_fetch": False})
    assert r.status_code == 200 and r.json()["status"] == "queued" and r.json()["job_id"].startswith("wow_")
    job = client.get(f"/api/sr/{r.json()['job_id']}").json()           # TestClient runs background tasks before returning
    assert job["status"] == "completed", job
    rgb = _check_result(job["result"], src, img.shape[:2])
    sr = R.enhance(_OracleESRGAN.sd[2], np.ascontiguousarray(img), 2, tile_size=256)     # file BGR in, BGR out
    assert np.array_equal(rgb, wow_cv2.enhance_for_crops(np.ascontiguousarray(sr[:, :, ::-1])))
    assert client.post("/api/wow", json={"input_file": str(tmp_path / "missing.png")}).status_code == 404
    # uploaded photo path (what the Angular page uses), anime model
    with open(src, "rb") as f:
        r = client.post("/api/enhance", files={"image": ("field.png", f, "image/png")}, data={"model": "realesrgan_anime"})
    assert r.status_code == 200 and r.json()["model"] == "realesrgan_anime"
    job = client.get(f"/api/sr/{r.json()['job_id']}").json()
    assert job["status"] == "completed", job
    assert set(job["result"]["sr_metadata"]) == METADATA_KEYS and job["result"]["sr_metadata"]["stages"][0]["model"] == "realesrgan_anime"
    with open(src, "rb") as f:
        assert client.post("/api/enhance", files={"image": ("field.png", f, "image/png")}, data={"model": "nope"}).status_code == 400


@pytest.mark.gpu
def test_request_path_on_the_gpu(ws, tmp_path, monkeypatch):
    from fastapi.testclient import TestClient
    from tests.mini_wow_server import build_app
    cnn = ws.app.cnn_super_resolution
    sds = {"realesrgan_x4": R.calibrate_conv_last(R.random_init_state_dict(0, 23), 23),
           "realesrgan_anime": R.calibrate_conv_last(R.random_init_state_dict(1, 6), 6)}
    real = cnn.RealESRGAN

    class WithWeights(real):          # the build box has no network for download_weights: supply the state dict
        def __init__(self, scale=4, device=None, tile_size=256, model_name=None, **kw):
            name = model_name or f"realesrgan_x{scale}"
            super().__init__(scale=scale, device=device, tile_size=tile_size, model_name=model_name, state_dict=sds[name], **kw)

    monkeypatch.setattr(cnn, "RealESRGAN", WithWeights)
    client = TestClient(build_app(ws.app.wow_sr, tmp_path / "data"))
    src = tmp_path / "field.png"
    img = _write_png(src, shape=(48, 64))
    r = client.post("/api/wow", json={"input_file": str(src), "auto_fetch": False})
    assert r.status_code == 200 and r.json()["status"] == "queued"
    job = client.get(f"/api/sr/{r.json()['job_id']}").json()
    assert job["status"] == "completed", job
    rgb = _check_result(job["result"], src, img.shape[:2])
    sr_ref = R.enhance(sds["realesrgan_x4"], np.ascontiguousarray(img), 23, tile_size=256)
    up = real(device="cuda", tile_size=256, state_dict=sds["realesrgan_x4"])
    sr = up.enhance(np.ascontiguousarray(img))
    assert (np.abs(sr.astype(int) - sr_ref.astype(int)) <= 1).mean() >= 0.999
    assert np.array_equal(rgb, wow_cv2.enhance_for_crops(np.ascontiguousarray(sr[:, :, ::-1])))
    with open(src, "rb") as f:
        r = client.post("/api/enhance", files={"image": ("field.png", f, "image/png")}, data={"model": "realesrgan_anime"})
    job = client.get(f"/api/sr/{r.json()['job_id']}").json()
    assert job["status"] == "completed" and job["result"]["sr_metadata"]["stages"][0]["model"] == "realesrgan_anime", job
    assert client.post("/api/wow", json={"input_file": str(tmp_path / "missing.png")}).status_code == 404
    with open(src, "rb") as f:
        assert client.post("/api/enhance", files={"image": ("f.png", f, "image/png")}, data={"model": "nope"}).status_code == 400
    # a failing job is reported, not raised (main.py:365-368)
    bad = tmp_path / "broken.png"
    bad.write_bytes(b"not an image")
    r = client.post("/api/wow", json={"input_file": str(bad)})
    assert client.get(f"/api/sr/{r.json()['job_id']}").json()["status"] == "failed"
