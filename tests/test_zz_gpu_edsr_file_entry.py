"""GPU: ``apply_super_resolution`` / ``process_sentinel2_sr`` (super_resolution.py:127-324, the /api/sr path) through the mirror
module, PNG in / PNG out.  EDSR parity itself is unpinned: this checks the file glue around ``upsample``.
Sorted last: added after the round's last hardware run."""
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_edsr_file_entry_points(ws, tmp_path, monkeypatch):
    import cv2

    from oracle import edsr_ref as E
    sr_mod = ws.app.super_resolution
    sd = E.random_init_state_dict(1, 16)
    monkeypatch.setattr(sr_mod, "get_model_dir", lambda: tmp_path)
    rgb = np.random.default_rng(5).integers(0, 256, (30, 44, 3), dtype=np.uint8)
    src = tmp_path / "in.png"
    cv2.imwrite(str(src), np.ascontiguousarray(rgb[:, :, ::-1]))
    with pytest.raises(FileNotFoundError, match="EDSR_x4.pth"):   # no converted weights yet: loud, no fallback
        sr_mod.apply_super_resolution(src, tmp_path / "o" / "x")
    torch.save(sd, tmp_path / "EDSR_x4.pth")

    out_path, meta = sr_mod.apply_super_resolution(src, tmp_path / "o" / "scene_sr.tif", scale=4, model_type="edsr", output_format="tif")
    assert out_path == tmp_path / "o" / "scene_sr.png"            # not georeferenced -> PNG, like the reference (:222-227)
    assert meta == {"input_file": str(src), "output_file": str(out_path), "scale": 4, "model": "edsr_x4", "original_size": [30, 44],
                    "output_size": [120, 176], "original_resolution_m": 10.0, "effective_resolution_m": 2.5}
    sr, _ = sr_mod.create_sr_model(4, "edsr", state_dict=sd)
    want_bgr = sr.upsample(np.ascontiguousarray(rgb[:, :, ::-1]))
    assert np.array_equal(cv2.imread(str(out_path)), want_bgr)

    res = sr_mod.process_sentinel2_sr(src, tmp_path / "job", scale=4)
    assert set(res) == {"timestamp", "input", "outputs", "sr_metadata"}
    assert res["outputs"]["sr_tif"] is None and res["outputs"]["sr_png"].endswith("in_sr_x4.png")
    assert json.load(open(tmp_path / "job" / "in_sr_metadata.json"))["sr_metadata"]["model"] == "edsr_x4"
    assert np.array_equal(cv2.imread(res["outputs"]["sr_png"]), want_bgr)
