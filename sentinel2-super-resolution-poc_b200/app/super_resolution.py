"""Drop-in for the EDSR x4 part of ``server/app/super_resolution.py`` (the "farm SR" variant BASELINE names).

The reference builds ``cv2.dnn_superres.DnnSuperResImpl`` from an external ``EDSR_x4.pb``
(:92-124) and calls ``sr.upsample(img_bgr)`` (:196).  This module keeps that surface —
``create_sr_model(scale, model_type) -> (sr, actual_scale)`` with ``sr.upsample(img_bgr)`` — over
libwowsr's EDSR-baseline network.  Parity is unpinned (see oracle/edsr_ref.py): the TensorFlow graph and the
contrib module are not available offline, so weights must be supplied as a state dict in the key order of
``oracle.edsr_ref.conv_specs`` (head, body.N.conv1/conv2, body_end, up1, up2, tail; OIHW).
"""
from __future__ import annotations

import json
from datetime import datetime
from pathlib import Path
from typing import Tuple

import numpy as np
import torch

from .. import _lib

SR_MODELS = {"edsr_x4": {"scale": 4, "description": "EDSR-baseline x4 (16 resblocks, 64 features)"}}


def get_model_dir() -> Path:
    """Same location as the reference (:62-66)."""
    model_dir = Path(__file__).parent.parent / "models"
    model_dir.mkdir(exist_ok=True)
    return model_dir


def edsr_keys(num_block=16):
    keys = ["head"]
    for b in range(num_block):
        keys += [f"body.{b}.conv1", f"body.{b}.conv2"]
    return keys + ["body_end", "up1", "up2", "tail"]


class EdsrSuperRes:
    """Object with the ``DnnSuperResImpl`` methods the reference uses: ``upsample``."""

    def __init__(self, state_dict, num_block=16, res_scale=1.0, precision="bf16", device=0, handle=None):
        self.num_block = num_block
        self._h = handle if handle is not None else _lib.Handle(device)
        tensors = [np.asarray(torch.as_tensor(state_dict[k + s]).cpu().numpy(), dtype=np.float32)
                   for k in edsr_keys(num_block) for s in (".weight", ".bias")]
        self._h.load_edsr(tensors, num_block, 64, res_scale, precision)

    def upsample(self, img_bgr: np.ndarray) -> np.ndarray:
        if img_bgr.ndim != 3 or img_bgr.shape[2] != 3:
            raise ValueError("expected an HxWx3 BGR uint8 image")
        return self._h.edsr_upsample_host(np.asarray(img_bgr).astype(np.uint8, copy=False))

    def upsample_float(self, img_bgr: np.ndarray):
        return self._h.edsr_upsample_host(np.asarray(img_bgr).astype(np.uint8, copy=False), want_float=True)


def create_sr_model(scale: int = 4, model_type: str = "edsr", *, state_dict=None, num_block=16, precision="bf16"):
    """Same return shape as the reference (:92-124): ``(sr, actual_scale)``."""
    name = f"{model_type}_x{scale}"
    if name not in SR_MODELS:
        raise ValueError(f"Unknown model: {name}. Available: {list(SR_MODELS.keys())}")
    if state_dict is None:
        # the reference keeps its models under <server>/models (:62-66); the converted weights live there as a torch state dict
        path = get_model_dir() / "EDSR_x4.pth"
        if not path.exists():
            raise FileNotFoundError(f"{path} is missing: EDSR weights are not bundled (the reference downloads a TensorFlow .pb); "
                                    "convert them to a state dict in the key order of edsr_keys() or pass state_dict=")
        state_dict = torch.load(path, map_location="cpu")
    if not torch.cuda.is_available():
        raise RuntimeError("this build runs on a B200 only (no CPU fallback)")
    return EdsrSuperRes(state_dict, num_block=num_block, precision=precision, device=torch.cuda.current_device()), 4


def apply_super_resolution(input_path: Path, output_path: Path, scale: int = 4, model_type: str = "edsr",
                           output_format: str = "tif") -> Tuple[Path, dict]:
    """Same signature, outputs and metadata as the reference's file entry point of the /api/sr path (:127-257): read
    (GeoTIFF bands 1-3 with the min-max stretch of :166-174, or any cv2-readable image), RGB->BGR, ``upsample``, BGR->RGB,
    GeoTIFF with the transform scaled when the input was georeferenced and ``output_format == "tif"``, PNG otherwise."""
    import cv2

    from . import wow_sr
    input_path = Path(input_path)
    img, transform, crs = wow_sr.read_image(input_path)                     # RGB, file dtype
    if img.dtype != np.uint8:
        host = img.astype(np.int32) if img.dtype == np.uint16 else img
        img = wow_sr.normalise_to_uint8_cuda(torch.from_numpy(np.ascontiguousarray(host)).cuda()).cpu().numpy()
    original_shape = img.shape[:2]
    sr_model, actual_scale = create_sr_model(scale=scale, model_type=model_type)
    output_bgr = sr_model.upsample(np.ascontiguousarray(img[:, :, ::-1]))
    output_rgb = np.ascontiguousarray(output_bgr[:, :, ::-1])
    output_path = Path(output_path)
    output_path.parent.mkdir(parents=True, exist_ok=True)
    if output_format == "tif" and transform is not None:
        import rasterio
        from rasterio.transform import Affine
        new_transform = Affine(transform.a / actual_scale, transform.b, transform.c, transform.d, transform.e / actual_scale,
                               transform.f)
        final_output = output_path.with_suffix(".tif")
        with rasterio.open(final_output, "w", driver="GTiff", height=output_rgb.shape[0], width=output_rgb.shape[1], count=3,
                           dtype="uint8", crs=crs, transform=new_transform, compress="lzw") as dst:
            for i in range(3):
                dst.write(output_rgb[:, :, i], i + 1)
    else:
        final_output = output_path.with_suffix(".png")
        cv2.imwrite(str(final_output), output_bgr)
    metadata = {
        "input_file": str(input_path),
        "output_file": str(final_output),
        "scale": actual_scale,
        "model": f"{model_type}_x{actual_scale}",
        "original_size": list(original_shape),
        "output_size": list(output_rgb.shape[:2]),
        "original_resolution_m": 10.0,
        "effective_resolution_m": 10.0 / actual_scale,
    }
    return final_output, metadata


def process_sentinel2_sr(input_tif: Path, output_dir: Path, scale: int = 4, model_type: str = "edsr") -> dict:
    """Same result dictionary and side files as the reference (:260-324); this is what ``/api/sr`` returns."""
    import cv2
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    timestamp = datetime.now().strftime("%Y%m%d_%H%M%S")
    base_name = Path(input_tif).stem
    sr_tif = output_dir / f"{base_name}_sr_x{scale}.tif"
    sr_png = output_dir / f"{base_name}_sr_x{scale}.png"
    output_path, sr_metadata = apply_super_resolution(input_path=input_tif, output_path=sr_tif, scale=scale, model_type=model_type,
                                                      output_format="tif")
    if output_path.suffix == ".tif":  # also a PNG copy of the GeoTIFF (:296-300)
        import rasterio
        with rasterio.open(output_path) as src:
            img = np.stack([src.read(i) for i in [1, 2, 3]], axis=-1)
        cv2.imwrite(str(sr_png), np.ascontiguousarray(img[:, :, ::-1]))
    result = {
        "timestamp": timestamp,
        "input": str(input_tif),
        "outputs": {"sr_tif": str(sr_tif) if sr_tif.exists() else None, "sr_png": str(sr_png) if sr_png.exists() else None},
        "sr_metadata": sr_metadata,
    }
    with open(output_dir / f"{base_name}_sr_metadata.json", "w") as f:
        json.dump(result, f, indent=2)
    return result
