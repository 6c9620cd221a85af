"""Drop-in for the hot-path parts of ``server/app/farm_sr.py``: ``apply_unsharp_mask`` (:61-71),
``enhance_local_contrast`` (:74-88), ``enhance_vegetation`` (:91-108) and the fused sequence
``apply_farm_sr`` runs at :170-178."""
from __future__ import annotations

import numpy as np
import torch

import json
from datetime import datetime
from pathlib import Path
from typing import Tuple

from .. import _lib
from .wow_sr import _handle, _to_host, normalise_to_uint8_cuda, read_image, write_image


def apply_unsharp_mask(img: np.ndarray, strength: float = 1.5, radius: float = 1.0) -> np.ndarray:
    p = _lib.post_params("farm", stages=_lib.STAGE_UNSHARP, sigma=float(radius), alpha=1.0 + strength, beta=-strength)
    return _handle().post_process_host(img, p)


def enhance_local_contrast(img: np.ndarray, clip_limit: float = 3.0, grid_size: int = 8) -> np.ndarray:
    p = _lib.post_params("farm", stages=_lib.STAGE_CLAHE, clip_limit=float(clip_limit), grid=int(grid_size))
    return _handle().post_process_host(img, p)


def enhance_vegetation(img: np.ndarray) -> np.ndarray:
    p = _lib.post_params("farm", stages=_lib.STAGE_VEG, sat_boost=1.3)
    return _handle().post_process_host(img, p)


def farm_post(img: np.ndarray) -> np.ndarray:
    """Steps 2-4 of ``apply_farm_sr`` with the call-site constants (:170-178), one fused pass."""
    return _handle().post_process_host(img, _lib.post_params("farm"))


def farm_post_cuda(img: torch.Tensor) -> torch.Tensor:
    assert img.is_cuda and img.dtype == torch.uint8 and img.is_contiguous()
    H, W = img.shape[:2]
    out = torch.empty_like(img)
    _handle(img.device.index).post_process_dev(img.data_ptr(), out.data_ptr(), H, W, _lib.post_params("farm"),
                                               stream=torch.cuda.current_stream(img.device).cuda_stream)
    return out


def apply_farm_sr(input_path: Path, output_path: Path, scale: int = 4) -> Tuple[Path, dict]:
    """Same signature and outputs as the reference (:110-240): Real-ESRGAN x``scale`` (:161-165), then CLAHE(2.5, 8),
    unsharp(1.2, 1.5), vegetation boost (:170-178) as one fused device pass; the SR image never leaves the GPU."""
    from .cnn_super_resolution import RealESRGAN
    input_path = Path(input_path)
    img, transform, crs = read_image(input_path)
    original_shape = img.shape[:2]
    esrgan = RealESRGAN(scale=scale, tile_size=256)
    host = np.ascontiguousarray(img)
    if host.dtype == np.uint16:
        host = host.astype(np.int32)
    x = normalise_to_uint8_cuda(torch.from_numpy(host).to(esrgan.device))
    sr_rgb = esrgan.enhance_cuda(x.flip(2).contiguous()).flip(2).contiguous()
    del esrgan
    final = _to_host(farm_post_cuda(sr_rgb))
    final_output = write_image(final, Path(output_path), transform, crs, scale)
    metadata = {
        "input_file": str(input_path),
        "output_file": str(final_output),
        "scale": scale,
        "model": f"RealESRGAN_farm_x{scale}",
        "enhancements": ["Real-ESRGAN super-resolution", "CLAHE local contrast", "Unsharp mask edge sharpening",
                         "Vegetation enhancement"],
        "original_size": list(original_shape),
        "output_size": list(final.shape[:2]),
        "original_resolution_m": 10.0,
        "optimized_for": "crop_row_visibility",
    }
    return final_output, metadata


def process_farm_sr(input_tif: Path, output_dir: Path, scale: int = 4) -> dict:
    """Same result dictionary and side files as the reference (:243-285); this is what ``/api/sr`` returns."""
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    base_name = Path(input_tif).stem
    sr_tif = output_dir / f"{base_name}_farm_sr_x{scale}.tif"
    _, sr_metadata = apply_farm_sr(input_path=input_tif, output_path=sr_tif, scale=scale)
    result = {
        "timestamp": datetime.now().strftime("%Y%m%d_%H%M%S"),
        "input": str(input_tif),
        "outputs": {
            "sr_tif": str(sr_tif) if sr_tif.exists() else None,
            "sr_png": str(sr_tif.with_suffix(".png")) if sr_tif.with_suffix(".png").exists() else None,
        },
        "sr_metadata": sr_metadata,
    }
    with open(output_dir / f"{base_name}_farm_sr_metadata.json", "w") as f:
        json.dump(result, f, indent=2)
    return result
