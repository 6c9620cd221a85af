"""Drop-in for the HSV vegetation mask of ``server/app/vector_extraction.py``: ``ExtractionConfig``'s HSV
fields (:50-59) and ``compute_green_mask_hsv`` (:222-270).  Only this function of that module is on the widened hot path —
polygonisation, NDVI and topology clean-up stay where they are.  The mask is computed by ``wowsr_green_mask``
(csrc/post.cu) with the same RGB->HSV arithmetic as the vegetation boost; there is no CPU fallback."""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Tuple

import numpy as np
import torch

from .wow_sr import _handle, read_image

BROWN_RANGE = ((10, 20, 40), (35, 200, 200))  # "brownish vegetation (dry crops)", vector_extraction.py:262-264


@dataclass
class ExtractionConfig:
    """The fields ``compute_green_mask_hsv`` reads (vector_extraction.py:57-59), same names and defaults."""
    hsv_green_hue_range: Tuple[int, int] = (35, 85)
    hsv_saturation_min: int = 30
    hsv_value_min: int = 30


def hsv_ranges(config) -> list:
    """The two ``cv2.inRange`` boxes of the reference (:255-264): green from the config, brown fixed."""
    hue_min, hue_max = config.hsv_green_hue_range
    return [((hue_min, config.hsv_saturation_min, config.hsv_value_min), (hue_max, 255, 255)), BROWN_RANGE]


def normalise_rgb(rgb: np.ndarray) -> np.ndarray:
    """:245-249 — rasters whose maximum exceeds 255 are scaled by their maximum (float64, truncating), others are cast."""
    if rgb.max() > 255:
        return (rgb / rgb.max() * 255).astype(np.uint8)
    return rgb.astype(np.uint8)


def green_mask_hsv_array(rgb: np.ndarray, config=None) -> np.ndarray:
    """RGB uint8 HxWx3 -> float32 HxW mask of 0 / 1 (the part of the reference function after the raster is read)."""
    return _handle().green_mask_host(rgb, hsv_ranges(config or ExtractionConfig()))


def green_mask_hsv_cuda(rgb: torch.Tensor, config=None) -> torch.Tensor:
    """Device-resident variant: uint8 HxWx3 CUDA tensor -> float32 HxW CUDA tensor (e.g. straight from the SR output)."""
    assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous() and rgb.dim() == 3 and rgb.shape[2] == 3
    H, W = rgb.shape[:2]
    mask = torch.empty((H, W), dtype=torch.float32, device=rgb.device)
    _handle(rgb.device.index).green_mask_dev(rgb.data_ptr(), H, W, hsv_ranges(config or ExtractionConfig()), mask.data_ptr(),
                                             stream=torch.cuda.current_stream(rgb.device).cuda_stream)
    return mask


def compute_green_mask_hsv(raster_path: Path, config=None) -> np.ndarray:
    """Same contract as the reference (:222-270): path of an RGB raster -> float32 mask (0 / 1)."""
    img, _, _ = read_image(Path(raster_path))
    return green_mask_hsv_array(normalise_rgb(np.ascontiguousarray(img[:, :, :3])), config)


__all__ = ["ExtractionConfig", "compute_green_mask_hsv", "green_mask_hsv_array", "green_mask_hsv_cuda", "hsv_ranges", "normalise_rgb"]
