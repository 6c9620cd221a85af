"""Drop-in for the hot-path parts of ``server/app/farm_sr.py``: ``apply_unsharp_mask`` (:61-71),
``enhance_local_contrast`` (:74-88), ``enhance_vegetation`` (:91-108) and the fused sequence
``apply_farm_sr`` runs at :170-178."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .wow_sr import _handle


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
