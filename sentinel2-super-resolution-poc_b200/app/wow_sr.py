"""Drop-in for the hot-path parts of ``server/app/wow_sr.py``: ``_enhance_for_crops`` (:187-209)
and the in-memory core of ``apply_wow_sr`` (:85-113)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def _handle(device=None):
    return _lib.default_handle(torch.cuda.current_device() if device is None else device)


def _enhance_for_crops(img: np.ndarray) -> np.ndarray:
    """CLAHE(2.5, 8x8) on L -> unsharp (sigma 1.2, 1.4/-0.4) -> green saturation x1.2; RGB uint8."""
    return _handle().post_process_host(img, _lib.post_params("wow"))


def enhance_for_crops_cuda(img: torch.Tensor) -> torch.Tensor:
    """Device-resident ``_enhance_for_crops``: uint8 HxWx3 CUDA tensor in/out."""
    assert img.is_cuda and img.dtype == torch.uint8 and img.is_contiguous()
    H, W = img.shape[:2]
    out = torch.empty_like(img)
    _handle(img.device.index).post_process_dev(img.data_ptr(), out.data_ptr(), H, W, _lib.post_params("wow"),
                                               stream=torch.cuda.current_stream(img.device).cuda_stream)
    return out


def wow_sr_array(img_rgb: np.ndarray, upsampler, enhance_crops: bool = True) -> np.ndarray:
    """The compute core of ``apply_wow_sr`` (:85-113) on an in-memory RGB uint8 image:
    RGB->BGR (:85), ``enhance`` (:94), BGR->RGB (:103), ``_enhance_for_crops`` (:110).
    The intermediate SR image stays on the GPU."""
    dev = upsampler.device
    x = torch.from_numpy(np.ascontiguousarray(img_rgb[:, :, ::-1])).to(dev)
    sr_bgr = upsampler.enhance_cuda(x)
    sr_rgb = sr_bgr.flip(2).contiguous()
    out = enhance_for_crops_cuda(sr_rgb) if enhance_crops else sr_rgb
    return out.cpu().numpy()
