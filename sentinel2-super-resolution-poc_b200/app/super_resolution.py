"""Drop-in for the EDSR x4 part of ``server/app/super_resolution.py`` (the "farm SR" variant BASELINE names).

The reference builds ``cv2.dnn_superres.DnnSuperResImpl`` from an external ``EDSR_x4.pb``
(:92-124) and calls ``sr.upsample(img_bgr)`` (:196).  This module keeps that surface —
``create_sr_model(scale, model_type) -> (sr, actual_scale)`` with ``sr.upsample(img_bgr)`` — over
libwowsr's EDSR-baseline network.  Parity is unpinned (see oracle/edsr_ref.py): the TensorFlow graph and the
contrib module are not available offline, so weights must be supplied as a state dict in the key order of
``oracle.edsr_ref.conv_specs`` (head, body.N.conv1/conv2, body_end, up1, up2, tail; OIHW).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib

SR_MODELS = {"edsr_x4": {"scale": 4, "description": "EDSR-baseline x4 (16 resblocks, 64 features)"}}


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
        raise FileNotFoundError("EDSR weights are not bundled (the reference downloads a TensorFlow .pb); pass state_dict=")
    if not torch.cuda.is_available():
        raise RuntimeError("this build runs on a B200 only (no CPU fallback)")
    return EdsrSuperRes(state_dict, num_block=num_block, precision=precision, device=torch.cuda.current_device()), 4
