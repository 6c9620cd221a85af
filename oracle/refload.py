"""TEST INFRASTRUCTURE ONLY — loader for the unmodified reference modules (build container only).

/root/reference does not exist on the GPU box, so nothing under ``-m gpu``, ``smoke()`` or
``bench.py`` may call this.  It is used by ``tests/golden/make_golden.py`` (fixture generator) and
by the CPU tests that pin the oracle against the reference when the tree is present.

``app.wow_sr`` / ``app.farm_sr`` import ``rasterio`` at module top (wow_sr.py:18-19,
farm_sr.py:8-9), which is not installed; a stub module is injected so they import unchanged.
``RealESRGAN.__init__`` needs the network (weights download, cnn_super_resolution.py:55-70), so
instances are built with ``__new__`` + the attributes ``enhance``/``_tile_process`` read.
"""
from __future__ import annotations

import os
import sys
import types

REF_SERVER = "/root/reference/server"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SERVER, "app"))


def _stub_rasterio():
    if "rasterio" in sys.modules:
        return
    r = types.ModuleType("rasterio")
    rt = types.ModuleType("rasterio.transform")
    rt.Affine = object
    r.transform = rt
    sys.modules["rasterio"] = r
    sys.modules["rasterio.transform"] = rt


def load():
    """Returns (cnn_super_resolution, wow_sr, farm_sr) reference modules."""
    if not available():
        raise RuntimeError("reference tree not present")
    _stub_rasterio()
    if REF_SERVER not in sys.path:
        sys.path.insert(0, REF_SERVER)
    import importlib
    cnn = importlib.import_module("app.cnn_super_resolution")
    wow = importlib.import_module("app.wow_sr")
    farm = importlib.import_module("app.farm_sr")
    return cnn, wow, farm


def load_vector_extraction(read_bands):
    """The unmodified ``app.vector_extraction`` with ``rasterio.open(path)`` served by ``read_bands(path) -> [band arrays]``
    (rasterio and its mask / features / warp submodules are not installed; only ``open(...).read(i)`` is exercised by
    compute_green_mask_hsv, vector_extraction.py:237-243)."""
    if not available():
        raise RuntimeError("reference tree not present")
    _stub_rasterio()
    r = sys.modules["rasterio"]
    for name, attrs in (("mask", ("mask",)), ("features", ("shapes",)), ("warp", ("calculate_default_transform", "reproject", "Resampling"))):
        m = types.ModuleType("rasterio." + name)
        for a in attrs:
            setattr(m, a, object)
        setattr(r, name, m)
        sys.modules["rasterio." + name] = m

    class _Dataset:
        def __init__(self, path):
            self.bands = read_bands(path)
            self.count = len(self.bands)

        def read(self, i):
            return self.bands[i - 1]

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    r.open = _Dataset
    if REF_SERVER not in sys.path:
        sys.path.insert(0, REF_SERVER)
    import importlib
    return importlib.import_module("app.vector_extraction")


def make_upsampler(cnn, model, tile_size=256, scale=4):
    """A reference RealESRGAN wrapper around ``model`` without running its network-bound __init__."""
    import torch
    up = cnn.RealESRGAN.__new__(cnn.RealESRGAN)
    up.tile_size = tile_size
    up.tile_pad = 10
    up.device = torch.device("cpu")
    up.scale = scale
    up.model_name = "realesrgan_x4"
    up.model = model.eval()
    return up
