"""TEST INFRASTRUCTURE ONLY — an in-memory stand-in for the few ``rasterio`` calls the GeoTIFF branches of the file entry points
make (``open(path)`` -> ``count`` / ``read(i)`` / ``transform`` / ``crs``; ``open(path, "w", ...)`` -> ``write(band, i)``;
``rasterio.transform.Affine``).  rasterio is not installed in this image, so without it those branches never run before
production.  ``install(monkeypatch)`` puts the fake into ``sys.modules``; "files" live in ``STORE`` keyed by path, and a written
dataset also leaves an empty file on disk so that ``Path.exists()`` checks of the callers see it."""
import sys
import types
from pathlib import Path

import numpy as np

STORE = {}


class Affine:
    def __init__(self, a, b, c, d, e, f):
        self.a, self.b, self.c, self.d, self.e, self.f = a, b, c, d, e, f

    def __eq__(self, o):
        return (self.a, self.b, self.c, self.d, self.e, self.f) == (o.a, o.b, o.c, o.d, o.e, o.f)

    def __repr__(self):
        return f"Affine({self.a}, {self.b}, {self.c}, {self.d}, {self.e}, {self.f})"


class _Dataset:
    def __init__(self, path, mode="r", **kw):
        self.path, self.mode = str(path), mode
        if mode == "r":
            rec = STORE[self.path]
            self.bands, self.transform, self.crs = rec["bands"], rec["transform"], rec["crs"]
            self.count = len(self.bands)
        else:
            self.kw = kw
            self.bands = [None] * kw["count"]
            self.count = kw["count"]

    def read(self, i):
        return self.bands[i - 1]

    def write(self, band, i):
        assert band.shape == (self.kw["height"], self.kw["width"]) and str(band.dtype) == self.kw["dtype"]
        self.bands[i - 1] = np.array(band)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        if self.mode == "w":
            assert all(b is not None for b in self.bands)
            STORE[self.path] = {"bands": self.bands, "transform": self.kw["transform"], "crs": self.kw["crs"], "kw": self.kw}
            Path(self.path).touch()
        return False


def put(path, bands, transform=None, crs="EPSG:32636"):
    """Registers an in-memory raster (list of HxW band arrays) under `path`."""
    STORE[str(path)] = {"bands": [np.asarray(b) for b in bands], "transform": transform or Affine(10.0, 0.0, 5e5, 0.0, -10.0, 4e6),
                        "crs": crs}
    Path(path).touch()


def install(monkeypatch):
    STORE.clear()
    r = types.ModuleType("rasterio")
    rt = types.ModuleType("rasterio.transform")
    rt.Affine = Affine
    r.transform = rt
    r.open = _Dataset
    monkeypatch.setitem(sys.modules, "rasterio", r)
    monkeypatch.setitem(sys.modules, "rasterio.transform", rt)
    return r
