"""Importable alias of the ``sentinel2-super-resolution-poc_b200`` package (its directory name is
not a Python identifier): ``import wowsr_b200 as ws; ws.app.cnn_super_resolution.RealESRGAN``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("sentinel2-super-resolution-poc_b200")
importlib.import_module("sentinel2-super-resolution-poc_b200.app.cnn_super_resolution")
importlib.import_module("sentinel2-super-resolution-poc_b200.app.wow_sr")
importlib.import_module("sentinel2-super-resolution-poc_b200.app.farm_sr")
importlib.import_module("sentinel2-super-resolution-poc_b200.app.vector_extraction")
importlib.import_module("sentinel2-super-resolution-poc_b200.app.super_resolution")
importlib.import_module("sentinel2-super-resolution-poc_b200.app.tiling")
sys.modules[__name__] = _pkg
