"""B200-native WOW super-resolution hot path (drop-in for the reference's upsampler surface).

    from <this package>.app.cnn_super_resolution import RealESRGAN      # cnn_super_resolution.py:161-280
    from <this package>.app.wow_sr import _enhance_for_crops             # wow_sr.py:187-209
    from <this package>.app.farm_sr import enhance_local_contrast, ...   # farm_sr.py:61-108

All arithmetic runs in libwowsr.so (hand-written sm_100a CUDA, include/wowsr.h); there is no CPU
fallback.  The directory name is not a valid Python identifier, so import it through
``wowsr_b200`` (repo root) or ``importlib.import_module``.
"""
from . import _lib  # noqa: F401
from ._lib import Handle, WowsrError, default_handle, plan_windows  # noqa: F401
