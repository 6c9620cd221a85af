"""Drop-in for ``server/app/wow_sr.py``: ``_enhance_for_crops`` (:187-209), ``apply_wow_sr`` (:27-184) and
``process_wow_sr`` (:212-266).  File decoding / encoding stays on the host (cv2; rasterio when it is installed);
everything between the decoded array and the encoded file runs on the GPU without a host round trip:
min-max normalisation of non-uint8 rasters (:66-72), RGB<->BGR swaps (:85,:103), the network, the post-process."""
from __future__ import annotations

import json
from datetime import datetime
from pathlib import Path
from typing import Tuple

import numpy as np
import torch

from .. import _lib


def _handle(device=None):
    return _lib.default_handle(torch.cuda.current_device() if device is None else device)


def _to_host(t: torch.Tensor) -> np.ndarray:
    """Finished device image -> numpy through the handle's pinned staging ring (chunked copies on a copy stream); a host
    tensor (the CPU test stand-ins) is returned as is."""
    if t.is_cuda:
        return _handle(t.device.index).download(t.contiguous())
    return t.numpy()


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
    return _to_host(out)


# ---------------------------------------------------------------------------------------------
# IO glue of the /api/wow path
# ---------------------------------------------------------------------------------------------

def normalise_to_uint8_cuda(img: torch.Tensor, eps: float = 0.0) -> torch.Tensor:
    """The reference's raster normalisation (:66-72) on the device, bit-exact with numpy: rasters whose maximum exceeds 255
    are min-max stretched and truncated; everything else is cast (wrapping, like ``astype``).  The stretch runs in the type
    numpy computes it in: float32 rasters in float32, integer and float64 rasters in float64.  A constant raster above 255
    gives 0/0 = NaN, which numpy's ``astype(uint8)`` turns into 0 (a NaN -> uint8 cast is undefined on CUDA): handled
    explicitly.  ``eps`` is the ``+ 1e-6`` that ``apply_cnn_sr`` adds to the range (cnn_super_resolution.py:310)."""
    if img.dtype == torch.uint8:
        return img
    mx, mn = img.max(), img.min()
    if float(mx) > 255:
        ft = torch.float32 if img.dtype == torch.float32 else torch.float64
        rng = (mx - mn).to(ft) + eps
        if float(rng) == 0.0:
            return torch.zeros(img.shape, dtype=torch.uint8, device=img.device)
        return ((img - mn).to(ft) / rng * 255).to(torch.uint8)
    return img.to(torch.int64).to(torch.uint8)


def read_image(input_path: Path):
    """(RGB array of the file's dtype, transform, crs) as the reference reads it (:57-79)."""
    import cv2
    input_path = Path(input_path)
    if input_path.suffix.lower() in (".tif", ".tiff"):
        try:
            import rasterio
        except ImportError as e:  # same failure mode as the reference module, whose import of rasterio is unconditional
            raise ImportError("reading GeoTIFF needs rasterio (server/requirements.txt); PNG/JPEG inputs do not") from e
        with rasterio.open(input_path) as src:
            if src.count >= 3:
                img = np.stack([src.read(i) for i in [1, 2, 3]], axis=-1)
            else:
                band = src.read(1)
                img = np.stack([band, band, band], axis=-1)
            return img, src.transform, src.crs
    bgr = cv2.imread(str(input_path))
    if bgr is None:
        raise FileNotFoundError(str(input_path))
    return np.ascontiguousarray(bgr[:, :, ::-1]), None, None


def write_image(rgb: np.ndarray, output_path: Path, transform, crs, scale: int) -> Path:
    """GeoTIFF (when the input was georeferenced) plus PNG, as the reference writes them (:126-164)."""
    import cv2
    output_path = Path(output_path)
    output_path.parent.mkdir(parents=True, exist_ok=True)
    final_output = output_path.with_suffix(".png")
    if transform is not None:
        import rasterio
        from rasterio.transform import Affine
        new_transform = Affine(transform.a / scale, transform.b, transform.c, transform.d, transform.e / scale, transform.f)
        final_output = output_path.with_suffix(".tif")
        with rasterio.open(final_output, "w", driver="GTiff", height=rgb.shape[0], width=rgb.shape[1], count=3, dtype="uint8",
                           crs=crs, transform=new_transform, compress="lzw") as dst:
            for i in range(3):
                dst.write(rgb[:, :, i], i + 1)
    cv2.imwrite(str(output_path.with_suffix(".png")), np.ascontiguousarray(rgb[:, :, ::-1]))
    return final_output


def apply_wow_sr(input_path: Path, output_path: Path, enhance_crops: bool = True, model: str = "realesrgan_x4") -> Tuple[Path, dict]:
    """Same signature, outputs and metadata as the reference (:27-184)."""
    from .cnn_super_resolution import RealESRGAN
    input_path = Path(input_path)
    img, transform, crs = read_image(input_path)
    original_shape = img.shape[:2]
    esrgan = RealESRGAN(model_name=model, tile_size=256)  # resident after the first request (cnn_super_resolution._MODEL_CACHE)
    scale = esrgan.scale
    dev = esrgan.device
    host = np.ascontiguousarray(img)
    if host.dtype == np.uint16:  # torch has no arithmetic on uint16: widen on the host copy (values are preserved)
        host = host.astype(np.int32)
    x = normalise_to_uint8_cuda(torch.from_numpy(host).to(dev))
    sr_bgr = esrgan.enhance_cuda(x.flip(2).contiguous())          # RGB -> BGR (:85), enhance (:94)
    del esrgan
    out = sr_bgr.flip(2).contiguous()                             # BGR -> RGB (:103)
    pipeline_stages = [{"model": model, "scale": scale, "purpose": "GAN upscaling"}]
    if enhance_crops:
        out = enhance_for_crops_cuda(out)
        pipeline_stages.append({"post_processing": "Enhanced", "purpose": "Crop visibility"})
    output_rgb = _to_host(out)
    final_output = write_image(output_rgb, Path(output_path), transform, crs, scale)
    metadata = {
        "input_file": str(input_path),
        "output_file": str(final_output),
        "scale": scale,
        "pipeline": "Real-ESRGAN x4 + Enhanced",
        "stages": pipeline_stages,
        "enhancements": ["CLAHE local contrast", "Unsharp mask", "Vegetation boost"] if enhance_crops else [],
        "original_size": list(original_shape),
        "output_size": list(output_rgb.shape[:2]),
        "original_resolution_m": 10.0,
        "effective_resolution_m": 10.0 / scale,
        "optimized_for": "z18_crop_visibility",
    }
    return final_output, metadata


def process_wow_sr(input_tif: Path, output_dir: Path, enhance_crops: bool = True, model: str = "realesrgan_x4") -> dict:
    """Same result dictionary and side files as the reference (:212-266); this is what ``/api/wow`` returns."""
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    base_name = Path(input_tif).stem
    wow_tif = output_dir / f"{base_name}_wow_sr.tif"
    _, sr_metadata = apply_wow_sr(input_path=input_tif, output_path=wow_tif, enhance_crops=enhance_crops, model=model)
    result = {
        "timestamp": datetime.now().strftime("%Y%m%d_%H%M%S"),
        "input": str(input_tif),
        "outputs": {
            "sr_tif": str(wow_tif) if wow_tif.exists() else None,
            "sr_png": str(wow_tif.with_suffix(".png")) if wow_tif.with_suffix(".png").exists() else None,
        },
        "sr_metadata": sr_metadata,
    }
    with open(output_dir / f"{base_name}_wow_sr_metadata.json", "w") as f:
        json.dump(result, f, indent=2)
    return result
