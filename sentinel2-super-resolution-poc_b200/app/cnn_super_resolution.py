"""Drop-in for ``server/app/cnn_super_resolution.py`` (hot-path parts only).

Same public surface as the reference: ``MODELS`` (:28-45), ``get_model_dir`` / ``download_weights``
(:48-70), ``RRDBNet`` (:110-158, here a parameter container with the official state_dict keys) and
``RealESRGAN(scale, device, tile_size, model_name).enhance(img)`` (:161-280).  The forward pass, the
tile planner and the stitching run in libwowsr.so on the GPU.
"""
from __future__ import annotations

import urllib.request
from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from .. import _lib

MODELS = {
    "realesrgan_x4": {
        "url": "https://github.com/xinntao/Real-ESRGAN/releases/download/v0.1.0/RealESRGAN_x4plus.pth",
        "scale": 4, "channels": 64, "blocks": 23, "num_in_ch": 3,
        "description": "General photos (best quality)",
    },
    "realesrgan_anime": {
        "url": "https://github.com/xinntao/Real-ESRGAN/releases/download/v0.2.2.4/RealESRGAN_x4plus_anime_6B.pth",
        "scale": 4, "channels": 64, "blocks": 6, "num_in_ch": 3,
        "description": "Sharp edges (best for text/plates)",
    },
}


def get_model_dir() -> Path:
    model_dir = Path(__file__).parent.parent / "models"
    model_dir.mkdir(exist_ok=True)
    return model_dir


def download_weights(model_name: str) -> Path:
    """Same contract as the reference (:55-70): returns the cached .pth, downloading it if absent."""
    if model_name not in MODELS:
        raise ValueError(f"Unknown model: {model_name}")
    weights_path = get_model_dir() / f"{model_name}.pth"
    if not weights_path.exists():
        urllib.request.urlretrieve(MODELS[model_name]["url"], weights_path)
    return weights_path


def state_dict_keys(num_block: int):
    """Official Real-ESRGAN key order == the reference's construction order (:122-136)."""
    keys = ["conv_first"]
    for b in range(num_block):
        for r in (1, 2, 3):
            for k in range(1, 6):
                keys.append(f"body.{b}.rdb{r}.conv{k}")
    keys += ["conv_body", "conv_up1", "conv_up2", "conv_hr", "conv_last"]
    return keys


class RRDBNet(nn.Module):
    """Parameter container with the reference's state_dict layout (:110-138).

    ``load_state_dict(strict=True)`` accepts the official ``params_ema`` / ``params`` dictionaries.
    The arithmetic of ``forward`` lives in libwowsr.so and is reached through ``RealESRGAN``.
    """

    def __init__(self, num_in_ch=3, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32, scale=4):
        super().__init__()
        if (num_in_ch, num_out_ch, num_feat, num_grow_ch, scale) != (3, 3, 64, 32, 4):
            raise ValueError("the B200 build supports the x4 RRDBNet family only (3->3, nf=64, gc=32)")
        self.scale = scale
        self.num_block = num_block
        self._names = state_dict_keys(num_block)
        # same creation order as the reference => same default-init RNG stream under manual_seed
        for name in self._names:
            cin, cout = self._shape(name, num_feat, num_grow_ch, num_in_ch, num_out_ch)
            conv = nn.Conv2d(cin, cout, 3, 1, 1)
            self.register_parameter(name.replace(".", "__") + "__weight", conv.weight)
            self.register_parameter(name.replace(".", "__") + "__bias", conv.bias)

    @staticmethod
    def _shape(name, nf, gc, cin0, cout0):
        if name == "conv_first":
            return cin0, nf
        if name == "conv_last":
            return nf, cout0
        if name.startswith("body."):
            k = int(name[-1])
            return nf + (k - 1) * gc, gc if k < 5 else nf
        return nf, nf

    # official key names <-> registered parameter names
    def state_dict(self, *a, **kw):
        return OrderedDict((n + s, getattr(self, n.replace(".", "__") + "__" + s[1:]).detach())
                           for n in self._names for s in (".weight", ".bias"))

    def load_state_dict(self, state_dict, strict=True):
        want = [n + s for n in self._names for s in (".weight", ".bias")]
        missing = [k for k in want if k not in state_dict]
        extra = [k for k in state_dict if k not in want]
        if strict and (missing or extra):
            raise RuntimeError(f"Error(s) in loading state_dict for RRDBNet: missing {missing[:4]} unexpected {extra[:4]}")
        with torch.no_grad():
            for k in want:
                if k in state_dict:
                    n, s = k.rsplit(".", 1)
                    p = getattr(self, n.replace(".", "__") + "__" + s)
                    v = torch.as_tensor(state_dict[k])
                    if v.shape != p.shape:
                        raise RuntimeError(f"size mismatch for {k}: {tuple(v.shape)} vs {tuple(p.shape)}")
                    p.copy_(v)
        return self

    def tensors(self):
        sd = self.state_dict()
        return [sd[n + s].cpu().numpy() for n in self._names for s in (".weight", ".bias")]

    def forward(self, x):
        """``RRDBNet.forward`` (cnn_super_resolution.py:139-157): float NCHW in [0, 1] -> float NCHW at 4x, through the CUDA
        kernels of the ``RealESRGAN`` this module belongs to (``RealESRGAN.model``).  The kernels fuse the ``/ 255`` of
        ``enhance`` (:220) into the first conv and therefore take inputs ON THE UINT8 GRID (k / 255, what ``enhance`` and
        ``_tile_process`` feed the model); anything else raises ``ValueError``."""
        bound = getattr(self, "_wowsr", None)          # (libwowsr handle holding these weights, device): set by RealESRGAN
        if bound is None:
            raise RuntimeError("RRDBNet.forward runs through the kernels of a RealESRGAN: use RealESRGAN(...).model(x)")
        return _forward_float(bound[0], bound[1], x, None)


@torch.no_grad()
def _forward_float(h, device, x: torch.Tensor, tile_size=None) -> torch.Tensor:
    """Float NCHW on the uint8 grid -> float NCHW pre-quantisation output of the network loaded in handle `h`; untiled
    (``tile_size=None``) or with the reference's tile / tile_pad stitching."""
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError("expected a float tensor of shape (N, 3, H, W)")
    x = x.detach().to(device, torch.float32)
    q = torch.round(x * 255.0)
    if not bool(torch.all((q >= 0) & (q <= 255) & (torch.abs(x * 255.0 - q) < 1e-3))):
        raise ValueError("the CUDA path takes inputs on the uint8 grid (k / 255): the first conv reads uint8 windows")
    N, _, H, W = x.shape
    u8 = q.to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    out = torch.empty((N, 4 * H, 4 * W, 3), dtype=torch.float32, device=device)
    scratch = torch.empty((4 * H, 4 * W, 3), dtype=torch.uint8, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream
    # forward: the whole image is one window; _tile_process: ALWAYS the window table (the h*w > 4 T^2 switch belongs to enhance, :226)
    wins = [_lib.Window(0, 0, W, H, 0, 0, W, H)] if tile_size is None else _lib.plan_windows(H, W, tile_size, 10)
    with h.lock:
        for n in range(N):
            h.forward_windows(u8[n].data_ptr(), H, W, W * 3, wins, scratch.data_ptr(), 4 * W * 3, dst_f32_ptr=out[n].data_ptr(),
                              dst_f32_pitch=4 * W * 3 * 4, stream=stream)
    return out.permute(0, 3, 1, 2)


# Loaded-model residency.  The reference constructs and deletes a ``RealESRGAN`` per request
# (wow_sr.py:93,97; farm_sr.py:162): here that would re-read the .pth, rebuild 351 parameter tensors and
# re-pack / re-upload 67 MB of weights every time.  Loaded models stay resident, keyed by what determines the
# device image: (device, model, precision, weights identity).
_MODEL_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_MODEL_CACHE_MAX = 2  # each resident model owns a workspace sized by its largest batch (up to mem_budget_mb)
_cache_lock = __import__("threading").Lock()


def _weights_fingerprint(state_dict=None, path=None):
    if path is not None:
        st = Path(path).stat()
        return ("file", str(path), st.st_size, st.st_mtime_ns)
    # content fingerprint of an in-memory state dict: every byte of every tensor (67 MB for x4plus: tens of milliseconds) — a
    # sampled hash would let two checkpoints that differ in a few weights share one resident model
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    for k in sorted(state_dict):
        t = torch.as_tensor(state_dict[k]).detach()
        h.update(k.encode())
        h.update(str(tuple(t.shape)).encode())
        h.update(t.to(torch.float32).contiguous().cpu().numpy().tobytes())
    return ("dict", h.hexdigest())


def clear_model_cache():
    """Drops every resident model (frees their device memory once no ``RealESRGAN`` refers to them)."""
    with _cache_lock:
        _MODEL_CACHE.clear()


class RealESRGAN:
    """Real-ESRGAN inference wrapper with the reference's constructor and ``enhance`` (:161-234).

    Extra keyword-only arguments (all optional): ``state_dict`` to supply weights directly (the
    build box has no network for ``download_weights``), ``precision`` ("bf16" | "fp16") for the
    tensor-core operand type, ``handle`` to share a libwowsr handle.
    """

    def __init__(self, scale: int = 4, device: str = None, tile_size: int = 256, model_name: str = None, *,
                 state_dict=None, precision: str = "bf16", handle=None):
        self.tile_size = tile_size
        self.tile_pad = 10
        if device is None:
            self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        else:
            self.device = torch.device(device)
        if model_name is None:
            model_name = f"realesrgan_x{scale}"
        if model_name not in MODELS:
            raise ValueError(f"Unknown model: {model_name}. Available: {list(MODELS.keys())}")
        if self.device.type != "cuda":
            raise RuntimeError("this build runs on a B200 only (no CPU fallback); got device=%s" % self.device)
        config = MODELS[model_name]
        self.scale = config["scale"]
        self.model_name = model_name
        self.precision = precision
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        weights_path = None
        if state_dict is None:
            weights_path = download_weights(model_name)
        elif "params_ema" in state_dict:
            state_dict = state_dict["params_ema"]
        elif "params" in state_dict:
            state_dict = state_dict["params"]
        key = None
        if handle is None:
            key = (dev_index, model_name, precision, _weights_fingerprint(state_dict, weights_path))
            with _cache_lock:
                hit = _MODEL_CACHE.get(key)
                if hit is not None:
                    _MODEL_CACHE.move_to_end(key)
            if hit is not None:
                self.model, self._h = hit
                return
        if state_dict is None:
            state_dict = torch.load(weights_path, map_location="cpu")
            if "params_ema" in state_dict:
                state_dict = state_dict["params_ema"]
            elif "params" in state_dict:
                state_dict = state_dict["params"]
        self.model = RRDBNet(3, 3, config["channels"], config["blocks"], 32, self.scale)
        self.model.load_state_dict(state_dict, strict=True)
        self.model.eval()
        self._h = handle if handle is not None else _lib.Handle(dev_index)
        self._h.load_rrdbnet(self.model.tensors(), config["blocks"], config["channels"], 32, precision)
        self.model._wowsr = (self._h, self.device)
        if key is not None:
            with _cache_lock:
                _MODEL_CACHE[key] = (self.model, self._h)
                while len(_MODEL_CACHE) > _MODEL_CACHE_MAX:
                    _MODEL_CACHE.popitem(last=False)

    @torch.no_grad()
    def enhance(self, img: np.ndarray, outscale: int = 4) -> np.ndarray:
        """``img`` HxWx3 uint8 (channel order as given) -> 4Hx4Wx3 uint8, synchronous (:217-234)."""
        if outscale != self.scale:
            raise ValueError(f"outscale={outscale} is not supported by {self.model_name} (x{self.scale})")
        if img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("expected an HxWx3 image")
        with self._h.lock:
            return self._h.enhance_host(np.asarray(img).astype(np.uint8, copy=False), self.tile_size)

    @torch.no_grad()
    def enhance_float(self, img: np.ndarray):
        """(uint8 output, float32 pre-quantisation output) — used by the parity tests."""
        with self._h.lock:
            return self._h.enhance_host(np.asarray(img).astype(np.uint8, copy=False), self.tile_size, want_float=True)

    @torch.no_grad()
    def enhance_cuda(self, img: torch.Tensor) -> torch.Tensor:
        """Device-resident variant: uint8 HxWx3 CUDA tensor in, uint8 4Hx4Wx3 CUDA tensor out."""
        assert img.is_cuda and img.dtype == torch.uint8 and img.is_contiguous()
        H, W = img.shape[:2]
        out = torch.empty((4 * H, 4 * W, 3), dtype=torch.uint8, device=img.device)
        with self._h.lock:
            self._h.enhance_dev(img.data_ptr(), H, W, self.tile_size, out.data_ptr(),
                                stream=torch.cuda.current_stream(img.device).cuda_stream)
        return out

    def _tile_process(self, img: torch.Tensor) -> torch.Tensor:
        """``_tile_process`` (:236-280): float (N, 3, H, W) on the uint8 grid -> stitched float output.  The window planner and
        the last-writer-wins stitching run inside libwowsr (wowsr_plan_windows / wowsr_rrdbnet_forward_windows)."""
        return _forward_float(self._h, self.device, img, self.tile_size)


def apply_cnn_sr(input_path: Path, output_path: Path, scale: int = 4):
    """Same signature, outputs and metadata as the reference's file entry point (:283-375, called by sr_cli.py:115-124).
    GeoTIFF input: bands 1-3 (or one band replicated), non-uint8 rasters min-max stretched with the reference's ``+ 1e-6``
    (:308-311) on the device, RGB->BGR, ``enhance``, GeoTIFF out with the transform scaled.  Any other file: ``cv2.imread``
    (BGR) -> ``enhance`` -> PNG."""
    import cv2

    from . import wow_sr
    input_path = Path(input_path)
    is_tif = input_path.suffix.lower() in (".tif", ".tiff")
    transform = crs = None
    model = RealESRGAN(scale=scale, tile_size=256)
    if is_tif:
        img, transform, crs = wow_sr.read_image(input_path)                 # RGB, file dtype
        host = np.ascontiguousarray(img)
        if host.dtype == np.uint16:
            host = host.astype(np.int32)
        x = wow_sr.normalise_to_uint8_cuda(torch.from_numpy(host).to(model.device), eps=1e-6)
        in_shape = img.shape
        out_bgr = model.enhance_cuda(x.flip(2).contiguous())                # RGB -> BGR (:318-322)
        output_rgb = wow_sr._to_host(out_bgr.flip(2).contiguous())
        output_bgr = None
    else:
        img = cv2.imread(str(input_path))
        if img is None:
            raise FileNotFoundError(str(input_path))
        in_shape = img.shape
        output_bgr = model.enhance(img)
        output_rgb = output_bgr[:, :, ::-1]
    output_path = Path(output_path)
    output_path.parent.mkdir(parents=True, exist_ok=True)
    if transform is not None:
        import rasterio
        from rasterio.transform import Affine
        new_transform = Affine(transform.a / scale, transform.b, transform.c, transform.d, transform.e / scale, transform.f)
        final_path = output_path.with_suffix(".tif")
        with rasterio.open(final_path, "w", driver="GTiff", height=output_rgb.shape[0], width=output_rgb.shape[1], count=3,
                           dtype="uint8", crs=crs, transform=new_transform, compress="lzw") as dst:
            for i in range(3):
                dst.write(output_rgb[:, :, i], i + 1)
    else:
        final_path = output_path.with_suffix(".png")
        if output_bgr is None:  # a GeoTIFF without georeferencing: the reference writes its BGR result (:361-362)
            output_bgr = np.ascontiguousarray(output_rgb[:, :, ::-1])
        cv2.imwrite(str(final_path), output_bgr)
    metadata = {
        "model": f"RealESRGAN_x{scale}",
        "scale": scale,
        "input_size": [in_shape[1], in_shape[0]],
        "output_size": [output_rgb.shape[1], output_rgb.shape[0]],
        "device": str(model.device),
        "original_resolution_m": 10.0,
        "effective_resolution_m": 10.0 / scale,
    }
    return final_path, metadata
