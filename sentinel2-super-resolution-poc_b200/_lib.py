"""ctypes binding of libwowsr.so (include/wowsr.h).  No CPU fallback: if the library or a CUDA
device is missing, calls raise."""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WOWSR_LIB", os.path.join(_HERE, "libwowsr.so"))  # override only for A/B kernel experiments

PREC = {"bf16_pure": 0, "fp16": 1, "bf16": 2, "mixed": 2}


class PostParams(C.Structure):
    _fields_ = [("clip_limit", C.c_double), ("sigma", C.c_double), ("alpha", C.c_float), ("beta", C.c_float),
                ("sat_boost", C.c_float), ("grid", C.c_int32), ("hue_lo", C.c_int32), ("hue_hi", C.c_int32),
                ("stages", C.c_int32), ("reserved", C.c_int32)]


class Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("pitch", C.c_int64), ("W", C.c_int32), ("H", C.c_int32),
                ("y0", C.c_int32), ("rows", C.c_int32)]


class Window(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
                ("ox0", C.c_int32), ("oy0", C.c_int32), ("ox1", C.c_int32), ("oy1", C.c_int32)]


class HsvRange(C.Structure):
    _fields_ = [("lo", C.c_uint8 * 3), ("hi", C.c_uint8 * 3)]


STAGE_CLAHE, STAGE_UNSHARP, STAGE_VEG, STAGE_ALL = 1, 2, 4, 7

_lib = None
_lock = threading.Lock()

_SIGS = {
    "wowsr_abi_version": (C.c_int, []),
    "wowsr_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "wowsr_destroy": (None, [C.c_void_p]),
    "wowsr_last_error": (C.c_char_p, [C.c_void_p]),
    "wowsr_launch_count": (C.c_uint64, [C.c_void_p]),
    "wowsr_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "wowsr_get_option": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "wowsr_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "wowsr_tiles_resample": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int64,
                                       C.c_int32, C.c_int32, C.c_void_p]),
    "wowsr_post_params_wow": (None, [C.POINTER(PostParams)]),
    "wowsr_post_params_farm": (None, [C.POINTER(PostParams)]),
    "wowsr_clahe_hist": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "wowsr_clahe_luts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]),
    "wowsr_post_apply": (C.c_int, [C.c_void_p, C.POINTER(Image), C.c_void_p, C.POINTER(PostParams), C.c_int32, C.c_int32,
                                   C.POINTER(Image), C.c_void_p]),
    "wowsr_post_process_dev": (C.c_int, [C.c_void_p, C.POINTER(Image), C.POINTER(PostParams), C.POINTER(Image), C.c_void_p]),
    "wowsr_post_process_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(PostParams), C.c_void_p]),
    "wowsr_green_mask": (C.c_int, [C.c_void_p, C.POINTER(Image), C.POINTER(HsvRange), C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "wowsr_green_mask_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(HsvRange), C.c_int32, C.c_void_p]),
    "wowsr_clahe_geometry": (None, [C.c_int32, C.c_int32, C.c_int32] + [C.POINTER(C.c_int32)] * 4),
    "wowsr_get_table": (C.c_int64, [C.c_int32, C.c_void_p, C.c_int64]),
    "wowsr_gaussian_taps": (C.c_int32, [C.c_double, C.POINTER(C.c_int32), C.c_int32]),
    "wowsr_plan_windows": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Window), C.c_int32]),
    "wowsr_load_rrdbnet": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]),
    "wowsr_rrdbnet_forward_windows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.POINTER(Window),
                                                C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "wowsr_enhance_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "wowsr_enhance_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wowsr_conv3x3_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "wowsr_get_timing": (C.c_int32, [C.c_void_p, C.POINTER(C.c_float), C.c_int32]),
    "wowsr_debug_trace": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]),
    "wowsr_debug_roll_plan": (C.c_int32, [C.c_int32] * 6 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "wowsr_load_edsr": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]),
    "wowsr_edsr_upsample_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "wowsr_edsr_upsample_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def lib():
    """Loads libwowsr.so (built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(there is no CPU fallback)")
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGS.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGS)


class WowsrError(RuntimeError):
    pass


def _locked(fn):
    """Serialises a Handle method on the handle's lock: a wowsr_ctx keeps its scratch (histograms, LUTs, staging buffers,
    workspace) inside the context, and the server runs /api/wow, /api/sr and /api/pipeline jobs on concurrent worker threads
    (main.py:519,607,670-675) that share the per-device default handle."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        with self.lock:
            return fn(self, *a, **kw)
    return wrapper


class Handle:
    """One wowsr_ctx bound to a CUDA device."""

    def __init__(self, device: int = 0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.wowsr_create(int(device), C.byref(h))
        if rc != 0:
            raise WowsrError(f"wowsr_create({device}) failed ({rc}): {self._L.wowsr_last_error(None).decode()}")
        self._h = h
        self.device = int(device)
        self._keep = []
        # a handle owns ONE workspace and stream set (include/wowsr.h): callers that share a handle between
        # threads (the server caches loaded models across requests) serialise network calls on this lock
        self.lock = threading.RLock()

    def close(self):
        if getattr(self, "_h", None):
            self._L.wowsr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise WowsrError(f"{what} failed ({rc}): {self._L.wowsr_last_error(self._h).decode()}")

    # -- misc ---------------------------------------------------------------------------------
    @_locked
    def set_option(self, key: str, value: int):
        self._check(self._L.wowsr_set_option(self._h, key.encode(), int(value)), "set_option")

    def launch_count(self) -> int:
        return int(self._L.wowsr_launch_count(self._h))

    def timing(self):
        buf = (C.c_float * 5)()
        n = self._L.wowsr_get_timing(self._h, buf, 5)
        return dict(zip(("total", "head", "trunk", "tail", "enqueue_host"), list(buf)[:n]))

    def debug_trace(self):
        buf = (C.c_int64 * 512)()  # rows 0..63: tile timeline; rows 64..127: epilogue breakdown (instrumented builds)
        n = self._L.wowsr_debug_trace(self._h, buf, 512)
        return np.array(list(buf)[:max(n, 0)], dtype=np.int64).reshape(-1, 4)

    @_locked
    def download(self, tensor) -> np.ndarray:
        """A contiguous CUDA tensor -> a new numpy array, through the pinned staging ring (no page-locking of the destination)."""
        import torch
        assert tensor.is_cuda and tensor.is_contiguous()
        out = np.empty(tuple(tensor.shape), dtype=np.dtype(str(tensor.dtype).replace("torch.", "")))
        rows = int(tensor.shape[0]) if tensor.dim() > 1 else 1
        row_bytes = tensor.numel() * tensor.element_size() // max(rows, 1)
        if tensor.numel() == 0:
            return out
        self._check(self._L.wowsr_download(self._h, C.c_void_p(tensor.data_ptr()), row_bytes, row_bytes, rows, out.ctypes.data, row_bytes,
                                           C.c_void_p(torch.cuda.current_stream(tensor.device).cuda_stream)), "download")
        return out

    @_locked
    def tiles_resample(self, src: Image, sx0, sy0, sxp, syp, out_ptr, out_pitch, OW, OH, stream=0):
        self._check(self._L.wowsr_tiles_resample(self._h, C.byref(src), float(sx0), float(sy0), float(sxp), float(syp), C.c_void_p(out_ptr),
                                                 out_pitch, OW, OH, C.c_void_p(stream)), "tiles_resample")

    # -- post-process -------------------------------------------------------------------------
    @_locked
    def post_process_host(self, img: np.ndarray, params: PostParams) -> np.ndarray:
        img = np.ascontiguousarray(img, dtype=np.uint8)
        if img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("expected HxWx3 uint8")
        out = np.empty_like(img)
        self._check(self._L.wowsr_post_process_host(self._h, img.ctypes.data, img.shape[0], img.shape[1], C.byref(params),
                                                    out.ctypes.data), "post_process_host")
        return out

    @_locked
    def post_process_dev(self, src_ptr, dst_ptr, H, W, params, stream=0, pitch=None):
        pitch = pitch or W * 3
        a = Image(src_ptr, pitch, W, H, 0, H)
        b = Image(dst_ptr, pitch, W, H, 0, H)
        self._check(self._L.wowsr_post_process_dev(self._h, C.byref(a), C.byref(params), C.byref(b), C.c_void_p(stream)),
                    "post_process_dev")

    @_locked
    def clahe_hist(self, image: Image, grid, prow0, prow1, hist_ptr, stream=0):
        self._check(self._L.wowsr_clahe_hist(self._h, C.byref(image), grid, prow0, prow1, C.c_void_p(hist_ptr), C.c_void_p(stream)),
                    "clahe_hist")

    @_locked
    def clahe_luts(self, hist_ptr, grid, tw, th, clip, luts_ptr, stream=0):
        self._check(self._L.wowsr_clahe_luts(self._h, C.c_void_p(hist_ptr), grid, tw, th, float(clip), C.c_void_p(luts_ptr),
                                             C.c_void_p(stream)), "clahe_luts")

    @_locked
    def post_apply(self, src: Image, luts_ptr, params, row0, row1, dst: Image, stream=0):
        self._check(self._L.wowsr_post_apply(self._h, C.byref(src), C.c_void_p(luts_ptr), C.byref(params), row0, row1, C.byref(dst),
                                             C.c_void_p(stream)), "post_apply")

    @staticmethod
    def _hsv_ranges(ranges):
        """[((h0, s0, v0), (h1, s1, v1)), ...] -> ctypes array; bounds are inclusive like cv2.inRange."""
        arr = (HsvRange * len(ranges))()
        for a, (lo, hi) in zip(arr, ranges):
            for c in range(3):
                if not (0 <= int(lo[c]) <= 255 and 0 <= int(hi[c]) <= 255):
                    raise ValueError("HSV bounds must fit uint8")
                a.lo[c], a.hi[c] = int(lo[c]), int(hi[c])
        return arr

    @_locked
    def green_mask_host(self, img: np.ndarray, ranges) -> np.ndarray:
        img = np.ascontiguousarray(img, dtype=np.uint8)
        if img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("expected HxWx3 uint8")
        out = np.empty(img.shape[:2], dtype=np.float32)
        arr = self._hsv_ranges(ranges)
        self._check(self._L.wowsr_green_mask_host(self._h, img.ctypes.data, img.shape[0], img.shape[1], arr, len(arr),
                                                  out.ctypes.data), "green_mask_host")
        return out

    @_locked
    def green_mask_dev(self, src_ptr, H, W, ranges, mask_ptr, stream=0, pitch=None, mask_pitch=None):
        a = Image(src_ptr, pitch or W * 3, W, H, 0, H)
        arr = self._hsv_ranges(ranges)
        self._check(self._L.wowsr_green_mask(self._h, C.byref(a), arr, len(arr), C.c_void_p(mask_ptr), mask_pitch or W * 4,
                                             C.c_void_p(stream)), "green_mask")

    # -- network ------------------------------------------------------------------------------
    @_locked
    def load_rrdbnet(self, tensors, num_block, num_feat=64, num_grow=32, precision="bf16"):
        arrs = [np.ascontiguousarray(t, dtype=np.float32) for t in tensors]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        self._check(self._L.wowsr_load_rrdbnet(self._h, num_block, num_feat, num_grow, ptrs, len(arrs), PREC[precision]),
                    "load_rrdbnet")

    @_locked
    def enhance_host(self, img: np.ndarray, tile_size: int, want_float=False):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        H, W = img.shape[:2]
        out = np.empty((H * 4, W * 4, 3), dtype=np.uint8)
        outf = np.empty((H * 4, W * 4, 3), dtype=np.float32) if want_float else None
        self._check(self._L.wowsr_enhance_host(self._h, img.ctypes.data, H, W, tile_size, out.ctypes.data,
                                               outf.ctypes.data if want_float else None), "enhance_host")
        return (out, outf) if want_float else out

    @_locked
    def enhance_dev(self, src_ptr, H, W, tile_size, dst_ptr, dst_f32_ptr=None, stream=0):
        self._check(self._L.wowsr_enhance_dev(self._h, C.c_void_p(src_ptr), H, W, tile_size, C.c_void_p(dst_ptr),
                                              C.c_void_p(dst_f32_ptr) if dst_f32_ptr else None, C.c_void_p(stream)), "enhance_dev")

    @_locked
    def forward_windows(self, src_ptr, H, W, pitch, windows, dst_ptr, dst_pitch, dst_f32_ptr=None, dst_f32_pitch=0, stream=0):
        arr = (Window * len(windows))(*windows)
        self._check(self._L.wowsr_rrdbnet_forward_windows(self._h, C.c_void_p(src_ptr), H, W, pitch, arr, len(windows),
                                                          C.c_void_p(dst_ptr), dst_pitch,
                                                          C.c_void_p(dst_f32_ptr) if dst_f32_ptr else None, dst_f32_pitch,
                                                          C.c_void_p(stream)), "rrdbnet_forward_windows")

    @_locked
    def conv3x3_host(self, x, weight, bias, act=0, precision="bf16"):
        x = np.ascontiguousarray(x, dtype=np.float32)
        weight = np.ascontiguousarray(weight, dtype=np.float32)
        bias = np.ascontiguousarray(bias, dtype=np.float32)
        n, h, w, cin = x.shape
        cout = weight.shape[0]
        out = np.empty((n, h, w, cout), dtype=np.float32)
        self._check(self._L.wowsr_conv3x3_host(self._h, x.ctypes.data, n, h, w, cin, weight.ctypes.data, bias.ctypes.data, cout,
                                               int(act), PREC[precision], out.ctypes.data), "conv3x3_host")
        return out

    @_locked
    def load_edsr(self, tensors, num_block=16, num_feat=64, res_scale=1.0, precision="bf16"):
        arrs = [np.ascontiguousarray(t, dtype=np.float32) for t in tensors]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        self._check(self._L.wowsr_load_edsr(self._h, num_block, num_feat, float(res_scale), ptrs, len(arrs), PREC[precision]),
                    "load_edsr")

    @_locked
    def edsr_upsample_host(self, img: np.ndarray, want_float=False):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        H, W = img.shape[:2]
        out = np.empty((H * 4, W * 4, 3), dtype=np.uint8)
        outf = np.empty((H * 4, W * 4, 3), dtype=np.float32) if want_float else None
        self._check(self._L.wowsr_edsr_upsample_host(self._h, img.ctypes.data, H, W, out.ctypes.data,
                                                     outf.ctypes.data if want_float else None), "edsr_upsample_host")
        return (out, outf) if want_float else out


    @_locked
    def edsr_upsample_dev(self, src_ptr, H, W, dst_ptr, dst_f32_ptr=None, stream=0):
        self._check(self._L.wowsr_edsr_upsample_dev(self._h, C.c_void_p(src_ptr), H, W, C.c_void_p(dst_ptr),
                                                    C.c_void_p(dst_f32_ptr) if dst_f32_ptr else None, C.c_void_p(stream)), "edsr_upsample_dev")


# -- pure host helpers (no GPU needed) ----------------------------------------------------------

def plan_windows(H, W, tile, pad=10):
    L = lib()
    n = L.wowsr_plan_windows(H, W, tile, pad, None, 0)
    if n < 0:
        raise ValueError("bad planner arguments")
    arr = (Window * n)()
    L.wowsr_plan_windows(H, W, tile, pad, arr, n)
    return list(arr)


def clahe_geometry(H, W, grid=8):
    v = [C.c_int32() for _ in range(4)]
    lib().wowsr_clahe_geometry(H, W, grid, *[C.byref(x) for x in v])
    return tuple(x.value for x in v)  # tile_w, tile_h, padded_w, padded_h


def gaussian_taps(sigma):
    buf = (C.c_int32 * 32)()
    n = lib().wowsr_gaussian_taps(float(sigma), buf, 32)
    if n < 0:
        raise ValueError("sigma too large")
    return list(buf)[:n]


def roll_plan(n_win, h, w, strip_x0=None, pair=True, max_units=74):
    """Work list of the rolling conv kernel (csrc/roll_kernel.cuh) as ((n, 8) int32 tasks, offsets, info dict)."""
    L = lib()
    strip_x0 = w if strip_x0 is None else strip_x0
    info = (C.c_int32 * 4)()
    n = L.wowsr_debug_roll_plan(n_win, h, w, strip_x0, int(pair), max_units, None, 0, None, 0, info)
    if n < 0:
        raise ValueError("bad rolling-plan arguments")
    tasks = np.empty((n, 8), dtype=np.int32)
    off = np.empty(info[2], dtype=np.int32)
    L.wowsr_debug_roll_plan(n_win, h, w, strip_x0, int(pair), max_units, tasks.ctypes.data, n, off.ctypes.data, len(off), info)
    return tasks, off, {"units": info[0], "units_h": info[1]}


_TABLES = {0: ("gam", np.uint16, 256), 1: ("cbrt", np.uint16, 3072), 2: ("lab_y", np.uint16, 256),
           3: ("lab_ify", np.uint16, 256), 4: ("invgam", np.uint8, 4096), 5: ("sdiv", np.uint32, 256),
           6: ("hdiv", np.uint32, 256)}


def get_tables():
    out = {}
    for tid, (name, dt, n) in _TABLES.items():
        a = np.empty(n, dtype=dt)
        got = lib().wowsr_get_table(tid, a.ctypes.data, a.nbytes)
        if got != a.nbytes:
            raise RuntimeError(f"table {name}: {got}")
        out[name] = a
    return out


def post_params(kind="wow", **over) -> PostParams:
    p = PostParams()
    (lib().wowsr_post_params_wow if kind == "wow" else lib().wowsr_post_params_farm)(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


_handles = {}


_handles_lock = threading.Lock()


def default_handle(device: int = 0) -> Handle:
    """The process-wide handle of `device` (created once, inside the critical section)."""
    with _handles_lock:
        h = _handles.get(device)
        if h is None:
            h = _handles[device] = Handle(device)
    return h
