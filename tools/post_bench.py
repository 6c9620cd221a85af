"""Times the post-process kernels alone (device-resident 4096x4096 and 16384x16384 RGB) with CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wowsr_b200 as ws
from tests.conftest import image_like
h = ws.Handle(0)
h.set_option('hist_match', int(sys.argv[1]) if len(sys.argv) > 1 else 0)   # 0 = production path (plain per-lane shared atomics)
for size in (4096, 16384):
    base = image_like(1024, 1024, seed=3)
    img = torch.from_numpy(base).cuda().repeat(size // 1024, size // 1024, 1).contiguous()
    out = torch.empty_like(img)
    hist = torch.zeros(64 * 256, dtype=torch.int32, device="cuda")
    luts = torch.zeros(64 * 256, dtype=torch.uint8, device="cuda")
    for kind in ("wow", "farm"):
        p = ws._lib.post_params(kind)
        tw, th, pw, ph = ws._lib.clahe_geometry(size, size, 8)
        im = ws._lib.Image(img.data_ptr(), size * 3, size, size, 0, size)
        om = ws._lib.Image(out.data_ptr(), size * 3, size, size, 0, size)
        def run_hist():
            hist.zero_(); h.clahe_hist(im, 8, 0, ph, hist.data_ptr())
        def run_apply():
            h.post_apply(im, luts.data_ptr(), p, 0, size, om)
        run_hist(); h.clahe_luts(hist.data_ptr(), 8, tw, th, p.clip_limit, luts.data_ptr()); run_apply()
        torch.cuda.synchronize()
        res = {}
        for name, fn in (("hist", run_hist), ("apply", run_apply)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3): fn()
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 10
        px = size * size
        tot = res["hist"] + res["apply"]
        print(f"{kind} {size}^2: hist {res['hist']*1e3:.0f} us ({px*3/res['hist']/1e6:.0f} GB/s)  apply {res['apply']*1e3:.0f} us ({px*6/res['apply']/1e6:.0f} GB/s)  "
              f"total {tot*1e3:.0f} us -> {px*9/tot/1e6:.0f} GB/s algorithmic = {px*9/tot/1e6/6544.3*100:.1f}% of HBM peak, {px/tot/1e3:.0f} Mpix/s", flush=True)
