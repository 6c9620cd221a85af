"""Quick GPU check of the post-process kernels against cv2 (run on the GPU box)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import numpy as np  # noqa: E402

import wowsr_b200 as ws  # noqa: E402
from oracle import wow_cv2  # noqa: E402


def mk(h, w, seed=3):
    r = np.random.default_rng(seed)
    a = r.integers(0, 256, (h, w, 3)).astype(np.float32)
    a = cv2.GaussianBlur(a, (0, 0), 3)
    a[..., 1] += 40
    a = (a - a.min()) / (a.max() - a.min()) * 255
    return a.astype(np.uint8)


h = ws.Handle(0)
for (H, W) in [(64, 64), (512, 512), (517, 1003), (300, 200), (1104, 1104), (2048, 2048)]:
    img = mk(H, W)
    for kind, fn in (("wow", wow_cv2.enhance_for_crops), ("farm", wow_cv2.farm_post)):
        ref = fn(img)
        t0 = time.time()
        got = h.post_process_host(img, ws._lib.post_params(kind))
        dt = time.time() - t0
        d = (ref != got)
        print(f"{kind} {H}x{W}: mismatched px={int(d.any(-1).sum())} maxdiff={int(np.abs(ref.astype(int)-got.astype(int)).max())} ({dt*1e3:.1f} ms)", flush=True)
