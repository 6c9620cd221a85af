"""CPU: the C-ABI library loads and exports every symbol include/wowsr.h declares; host-side logic
(window planner, geometry, tables, Gaussian taps) agrees with the oracle.  No compute calls."""
import os
import re

import numpy as np
import pytest

from oracle import postproc_np as P
from oracle import rrdbnet_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header(ws):
    hdr = open(os.path.join(ROOT, "include", "wowsr.h")).read()
    declared = set(re.findall(r"\b(wowsr_[a-z0-9_]+)\s*\(", hdr))
    lib = ws._lib.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in wowsr.h but not exported"
    assert declared == set(ws._lib.exported_symbols())
    assert lib.wowsr_abi_version() == 1


def test_no_cpu_fallback(ws):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ws.WowsrError, match="no CPU fallback"):
        ws.Handle(0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ws.app.cnn_super_resolution.RealESRGAN(device="cpu", state_dict={})


def test_tables_match_oracle(ws):
    t = ws._lib.get_tables()
    gam, cbrt = P.lab_tables()
    y, ify, _, ig = P.lab2rgb_tables()
    sd, hd = P.hsv_tables()
    for name, ref in (("gam", gam), ("cbrt", cbrt), ("lab_y", y), ("lab_ify", ify), ("invgam", ig), ("sdiv", sd), ("hdiv", hd)):
        assert np.array_equal(t[name].astype(np.int64), np.asarray(ref).astype(np.int64)), name


def test_gaussian_taps_and_geometry(ws):
    for s in (1.0, 1.2, 1.5, 2.0, 0.8):
        assert ws._lib.gaussian_taps(s) == list(P.gaussian_kernel_u8(s))
    for (h, w) in [(512, 512), (517, 1003), (43920, 43920), (1104, 1104), (9, 9)]:
        tw, th, pw, ph = ws._lib.clahe_geometry(h, w, 8)
        assert (tw, th, pw, ph) == P.clahe_geometry(h, w, 8)


def _simulate_last_writer(H, W, T):
    """Owner id per LR pixel by replaying the reference loop order (cnn_super_resolution.py:247-278)."""
    own = -np.ones((H, W), dtype=np.int64)
    i = 0
    for (y1, y2, ylo, yhi) in R.plan_axis(H, T):
        for (x1, x2, xlo, xhi) in R.plan_axis(W, T):
            own[ylo:yhi, xlo:xhi] = i
            i += 1
    return own


@pytest.mark.parametrize("H,W,T", [(513, 512, 256), (600, 700, 256), (1100, 1100, 256), (300, 2000, 256), (2048, 2048, 512),
                                   (257, 1030, 256), (70, 50, 16), (10980, 10980, 256), (4096, 4096, 512), (528, 276, 128)])
def test_planner_matches_reference_loop(ws, H, W, T):
    wins = ws._lib.plan_windows(H, W, T)
    if H * W <= 4 * T * T:
        assert len(wins) == 1 and (wins[0].x0, wins[0].y0, wins[0].x1, wins[0].y1) == (0, 0, W, H)
        return
    ys, xs = R.plan_axis(H, T), R.plan_axis(W, T)
    assert len(wins) == len(ys) * len(xs)
    big = H * W > 4_000_000
    own = None if big else _simulate_last_writer(H, W, T)
    got = None if big else -np.ones((H, W), dtype=np.int64)
    sizes = set()
    i = 0
    for (y1, y2, _, _) in ys:
        for (x1, x2, _, _) in xs:
            w = wins[i]
            assert (w.x0, w.y0, w.x1, w.y1) == (x1, y1, x2, y2)
            sizes.add((w.x1 - w.x0, w.y1 - w.y0))
            if not big and w.ox1 > w.ox0 and w.oy1 > w.oy0:
                assert (got[w.oy0:w.oy1, w.ox0:w.ox1] == -1).all(), "owned rectangles overlap"
                got[w.oy0:w.oy1, w.ox0:w.ox1] = i
            i += 1
    assert len(sizes) == 1                      # all windows share one size -> one batched launch
    if not big:
        assert np.array_equal(got, own)         # ownership == last-writer-wins of the reference loop
    else:
        area = sum(max(0, w.ox1 - w.ox0) * max(0, w.oy1 - w.oy0) for w in wins)
        assert area == H * W


def test_rrdbnet_container_state_dict_roundtrip(ws):
    import torch
    cnn = ws.app.cnn_super_resolution
    torch.manual_seed(0)
    m = cnn.RRDBNet(num_block=2)
    sd = R.random_init_state_dict(0, 2)
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert all(torch.equal(sd[k], v) for k, v in m.state_dict().items())   # same RNG stream as the reference order
    m2 = cnn.RRDBNet(num_block=2)
    m2.load_state_dict(sd, strict=True)
    assert len(m2.tensors()) == 72
    with pytest.raises(RuntimeError):
        m2.load_state_dict({"conv_first.weight": sd["conv_first.weight"]}, strict=True)
    with pytest.raises(ValueError):
        cnn.download_weights("nope")
    assert set(cnn.MODELS) == {"realesrgan_x4", "realesrgan_anime"} and cnn.MODELS["realesrgan_x4"]["blocks"] == 23


def test_option_key_list_matches_the_keys_the_library_reads():
    """wowsr_set_option rejects unknown keys (include/wowsr.h); its list must hold exactly the keys some wowsr_opt() reads."""
    csrc = os.path.join(ROOT, "sentinel2-super-resolution-poc_b200", "csrc")
    read = set()
    for fn in os.listdir(csrc):
        if fn.endswith((".cu", ".cuh", ".h")):
            read |= set(re.findall(r'wowsr_opt\(\s*ctx,\s*"([a-z0-9_]+)"', open(os.path.join(csrc, fn)).read()))
    src = open(os.path.join(csrc, "ctx.cu")).read()
    block = src[src.index("kOptionKeys[] = {"):]
    listed = set(re.findall(r'"([a-z0-9_]+)"', block[:block.index("};")]))
    assert listed == read
    for path in ("tests/test_gpu_rrdbnet.py", "tools/gpu_probe.py"):
        text = open(os.path.join(ROOT, path)).read()
        used = set(re.findall(r"\b((?:tc|trunk|conv|hist|mem)_[a-z0-9_]+)\s*=\s*\d", text)) | set(re.findall(r'"((?:tc|trunk|conv|hist|mem)_[a-z0-9_]+)"\s*:', text))
        assert used <= listed, (path, used - listed)
