"""GPU: BASELINE.json's full sizes (configs 2, 3, 5) through size-independent properties, plus one oracle window.

* tiled path: a window's owned output depends only on that window's LR pixels (cnn_super_resolution.py:256-257), so
  the full run must equal, bit for bit, the same window run on its own — checked for corner / edge / interior / shifted
  windows; the owned rectangles tile the output exactly once (planner);
* one 532x532 window of config 2 against the fp32 oracle (23 blocks, ~1 min of CPU) at the north_star tolerance;
* EDSR config 3 (1024x1024): an interior crop with its receptive field reproduces the full image's pixels."""
import os

import numpy as np
import pytest
import torch

from oracle import rrdbnet_ref as R

pytestmark = pytest.mark.gpu


def _lr_image(H, W, seed):
    import bench
    return bench.make_lr_image(H, W, seed=seed)


def _metrics(got_u8, ref_u8):
    d = np.abs(got_u8.astype(int) - ref_u8.astype(int))
    mse = float(((got_u8.astype(np.float64) - ref_u8) ** 2).mean())
    return float((d <= 1).mean()), (99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)), int(d.max())


def _window_alone(up, dimg, w):
    """The window as an untiled image of its own (tile_size large enough that enhance does not tile it)."""
    crop = dimg[w.y0:w.y1, w.x0:w.x1].contiguous()
    keep = up.tile_size
    up.tile_size = 4096
    try:
        out = up.enhance_cuda(crop)
    finally:
        up.tile_size = keep
    return out[4 * (w.oy0 - w.y0):4 * (w.oy1 - w.y0), 4 * (w.ox0 - w.x0):4 * (w.ox1 - w.x0)]


def test_cfg2_full_size_windows_and_oracle(ws):
    """Config 2: 4096x4096, tile_size=512 -> 64 windows of 532x532 (the last row/column shifted to start at 3564)."""
    blocks = 23
    sd = R.calibrate_conv_last(R.random_init_state_dict(0, blocks), blocks)
    up = ws.app.cnn_super_resolution.RealESRGAN(device="cuda", tile_size=512, state_dict=sd)
    img = _lr_image(4096, 4096, seed=1)
    dimg = torch.from_numpy(img).cuda()
    full = up.enhance_cuda(dimg)
    wins = ws._lib.plan_windows(4096, 4096, 512)
    assert len(wins) == 64 and all(w.x1 - w.x0 == 532 and w.y1 - w.y0 == 532 for w in wins)
    assert wins[63].x0 == 3564 and wins[63].y0 == 3564
    cover = torch.zeros((4096, 4096), dtype=torch.int32)
    for w in wins:
        cover[w.oy0:w.oy1, w.ox0:w.ox1] += 1
    assert bool((cover == 1).all())
    for i in (0, 3, 27, 56, 63):                      # corner, top edge, interior, bottom-left (shifted row), shifted corner
        w = wins[i]
        alone = _window_alone(up, dimg, w)
        assert torch.equal(full[4 * w.oy0:4 * w.oy1, 4 * w.ox0:4 * w.ox1], alone), i
    # corner, top edge, interior and the shifted corner window against the fp32 oracle (>= 4 windows)
    torch.set_num_threads(os.cpu_count())
    for i in (0, 3, 27, 63):
        w = wins[i]
        ref_f = R.enhance_float(sd, img[w.y0:w.y1, w.x0:w.x1], blocks, 4096)
        ref = R.quantise(ref_f)[4 * (w.oy0 - w.y0):4 * (w.oy1 - w.y0), 4 * (w.ox0 - w.x0):4 * (w.ox1 - w.x0)]
        got = full[4 * w.oy0:4 * w.oy1, 4 * w.ox0:4 * w.ox1].cpu().numpy()
        w1, psnr, mx = _metrics(got, ref)
        print(f"cfg2 window {i} vs fp32 oracle:", w1, psnr, mx)
        assert w1 >= 0.999 and psnr >= 50.0, (i, w1, psnr, mx)


def test_cfg5_full_scene_windows(ws):
    """Config 5: 10980x10980 scene, tile_size=256 -> 43 x 43 windows of 276x276, 43920x43920 output (5.8 GB)."""
    blocks = 23
    sd = R.calibrate_conv_last(R.random_init_state_dict(0, blocks), blocks)
    up = ws.app.cnn_super_resolution.RealESRGAN(device="cuda", tile_size=256, state_dict=sd)
    img = _lr_image(10980, 10980, seed=4)
    dimg = torch.from_numpy(img).cuda()
    full = up.enhance_cuda(dimg)
    assert tuple(full.shape) == (43920, 43920, 3)
    wins = ws._lib.plan_windows(10980, 10980, 256)
    assert len(wins) == 1849 and wins[-1].x0 == 10980 - 276
    for i in (0, 42, 43 * 20 + 21, 1848):
        w = wins[i]
        assert torch.equal(full[4 * w.oy0:4 * w.oy1, 4 * w.ox0:4 * w.ox1], _window_alone(up, dimg, w)), i
    # corner, interior and shifted-corner windows against the fp32 oracle (276x276: ~10 s of CPU each)
    torch.set_num_threads(os.cpu_count())
    for i in (0, 43 * 20 + 21, 1848):
        w = wins[i]
        ref_f = R.enhance_float(sd, img[w.y0:w.y1, w.x0:w.x1], blocks, 4096)
        ref = R.quantise(ref_f)[4 * (w.oy0 - w.y0):4 * (w.oy1 - w.y0), 4 * (w.ox0 - w.x0):4 * (w.ox1 - w.x0)]
        w1, psnr, mx = _metrics(full[4 * w.oy0:4 * w.oy1, 4 * w.ox0:4 * w.ox1].cpu().numpy(), ref)
        print(f"cfg5 window {i} vs fp32 oracle:", w1, psnr, mx)
        assert w1 >= 0.999 and psnr >= 50.0, (i, w1, psnr, mx)
    # determinism of the whole scene (checksum of checksums over row blocks)
    again = up.enhance_cuda(dimg)
    sums = [(int(full[y:y + 4392].to(torch.int64).sum()), int(again[y:y + 4392].to(torch.int64).sum())) for y in range(0, 43920, 4392)]
    assert all(a == b for a, b in sums)
    del full, again
    torch.cuda.empty_cache()


def test_cfg3_edsr_1024_interior_crop(ws):
    """Config 3: EDSR-baseline x4 on 1024x1024 (parity unpinned: the oracle is our restatement).  The network's receptive
    field is 2*16+3 convs at LR plus the tail: a crop with a 48-pixel margin reproduces the interior exactly."""
    import importlib
    from oracle import edsr_ref as E
    sr_mod = importlib.import_module("sentinel2-super-resolution-poc_b200.app.super_resolution")
    sd = E.random_init_state_dict(0, 16)
    sr, _ = sr_mod.create_sr_model(4, "edsr", state_dict=sd)
    img = _lr_image(1024, 1024, seed=2)
    out = sr.upsample(img)
    assert out.shape == (4096, 4096, 3)
    y0, x0, S, M = 400, 520, 96, 48
    torch.set_num_threads(os.cpu_count())
    ref = E.quantise(E.forward_float(sd, img[y0 - M:y0 + S + M, x0 - M:x0 + S + M], 16))[4 * M:4 * (M + S), 4 * M:4 * (M + S)]
    w1, psnr, mx = _metrics(out[4 * y0:4 * (y0 + S), 4 * x0:4 * (x0 + S)], ref)
    print("cfg3 edsr interior crop:", w1, psnr, mx)
    assert w1 >= 0.999 and psnr >= 50.0, (w1, psnr, mx)
