"""CPU (gloo, world_size 2 and 3): the multi-GPU scene decomposition of scene.py — tile-row bands, CLAHE
histogram all-reduce, seam halo exchange, gather — reproduces the single-process oracle bit for bit.
The compute backend here is the CPU oracle (this is a test of the sharding logic, not of the kernels)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleBackend:
    """CPU stand-in for scene.GpuBackend built on oracle/ (numpy) with a cheap fake x4 'network'."""

    def __init__(self, params, lib):
        self.params = params
        self.lib = lib

    def blur_radius(self):
        taps = self.lib.gaussian_taps(self.params.sigma)
        half = len(taps) // 2
        r = half
        while r > 0 and taps[half - r] == 0:
            r -= 1
        return r

    @staticmethod
    def fake_sr(win):
        """Deterministic x4 'model' whose output depends on position inside the window (so wrong window
        geometry or ownership shows up), applied per window like the reference applies its model."""
        up = np.repeat(np.repeat(win, 4, 0), 4, 1).astype(np.int32)
        yy, xx = np.mgrid[0:up.shape[0], 0:up.shape[1]]
        return ((up + (yy * 3 + xx * 5)[..., None]) % 256).astype(np.uint8)

    def sr_band(self, img, plan, band, row_off):
        a = img.numpy()
        for w in plan.windows:
            out = self.fake_sr(a[w.y0:w.y1, w.x0:w.x1])
            ys, xs = slice(4 * (w.oy0 - w.y0), 4 * (w.oy1 - w.y0)), slice(4 * (w.ox0 - w.x0), 4 * (w.ox1 - w.x0))
            band[4 * w.oy0 - row_off:4 * w.oy1 - row_off, 4 * w.ox0:4 * w.ox1] = torch.from_numpy(out[ys, xs].copy())

    def hist(self, band, plan, row_off, prow0, prow1, hist):
        from oracle import postproc_np as P
        tw, th, pw, ph = P.clahe_geometry(plan.OH, plan.OW, self.params.grid)
        L = P.rgb2l_u8(band.numpy())
        g = self.params.grid
        h = hist.view(g, g, 256)
        for pr in range(prow0, prow1):
            sy = pr if pr < plan.OH else 2 * plan.OH - 2 - pr
            row = L[sy - row_off]
            xs = np.arange(pw)
            xs = np.where(xs >= plan.OW, 2 * plan.OW - 2 - xs, xs)
            rowp = row[xs]
            for tx in range(g):
                h[pr // th, tx] += torch.from_numpy(np.bincount(rowp[tx * tw:(tx + 1) * tw], minlength=256).astype(np.int32))

    def luts(self, hist, plan):
        from oracle import postproc_np as P
        tw, th, _, _ = P.clahe_geometry(plan.OH, plan.OW, self.params.grid)
        g = self.params.grid
        return torch.from_numpy(P.clahe_luts(hist.view(g, g, 256).numpy().astype(np.uint32), tw * th, self.params.clip_limit))

    def apply(self, band, plan, row_off, luts, out):
        from oracle import postproc_np as P
        p = self.params
        r = self.blur_radius()
        # evaluate the per-pixel part on the stored rows, then blur with reflect-101 only at true image borders
        tw, th, _, _ = P.clahe_geometry(plan.OH, plan.OW, p.grid)
        a = band.numpy()
        lab = P.rgb2lab_u8(a)
        full_L = np.zeros((plan.OH, plan.OW), np.uint8)
        full_L[row_off:row_off + a.shape[0]] = lab[..., 0]
        lab[..., 0] = P.clahe_interp(full_L, luts.numpy(), tw, th)[row_off:row_off + a.shape[0]]
        enh = P.lab2rgb_u8(lab)
        top_pad = r if row_off == 0 else 0
        bot_pad = r if row_off + a.shape[0] == plan.OH else 0
        ext = np.pad(enh, ((top_pad, bot_pad), (0, 0), (0, 0)), mode="reflect")
        blur = P.gaussian_blur_u8(ext, p.sigma)
        y0 = plan.Y0 - row_off + top_pad
        sharp = P.add_weighted_u8(ext[y0:y0 + plan.Y1 - plan.Y0], p.alpha, blur[y0:y0 + plan.Y1 - plan.Y0], p.beta)
        out[:plan.Y1 - plan.Y0] = torch.from_numpy(P.enhance_vegetation(sharp, p.sat_boost, p.hue_lo, p.hue_hi))

    def new_band(self, rows, width):
        return torch.zeros((rows, width, 3), dtype=torch.uint8)

    def new_hist(self):
        return torch.zeros(self.params.grid ** 2 * 256, dtype=torch.int32)


def _expected(img, tile, params):
    """Single-process truth: reference tiling loop (oracle) with the same fake model, then the oracle post-process."""
    from oracle import postproc_np as P
    from oracle import rrdbnet_ref as R
    H, W = img.shape[:2]
    out = np.zeros((4 * H, 4 * W, 3), np.uint8)
    if H * W > 4 * tile * tile:
        for (y1, y2, ylo, yhi) in R.plan_axis(H, tile):
            for (x1, x2, xlo, xhi) in R.plan_axis(W, tile):
                t = OracleBackend.fake_sr(img[y1:y2, x1:x2])
                out[4 * ylo:4 * yhi, 4 * xlo:4 * xhi] = t[4 * (ylo - y1):4 * (yhi - y1), 4 * (xlo - x1):4 * (xhi - x1)]
    else:
        out = OracleBackend.fake_sr(img)
    pp = dict(clip=params.clip_limit, grid=params.grid, sigma=params.sigma, alpha=params.alpha, beta=params.beta,
              hue_lo=params.hue_lo, hue_hi=params.hue_hi, sat=params.sat_boost)
    return P.post_process(out, pp)


def _worker(rank, world, port, H, W, tile, kind, q, balance="windows"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    import wowsr_b200 as ws
    scene = importlib.import_module("sentinel2-super-resolution-poc_b200.scene")
    params = ws._lib.post_params(kind)
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    if balance == "shared_host":   # every rank writes its band into one shared host image (no gather on rank 0)
        shared = scene.SharedHostImage(4 * H, 4 * W)
        plan, up_bytes = scene.run_scene_to_host(OracleBackend(params, ws._lib), torch.from_numpy(img), tile, shared)
        assert 0 <= up_bytes <= img.nbytes
        full = shared.array.clone()
        shared.close()
    else:
        plan, band, full = scene.run_scene(OracleBackend(params, ws._lib), torch.from_numpy(img), tile, post=True, gather=True,
                                           balance=balance)
    if rank == 0:
        want = _expected(img, tile, params)
        q.put((bool(np.array_equal(full.numpy(), want)), [tuple(b) for b in plan.bands]))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,H,W,tile,kind,balance", [(2, 70, 50, 16, "wow", "windows"), (3, 70, 50, 16, "farm", "windows"),
                                                         (2, 37, 45, 64, "wow", "windows"), (3, 70, 50, 16, "wow", "rows"),
                                                         (3, 40, 90, 16, "wow", "windows"), (2, 70, 50, 16, "wow", "shared_host")])
def test_sharded_scene_equals_single_process(world, H, W, tile, kind, balance):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, tile, kind, q, balance)) for r in range(world)]
    for p in procs:
        p.start()
    ok, bands = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, bands
    assert bands[0][0] == 0 and bands[-1][1] == 4 * H
    assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))


def test_split_rows_and_plan():
    import importlib
    scene = importlib.import_module("sentinel2-super-resolution-poc_b200.scene")
    assert scene.split_rows(43, 8) == [(0, 6), (6, 12), (12, 18), (18, 23), (23, 28), (28, 33), (33, 38), (38, 43)]
    assert scene.split_rows(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for balance, spread in (("rows", 43), ("windows", 1)):
        total, counts = 0, []
        for r in range(8):
            p = scene.ScenePlan(10980, 10980, 256, 8, r, balance=balance)
            total += len(p.windows)
            counts.append(len(p.windows))
            assert p.Y1 > p.Y0
        assert total == 1849 and max(counts) - min(counts) <= spread, (balance, counts)
    p = scene.ScenePlan(10980, 10980, 256, 8, 3)
    # every cut tile row is shipped exactly once, to the rank that owns the row
    assert len(p.pieces) == 7 and all(src != dst for (src, dst, *_rest) in p.pieces)
    assert p.bands[0][0] == 0 and p.bands[-1][1] == 43920 and all(p.bands[i][1] == p.bands[i + 1][0] for i in range(7))
    p = scene.ScenePlan(100, 100, 256, 4, 2)          # untiled image: rank 0 does everything
    assert len(p.windows) == 0 and p.Y0 == p.Y1


def test_scene_plan_properties():
    """Host logic of the window-balanced sharding, over many scene shapes and world sizes: every window is computed
    exactly once; bands are contiguous, ordered and cover the output; every tile row cut between ranks is shipped, as
    one rectangle per (source, row), to the rank that owns the row, and those rectangles plus the owner's own windows
    cover the owner's band exactly once."""
    import importlib
    from hypothesis import given, settings, strategies as st
    scene = importlib.import_module("sentinel2-super-resolution-poc_b200.scene")

    @settings(max_examples=120, deadline=None, derandomize=True)
    @given(st.integers(20, 700), st.integers(20, 700), st.sampled_from([16, 32, 64, 100, 256]), st.integers(1, 8))
    def check(H, W, tile, world):
        plans = [scene.ScenePlan(H, W, tile, world, r) for r in range(world)]
        p0 = plans[0]
        all_w = [(w.x0, w.y0) for p in plans for w in p.windows]
        import wowsr_b200 as ws
        ref = [(w.x0, w.y0) for w in ws._lib.plan_windows(H, W, tile)]
        assert all_w == ref                                              # contiguous ranges in row-major order, no gaps
        assert p0.bands[0][0] == 0 and p0.bands[-1][1] == 4 * H
        assert all(p0.bands[i][1] == p0.bands[i + 1][0] for i in range(world - 1))
        assert all(p.pieces == p0.pieces for p in plans)                 # every rank derives the same exchange list
        for r, p in enumerate(plans):
            cover = np.zeros((p.Y1 - p.Y0, 4 * W), np.int32)
            for w in p.windows:                                          # own windows inside the own band
                y0, y1 = max(4 * w.oy0, p.Y0), min(4 * w.oy1, p.Y1)
                if y1 > y0:
                    cover[y0 - p.Y0:y1 - p.Y0, 4 * w.ox0:4 * w.ox1] += 1
            for (src, dst, y0, y1, x0, x1) in p0.pieces:
                assert src != dst
                if dst == r:
                    assert p.Y0 <= y0 and y1 <= p.Y1
                    cover[y0 - p.Y0:y1 - p.Y0, x0:x1] += 1
                if src == r:                                             # shipped rows are rows this rank computed
                    assert p.SY0 <= y0 and y1 <= p.SY1 and not (p.Y0 <= y0 and y1 <= p.Y1 and p.Y1 > p.Y0)
            assert (cover == 1).all(), (H, W, tile, world, r)

    check()
