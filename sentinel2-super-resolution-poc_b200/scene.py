"""Sharding a scene across the GPUs of one box (BASELINE config 5 / SURVEY 8e).

The reference has no distributed code; its unit of independence is the tile window of
``RealESRGAN._tile_process`` (cnn_super_resolution.py:236-280).  One process per GPU:

1. the ``tiles_y x tiles_x`` window grid is split into contiguous tile-row bands, one per rank; each rank
   runs RRDBNet on its windows and owns the output rows those windows own (last-writer-wins resolved by
   the planner, so bands are disjoint and contiguous);
2. CLAHE is global: every rank histograms its band, ONE all-reduce (grid*grid*256 counters) gives every
   rank the same LUTs;
3. the unsharp blur needs ``r`` rows of the neighbouring bands' SR output: exchanged point-to-point;
4. every rank post-processes its band; bands are gathered on rank 0.

The compute is pluggable (``GpuBackend`` calls libwowsr; the tests use an oracle-backed CPU backend over
gloo to check the decomposition itself).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def split_rows(n: int, world: int):
    """Contiguous near-equal split of n tile rows over `world` ranks: list of (r0, r1)."""
    base, rem = divmod(n, world)
    out, r = [], 0
    for i in range(world):
        k = base + (1 if i < rem else 0)
        out.append((r, r + k))
        r += k
    return out


class ScenePlan:
    """Window / band assignment of one rank for an H x W LR scene."""

    def __init__(self, H, W, tile, world, rank, scale=4, pad=10):
        self.H, self.W, self.tile, self.world, self.rank, self.scale = H, W, tile, world, rank, scale
        wins = _lib.plan_windows(H, W, tile, pad)
        if len(wins) == 1:
            tiles_y, tiles_x = 1, 1
        else:
            tiles_x = (W + tile - 1) // tile
            tiles_y = (H + tile - 1) // tile
        self.tiles_y, self.tiles_x = tiles_y, tiles_x
        self.row_bands = split_rows(tiles_y, world)
        # output band [Y0, Y1) of every rank (output pixels)
        self.bands = []
        for (r0, r1) in self.row_bands:
            if r1 > r0:
                y0 = wins[r0 * tiles_x].oy0 * scale
                y1 = wins[(r1 - 1) * tiles_x].oy1 * scale
            else:
                y0 = y1 = (self.bands[-1][1] if self.bands else 0)
            self.bands.append((y0, y1))
        r0, r1 = self.row_bands[rank]
        self.windows = wins[r0 * tiles_x:r1 * tiles_x]
        self.Y0, self.Y1 = self.bands[rank]
        self.OH, self.OW = H * scale, W * scale

    def neighbours(self):
        """Ranks holding the band directly above / below this one (skipping empty bands)."""
        up = next((r for r in range(self.rank - 1, -1, -1) if self.bands[r][1] > self.bands[r][0]), None)
        dn = next((r for r in range(self.rank + 1, self.world) if self.bands[r][1] > self.bands[r][0]), None)
        return up, dn


class GpuBackend:
    """libwowsr-backed compute on this rank's GPU."""

    def __init__(self, upsampler, params):
        self.up = upsampler
        self.h = upsampler._h
        self.params = params
        self.dev = upsampler.device

    def blur_radius(self):
        taps = _lib.gaussian_taps(self.params.sigma)
        half = len(taps) // 2
        r = half
        while r > 0 and taps[half - r] == 0:
            r -= 1
        return r

    def sr_band(self, img_dev, plan, band, row_off):
        """Runs the plan's windows; output row Y lands in band[Y - row_off]."""
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        pitch = plan.OW * 3
        self.h.forward_windows(img_dev.data_ptr(), plan.H, plan.W, plan.W * 3, plan.windows, band.data_ptr() - row_off * pitch,
                               pitch, stream=stream)

    def hist(self, band, plan, row_off, prow0, prow1, hist):
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        im = _lib.Image(band.data_ptr(), plan.OW * 3, plan.OW, plan.OH, row_off, band.shape[0])
        self.h.clahe_hist(im, self.params.grid, prow0, prow1, hist.data_ptr(), stream=stream)

    def luts(self, hist, plan):
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        tw, th, _, _ = _lib.clahe_geometry(plan.OH, plan.OW, self.params.grid)
        luts = torch.empty(self.params.grid ** 2 * 256, dtype=torch.uint8, device=self.dev)
        self.h.clahe_luts(hist.data_ptr(), self.params.grid, tw, th, self.params.clip_limit, luts.data_ptr(), stream=stream)
        return luts

    def apply(self, band, plan, row_off, luts, out):
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        src = _lib.Image(band.data_ptr(), plan.OW * 3, plan.OW, plan.OH, row_off, band.shape[0])
        dst = _lib.Image(out.data_ptr(), plan.OW * 3, plan.OW, plan.OH, plan.Y0, out.shape[0])
        self.h.post_apply(src, luts.data_ptr(), self.params, plan.Y0, plan.Y1, dst, stream=stream)

    def new_band(self, rows, width):
        return torch.empty((rows, width, 3), dtype=torch.uint8, device=self.dev)

    def new_hist(self):
        return torch.zeros(self.params.grid ** 2 * 256, dtype=torch.int32, device=self.dev)


def run_scene(backend, img, tile, post=True, gather=True, group=None):
    """One pass of the sharded pipeline.  `img`: HxWx3 uint8 tensor on the backend's device (every rank holds
    the LR scene; it is 1/16 of the output).  Returns (plan, local post-processed band, full image on rank 0
    or None)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    H, W = img.shape[:2]
    plan = ScenePlan(H, W, tile, world, rank)
    r = backend.blur_radius() if post else 0
    have = plan.Y1 > plan.Y0
    lo = max(plan.Y0 - r, 0) if have else 0
    hi = min(plan.Y1 + r, plan.OH) if have else 0
    band = backend.new_band(max(hi - lo, 1), plan.OW)
    if have:
        backend.sr_band(img, plan, band, lo)
    if not post:
        out = band[plan.Y0 - lo:plan.Y1 - lo] if have else band[:0]
    else:
        # (2) global CLAHE histogram: local partial + one all-reduce
        hist = backend.new_hist()
        if have:
            _, _, _, ph = _lib.clahe_geometry(plan.OH, plan.OW, backend.params.grid)
            last = plan.Y1 == plan.OH
            backend.hist(band, plan, lo, plan.Y0, ph if last else plan.Y1, hist)
        if world > 1:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        luts = backend.luts(hist, plan)
        # (3) seam halo rows
        if world > 1 and r > 0:
            up, dn = plan.neighbours() if have else (None, None)
            ops = []
            if have and up is not None:
                ops.append(dist.P2POp(dist.isend, band[plan.Y0 - lo:plan.Y0 - lo + r].contiguous(), up, group))
                ops.append(dist.P2POp(dist.irecv, band[0:plan.Y0 - lo], up, group))
            if have and dn is not None:
                ops.append(dist.P2POp(dist.isend, band[plan.Y1 - lo - r:plan.Y1 - lo].contiguous(), dn, group))
                ops.append(dist.P2POp(dist.irecv, band[plan.Y1 - lo:hi - lo], dn, group))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        out = backend.new_band(max(plan.Y1 - plan.Y0, 1), plan.OW)
        if have:
            backend.apply(band, plan, lo, luts, out)
        out = out[:plan.Y1 - plan.Y0]
    full = None
    if gather:
        if world == 1:
            full = out
        else:
            # (4) gather the uint8 bands on rank 0 (NVLink point-to-point)
            if rank == 0:
                full = backend.new_band(plan.OH, plan.OW)
                full[plan.Y0:plan.Y1] = out
                reqs = [dist.irecv(full[y0:y1], src, group=group) for src, (y0, y1) in enumerate(plan.bands) if src != 0 and y1 > y0]
                for q in reqs:
                    q.wait()
            elif have:
                dist.send(out.contiguous(), 0, group=group)
    return plan, out, full
