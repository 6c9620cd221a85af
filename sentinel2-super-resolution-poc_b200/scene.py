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

import os

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
    """Window / band assignment of one rank for an H x W LR scene.

    ``balance="rows"``: whole tile rows per rank (no SR exchange; 43 tile rows over 8 ranks leaves 6 vs 5 rows,
    12 % imbalance).  ``balance="windows"`` (default): the row-major window list is cut into near-equal contiguous
    ranges; a tile row cut between two ranks is post-processed by the rank that computes its first window, the other
    rank ships its piece of that row's SR output (``pieces``) point-to-point before the CLAHE histogram."""

    def __init__(self, H, W, tile, world, rank, scale=4, pad=10, balance="windows"):
        self.H, self.W, self.tile, self.world, self.rank, self.scale = H, W, tile, world, rank, scale
        wins = _lib.plan_windows(H, W, tile, pad)
        if len(wins) == 1:
            tiles_y, tiles_x = 1, 1
        else:
            tiles_x = (W + tile - 1) // tile
            tiles_y = (H + tile - 1) // tile
        self.tiles_y, self.tiles_x = tiles_y, tiles_x
        self.OH, self.OW = H * scale, W * scale
        if balance == "rows" or len(wins) < 2 * world:
            self.row_bands = split_rows(tiles_y, world)
            self.win_ranges = [(r0 * tiles_x, r1 * tiles_x) for (r0, r1) in self.row_bands]
        else:
            self.win_ranges = split_rows(len(wins), world)
            # tile row t belongs to the rank whose range holds its first window
            self.row_bands = []
            for (k0, k1) in self.win_ranges:
                t0 = (k0 + tiles_x - 1) // tiles_x
                t1 = (k1 + tiles_x - 1) // tiles_x
                self.row_bands.append((min(t0, tiles_y), min(t1, tiles_y)))
        # output band [Y0, Y1) of every rank (output pixels)
        self.bands = []
        for (r0, r1) in self.row_bands:
            if r1 > r0:
                y0 = wins[r0 * tiles_x].oy0 * scale
                y1 = wins[(r1 - 1) * tiles_x].oy1 * scale
            else:
                y0 = y1 = (self.bands[-1][1] if self.bands else 0)
            self.bands.append((y0, y1))
        k0, k1 = self.win_ranges[rank]
        self.windows = wins[k0:k1]
        self.Y0, self.Y1 = self.bands[rank]
        # output rows this rank's windows write (a superset of its band when rows are cut between ranks)
        self.SY0 = min((w.oy0 for w in self.windows), default=0) * scale
        self.SY1 = max((w.oy1 for w in self.windows), default=0) * scale
        # SR pieces computed by one rank and post-processed by another: (src, dst, y0, y1, x0, x1) in output pixels
        self.pieces = []
        for src, (a, b) in enumerate(self.win_ranges):
            t = a // tiles_x if b > a else 0
            while b > a and t * tiles_x < b:
                c0, c1 = max(a, t * tiles_x), min(b, (t + 1) * tiles_x)
                dst = next(r for r, (r0, r1) in enumerate(self.row_bands) if r0 <= t < r1)
                piece = (src, dst, wins[c0].oy0 * scale, wins[c0].oy1 * scale, wins[c0].ox0 * scale, wins[c1 - 1].ox1 * scale)
                if dst != src and piece[3] > piece[2] and piece[5] > piece[4]:  # windows can own nothing (image < tile + 2 pad)
                    self.pieces.append(piece)
                t += 1

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


class DistComm:
    """The exchanges of ``run_scene`` over ``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def exchange(self, sends, recvs):
        """sends: [(tensor, dst)], recvs: [(tensor, src)] — one batched point-to-point round (matched in order per pair)."""
        ops = [dist.P2POp(dist.isend, t.contiguous(), dst, self.group) for t, dst in sends]
        ops += [dist.P2POp(dist.irecv, t, src, self.group) for t, src in recvs]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def all_reduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def barrier(self):
        if self.world > 1:
            dist.barrier(self.group)


class ThreadComm:
    """Same interface for `world` ranks that are THREADS of one process sharing one GPU (each with its own libwowsr handle):
    lets a single-GPU box run the sharded pipeline — cut tile rows, histogram all-reduce, seam halos, gather — against the
    single-rank output (tests/test_gpu_scene_ranks.py).  Messages between a pair of ranks are matched in order, like NCCL's."""

    class Shared:
        def __init__(self, world):
            import collections
            import threading
            self.world = world
            self.cv = threading.Condition()
            self.box = collections.defaultdict(collections.deque)   # (src, dst) -> queue of tensors
            self.red = {}
            self.red_gen = [0] * world
            self.bar = threading.Barrier(world)

    def __init__(self, shared, rank):
        self.s, self.rank, self.world, self.group = shared, rank, shared.world, None

    @staticmethod
    def _settled(t):
        """A private copy of `t` that is complete in memory (other threads read it on THEIR streams)."""
        c = t.detach().clone()
        if c.is_cuda:
            torch.cuda.current_stream(c.device).synchronize()
        return c

    def exchange(self, sends, recvs):
        msgs = [(self._settled(t), dst) for t, dst in sends]
        with self.s.cv:
            for m, dst in msgs:
                self.s.box[(self.rank, dst)].append(m)
            self.s.cv.notify_all()
        for t, src in recvs:
            with self.s.cv:
                ok = self.s.cv.wait_for(lambda: len(self.s.box[(src, self.rank)]) > 0, timeout=120)
                if not ok:
                    raise RuntimeError(f"rank {self.rank}: no message from rank {src}")
                m = self.s.box[(src, self.rank)].popleft()
            t.copy_(m)

    def all_reduce_sum(self, t):
        g = self.s.red_gen[self.rank]
        self.s.red_gen[self.rank] += 1
        mine = self._settled(t)
        with self.s.cv:
            self.s.red.setdefault(g, []).append(mine)
            self.s.cv.notify_all()
            if not self.s.cv.wait_for(lambda: len(self.s.red[g]) == self.world, timeout=120):
                raise RuntimeError(f"rank {self.rank}: all-reduce {g} incomplete")
            parts = list(self.s.red[g])
        total = parts[0].clone()
        for q in parts[1:]:
            total += q
        t.copy_(total)

    def barrier(self):
        self.s.bar.wait(timeout=120)


def run_scene(backend, img, tile, post=True, gather=True, group=None, balance="windows", comm=None):
    """One pass of the sharded pipeline.  `img`: HxWx3 uint8 tensor on the backend's device (every rank holds
    the LR scene; it is 1/16 of the output).  Returns (plan, local post-processed band, full image on rank 0
    or None).  `comm`: the exchange layer (default: torch.distributed on `group`)."""
    comm = comm if comm is not None else DistComm(group)
    world, rank = comm.world, comm.rank
    H, W = img.shape[:2]
    plan = ScenePlan(H, W, tile, world, rank, balance=balance)
    r = backend.blur_radius() if post else 0
    have = plan.Y1 > plan.Y0            # owns an output band
    have_win = len(plan.windows) > 0     # computes SR windows
    lo = max(plan.Y0 - r, 0) if have else plan.SY0
    hi = min(plan.Y1 + r, plan.OH) if have else plan.SY1
    if have_win:
        lo, hi = min(lo, plan.SY0), max(hi, plan.SY1)
    band = backend.new_band(max(hi - lo, 1), plan.OW)
    if have_win:
        backend.sr_band(img, plan, band, lo)
    # (1b) tile rows cut between two ranks: ship the SR pieces to the rank that post-processes the row
    mine = [p for p in plan.pieces if p[0] == rank or p[1] == rank]
    if mine:
        sends, recvs, pastes = [], [], []
        for (src, dst, y0, y1, x0, x1) in mine:
            if src == rank:
                sends.append((band[y0 - lo:y1 - lo, x0:x1].contiguous(), dst))
            else:
                buf = backend.new_band(y1 - y0, x1 - x0)
                recvs.append((buf, src))
                pastes.append((buf, y0, y1, x0, x1))
        comm.exchange(sends, recvs)
        for (buf, y0, y1, x0, x1) in pastes:
            band[y0 - lo:y1 - lo, x0:x1] = buf
    if not post:
        out = band[plan.Y0 - lo:plan.Y1 - lo] if have else band[:0]
    else:
        # (2) global CLAHE histogram: local partial + one all-reduce
        hist = backend.new_hist()
        if have:
            _, _, _, ph = _lib.clahe_geometry(plan.OH, plan.OW, backend.params.grid)
            last = plan.Y1 == plan.OH
            backend.hist(band, plan, lo, plan.Y0, ph if last else plan.Y1, hist)
        comm.all_reduce_sum(hist)
        luts = backend.luts(hist, plan)
        # (3) seam halo rows
        if world > 1 and r > 0:
            up, dn = plan.neighbours() if have else (None, None)
            sends, recvs = [], []
            if have and up is not None:
                sends.append((band[plan.Y0 - lo:plan.Y0 - lo + r].contiguous(), up))
                recvs.append((band[plan.Y0 - r - lo:plan.Y0 - lo], up))
            if have and dn is not None:
                sends.append((band[plan.Y1 - lo - r:plan.Y1 - lo].contiguous(), dn))
                recvs.append((band[plan.Y1 - lo:plan.Y1 + r - lo], dn))
            comm.exchange(sends, recvs)
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
                comm.exchange([], [(full[y0:y1], src) for src, (y0, y1) in enumerate(plan.bands) if src != 0 and y1 > y0])
            elif have:
                comm.exchange([(out.contiguous(), 0)], [])
    return plan, out, full


def _clear_cuda_error():
    """A failed cudaHostRegister leaves its error code as the runtime's 'last error'; the next PyTorch launch check
    would raise it.  Reading it resets it."""
    import ctypes
    for name in ("libcudart.so.12", "libcudart.so.13", "libcudart.so"):
        try:
            ctypes.CDLL(name).cudaGetLastError()
            return
        except OSError:
            continue


class SharedHostImage:
    """Host-side result buffer shared by the ranks of ONE box (POSIX shared memory).

    The NVLink gather above leaves the whole stitched image on rank 0, whose single PCIe link then carries all of it to
    the host (5.8 GB for a Sentinel-2 scene).  When the consumer is host code (the server writes PNG / GeoTIFF files),
    every rank instead copies its own band device->host into this buffer over its own PCIe link, in parallel.  Each
    rank page-locks only the rows it writes (``pin_rows``): registering the whole image in every process multiplies the
    locked-page accounting by the world size and fails for large scenes (cudaErrorOperatingSystem at 8 x 6.4 GB)."""

    def __init__(self, OH, OW, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.OH, self.OW = OH, OW
        self.nbytes = OH * OW * 3
        name = [None]
        if self.rank == 0:
            # tmpfs pages are allocated on first touch: a too-small /dev/shm would kill the writers with SIGBUS, so check
            # the free space up front and let EVERY rank fail the same way (callers fall back to the rank-0 gather)
            try:
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize >= self.nbytes + (256 << 20):
                    name = [f"/dev/shm/wowsr_{os.getpid()}_{OH}x{OW}"]
            except OSError:
                pass
        if self.world > 1:
            dist.broadcast_object_list(name, src=0, group=group)
        if name[0] is None:
            raise OSError(f"/dev/shm cannot hold a {self.nbytes >> 20} MiB image")
        self.path = name[0]
        if self.rank == 0:
            self.array = torch.from_file(self.path, shared=True, size=self.nbytes, dtype=torch.uint8)
        if self.world > 1:
            dist.barrier(group)
        if self.rank != 0:
            self.array = torch.from_file(self.path, shared=True, size=self.nbytes, dtype=torch.uint8)
        self.array = self.array.view(OH, OW, 3)
        self._range = None
        self.pinned = False

    def pin_rows(self, y0, y1):
        """Page-locks the (page-aligned) byte range of rows [y0, y1); returns False (and copies stay pageable) on failure."""
        if not torch.cuda.is_available() or y1 <= y0:
            return False
        base, row = self.array.data_ptr(), self.OW * 3
        a = (base + y0 * row) & ~4095
        b = min((base + y1 * row + 4095) & ~4095, (base + self.nbytes + 4095) & ~4095)
        if self._range == (a, b):
            return True
        self._unpin()
        rc = torch.cuda.cudart().cudaHostRegister(a, b - a, 0)
        if rc is not None and int(rc) != 0:
            _clear_cuda_error()
            self.pinned = False
            return False
        self._range, self.pinned = (a, b), True
        return True

    def _unpin(self):
        if self._range is not None:
            rc = torch.cuda.cudart().cudaHostUnregister(self._range[0])
            if rc is not None and int(rc) != 0:
                _clear_cuda_error()
            self._range, self.pinned = None, False

    def close(self):
        if getattr(self, "array", None) is None:
            return
        self._unpin()
        self.array = None
        if self.world > 1:
            dist.barrier(self.group)
        if self.rank == 0 and os.path.exists(self.path):
            os.unlink(self.path)


def lr_rows_needed(H, W, tile, world, rank, balance="windows"):
    """LR rows [y0, y1) this rank's windows read (every other row of the scene is never touched by its kernels)."""
    plan = ScenePlan(H, W, tile, world, rank, balance=balance)
    if not plan.windows:
        return 0, 0
    return min(w.y0 for w in plan.windows), max(w.y1 for w in plan.windows)


def run_scene_to_host(backend, host_img, tile, shared: SharedHostImage, post=True, balance="windows"):
    """Host-to-host pass: `host_img` (pinned HxWx3 uint8) -> device, sharded pipeline, every rank's band -> `shared`.
    Each rank uploads only the LR rows its windows read (the ranks' uploads add up to about one copy of the scene).
    On return (after the closing barrier) ``shared.array`` holds the complete image in every process.  Returns
    (plan, bytes uploaded by this rank)."""
    dev = getattr(backend, "dev", "cpu")
    H, W = host_img.shape[:2]
    y0, y1 = lr_rows_needed(H, W, tile, shared.world, shared.rank, balance)
    if shared.world == 1 or (y0, y1) == (0, H):
        d = host_img.to(dev, non_blocking=True)
        up_bytes = H * W * 3
    else:
        d = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
        if y1 > y0:
            d[y0:y1].copy_(host_img[y0:y1], non_blocking=True)
        up_bytes = (y1 - y0) * W * 3
    plan, out, _ = run_scene(backend, d, tile, post=post, gather=False, group=shared.group, balance=balance)
    if plan.Y1 > plan.Y0:
        shared.pin_rows(plan.Y0, plan.Y1)
        shared.array[plan.Y0:plan.Y1].copy_(out, non_blocking=True)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if shared.world > 1:
        dist.barrier(shared.group)
    return plan, up_bytes
