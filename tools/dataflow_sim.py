#!/usr/bin/env python
"""Timing model of the persistent dataflow trunk kernel (csrc/trunk_kernel.cuh) — CPU only, no GPU needed.

Purpose: decide, before spending GPU minutes, which (group size, tile height, dependency granularity) can run the
residual trunk without pipeline bubbles, and how much L2 the tasks in flight pin down.  Every CTA executes its tasks
b, b + grid, b + 2 grid, ... in order (exactly the kernel's static assignment); per task

    mma_start = max(mma_end of the CTA's previous task,            # issuer busy
                    dep_ready + L_FILL,                            # dependency seen -> TMA -> first stage full
                    epi_end of the CTA's task before the previous) # TMEM accumulators are double buffered
    mma_end   = mma_start + D_mma(layer)
    epi_end   = max(mma_end, epi_end of the previous task) + D_epi(layer)
    published = epi_end + L_PUB                                    # fence + counter atomic, visible to pollers

Because a task only depends on tasks with a smaller index and every CTA walks increasing indices, evaluating tasks in
index order is an exact event simulation of this model.  Per-layer durations are the measured tile timelines of the
layer-by-layer kernel (profiles/r01_tile_timeline_final.txt, cycles): rdb.conv1 9048 MMA / 4900 epilogue, rdb.conv4
19606 / 4500, rdb.conv5 25284 / 19208 (conv2, conv3 interpolated).  R = 4 tiles of the N = 32 layers: half the rows, but
the two tile-edge rows weigh more: x 0.5625 (DESIGN.md 4.1 cost model: 288 vs 512 cycles per chunk and run-axis tap).

Calibration: `--clock` (GHz) is chosen so that the layer-by-layer model reproduces the measured 72 ms (25 windows of 276 x 276,
23 blocks, profiles/r01_dataflow_trunk_trial.txt); the same clock then has to reproduce the two measured dataflow points
(window-level dependencies: groups of 2 -> 90 ms, groups of 4 -> 74-78 ms).
"""
from __future__ import annotations

import argparse
import math

MMA8 = [9048.0, 12567.0, 16086.0, 19606.0]  # rdb.conv1..4, R = 8 tiles (N = 32)
EPI8 = [4900.0, 4800.0, 4650.0, 4500.0]
MMA5, EPI5 = 25284.0, 19208.0                # rdb.conv5, R = 4 tiles (N = 64); HBM-bound figures
R4_FACTOR = 288.0 / 512.0


def tiles(h, w, R):
    """(n_horizontal, n_vertical, bands): tile counts of one window and layer, like run_trunk_dataflow (conv.cu)."""
    wm = w // 128 * 128
    rem = w - wm
    strip = rem > 0 and wm > 0 and h >= 64
    x0 = wm if strip else w
    tiles_x = (x0 + 127) // 128
    tiles_y = (h + R - 1) // R
    n_v = ((h + 127) // 128) * ((rem + R - 1) // R) if strip else 0
    return tiles_x, tiles_y, n_v


class Model:
    def __init__(self, a):
        self.a = a
        self.R32 = a.r32
        tx8, ty8, nv8 = tiles(a.h, a.w, self.R32)
        tx4, ty4, nv4 = tiles(a.h, a.w, 4)
        self.kind = [(tx8, ty8, nv8, self.R32), (tx4, ty4, nv4, 4)]
        f32 = R4_FACTOR if self.R32 == 4 else 1.0
        self.mma = [m * f32 for m in MMA8] + [MMA5 * a.conv5_scale]
        self.epi = [e * (0.5 if self.R32 == 4 else 1.0) for e in EPI8] + [EPI5 * a.conv5_epi_scale]

    def n_tasks_layer(self, k):
        tx, ty, nv, _ = self.kind[k == 4]
        return tx * ty + nv

    def task_list(self, G):
        """One RDB period of the task order: (k, window, first_row, last_row) per task; rows = output rows the task covers."""
        out = []
        for k in range(5):
            tx, ty, nv, R = self.kind[k == 4]
            for wgi in range(G):
                per = []
                for vb in range(ty):
                    for _ in range(tx):
                        per.append((k, wgi, vb * R, min(vb * R + R, self.a.h) - 1))
                vruns = (self.a.h + 127) // 128
                vt = [[] for _ in range(vruns)]
                for i in range(nv):
                    run = i % vruns
                    vt[run].append((k, wgi, run * 128, min(run * 128 + 128, self.a.h) - 1))
                if self.a.interleave_strip:  # vertical strip tiles issued right before the horizontal rows they overlap
                    merged, vi = [], 0
                    for t in per:
                        while vi < vruns and t[2] >= vi * 128:
                            merged += vt[vi]
                            vi += 1
                        merged.append(t)
                    per = merged
                else:
                    for v in vt:
                        per += v
                out += per
        return out

    def task_list_skewed(self, G, lag):
        """One RDB over G windows as a skewed wavefront: at step s the list holds tile s of conv1, tile s - lag of conv2, ...,
        tile(s) s - 4 lag of conv5 (tile = index into the layer-major list of the layer over all G windows, so the wavefront
        sweeps window after window).  A consumer trails its producers by `lag` steps = ~6 lag tasks, which is what hides the
        publish -> poll -> TMA latency; the data it re-reads was touched at most 4 lag steps earlier."""
        full = [t for t in self.task_list(G) if t[0] in self.a.fused]
        ks = sorted(self.a.fused)
        per_layer = [[t for t in full if t[0] == k] for k in ks]
        n0 = len(per_layer[0])
        out = []
        for s in range(n0 + (len(ks) - 1) * lag):
            for i in range(len(ks)):
                u = s - i * lag
                if 0 <= u < n0:
                    nk = len(per_layer[i])
                    out += per_layer[i][u * nk // n0:(u + 1) * nk // n0]
        assert len(out) == len(full)
        return out

    def run_group(self, G, n_rdb, grid):
        """Returns (cycles until the last task is published, mean bubble cycles per task)."""
        a = self.a
        period = self.task_list_skewed(G, a.lag) if a.lag > 0 else self.task_list(G)
        band = a.band_rows
        nb = (a.h + band - 1) // band
        n_tasks = len(period) * n_rdb
        cta_mma_end = [0.0] * grid
        cta_epi_end = [0.0] * grid
        cta_epi_prev = [0.0] * grid
        # publish time per (layer index, window[, band]); "ready" needs the max over all producers
        ready_win = {}
        ready_band = {}
        count_layer = {}
        bubble = 0.0
        t_last = 0.0
        for t in range(n_tasks):
            rdb, e = divmod(t, len(period))
            k, wgi, r0, r1 = period[e]
            li = rdb * 5 + k
            cta = t % grid
            dep = 0.0
            has_dep = li > 0 and not (a.lag > 0 and k == min(a.fused))  # the first fused layer reads the previous launch's output
            if has_dep:
                if a.deps == "window":
                    dep = ready_win[(li - 1, wgi)]
                else:
                    b0, b1 = max(0, (r0 - 1) // band), min(nb - 1, (r1 + 1) // band)
                    dep = max(ready_band[(li - 1, wgi, b)] for b in range(b0, b1 + 1))
            start = max(cta_mma_end[cta], dep + a.l_fill if has_dep else 0.0, cta_epi_prev[cta])
            # dependency bubble only: the part of the issuer's idle time that neither its own previous task nor the TMEM
            # double buffer explains
            bubble += max(0.0, start - max(cta_mma_end[cta], cta_epi_prev[cta])) if t >= grid else 0.0
            mma_end = start + self.mma[k]
            epi_end = max(mma_end, cta_epi_end[cta]) + self.epi[k]
            cta_epi_prev[cta] = cta_epi_end[cta]
            cta_mma_end[cta], cta_epi_end[cta] = mma_end, epi_end
            pub = epi_end + a.l_pub
            key = (li, wgi)
            ready_win[key] = max(ready_win.get(key, 0.0), pub)
            for b in range(r0 // band, r1 // band + 1):
                kb = (li, wgi, b)
                ready_band[kb] = max(ready_band.get(kb, 0.0), pub)
            count_layer[key] = count_layer.get(key, 0) + 1
            t_last = max(t_last, pub)
        return t_last, bubble / max(1, n_tasks - grid)

    def layer_by_layer(self, n_win, n_rdb, grid, layers=range(5)):
        """One launch per layer over all windows: ceil(tiles / grid) tile periods + launch overhead."""
        total = 0.0
        for k in layers:
            n = self.n_tasks_layer(k) * n_win
            waves = n / grid  # tiles are spread evenly; the tail of the last wave is counted through the ceil below
            period = max(self.mma[k], self.epi[k]) + 350.0  # measured tile period = issue length + hand-over
            total += (math.ceil(waves) if self.a.quantise else waves) * period + self.a.launch_cycles
        return total * n_rdb


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--h", type=int, default=276)
    ap.add_argument("--w", type=int, default=276)
    ap.add_argument("--windows", type=int, default=25)
    ap.add_argument("--blocks", type=int, default=23)
    ap.add_argument("--grid", type=int, default=148)
    ap.add_argument("--clock", type=float, default=0.0, help="GHz; 0 = calibrate on the measured layer-by-layer 72 ms")
    ap.add_argument("--l-fill", dest="l_fill", type=float, default=2500.0, help="dependency visible -> first stage full (cycles)")
    ap.add_argument("--l-pub", dest="l_pub", type=float, default=800.0, help="epilogue end -> counter visible (cycles)")
    ap.add_argument("--launch-cycles", dest="launch_cycles", type=float, default=6000.0, help="per-launch drain + launch gap")
    ap.add_argument("--conv5-scale", dest="conv5_scale", type=float, default=1.0, help="conv5 MMA period scale (L2-resident: ~0.7)")
    ap.add_argument("--conv5-epi-scale", dest="conv5_epi_scale", type=float, default=1.0)
    ap.add_argument("--quantise", type=int, default=1)
    ap.add_argument("--band-rows", dest="band_rows", type=int, default=4)
    ap.add_argument("--interleave-strip", dest="interleave_strip", type=int, default=0)
    ap.add_argument("--r32", type=int, default=8)
    ap.add_argument("--deps", default="window")
    ap.add_argument("--cases", default="window:8:0,band:8:1,band:4:1", help="deps:R of the N=32 layers:interleave strip tiles, comma separated")
    ap.add_argument("--groups", default="1,2,3,4,6")
    ap.add_argument("--lag", type=int, default=0, help="> 0: skewed wavefront order, one launch per RDB over a whole group (use --groups 25)")
    ap.add_argument("--fused", default="0,1,2,3,4", help="with --lag: the layers (0 = conv1 .. 4 = conv5) of the skewed launch; the others stay one launch each")
    a = ap.parse_args()
    n_rdb = 3 * a.blocks
    a.fused = {int(v) for v in a.fused.split(",")}

    c5, c5e = a.conv5_scale, a.conv5_epi_scale
    a.conv5_scale = a.conv5_epi_scale = 1.0  # the layer-by-layer path is measured with the HBM-bound conv5
    lbl = Model(a).layer_by_layer(a.windows, n_rdb, a.grid)
    a.conv5_scale, a.conv5_epi_scale = c5, c5e
    clock = a.clock or lbl / 72e-3 / 1e9
    print(f"# window {a.h}x{a.w}, {a.windows} windows, {a.blocks} blocks, grid {a.grid}; clock {clock:.3f} GHz "
          f"({'calibrated: layer-by-layer = 72 ms' if not a.clock else 'given'})")
    print(f"layer-by-layer model: {lbl / clock / 1e6:.1f} ms")
    print(f"{'deps':8s} {'R(N=32)':8s} {'strip':6s} {'G':>2s} {'tasks/layer':>11s} {'live MB':>8s} {'ms':>7s} {'bubble/task':>11s}")
    for deps, r32, inter in [tuple(int(v) if v.isdigit() else v for v in c.split(":")) for c in a.cases.split(",")]:
        for G in [int(g) for g in a.groups.split(",")]:
            a.deps, a.r32, a.interleave_strip = deps, r32, inter
            m = Model(a)
            groups = math.ceil(a.windows / G)
            cyc = 0.0
            bub = 0.0
            for g in range(groups):
                gg = min(G, a.windows - g * G)
                if a.lag > 0:  # one skewed launch per RDB (fills and drains 69 times) + the other layers one launch each
                    c, b = m.run_group(gg, 1, a.grid)
                    c5, c5e = a.conv5_scale, a.conv5_epi_scale
                    a.conv5_scale = a.conv5_epi_scale = 1.0
                    rest = Model(a).layer_by_layer(gg, n_rdb, a.grid, [k for k in range(5) if k not in a.fused])
                    a.conv5_scale, a.conv5_epi_scale = c5, c5e
                    cyc += (c + a.launch_cycles) * n_rdb + rest
                else:
                    c, b = m.run_group(gg, n_rdb, a.grid)
                    cyc += c + a.launch_cycles
                bub += b
            live = G * a.h * a.w * 640 / 1e6  # dense 384 + next hi 128 + lo 128 B per pixel
            print(f"{deps:8s} {r32:<8d} {inter:<6d} {G:2d} {G * m.n_tasks_layer(0):11d} {live:8.1f} {cyc / clock / 1e6:7.1f} {bub / groups:11.0f}")


if __name__ == "__main__":
    main()
