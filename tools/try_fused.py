"""Round-2 queue, step 2: the EXPERIMENTAL fused-tail launch (csrc/sched_kernel.cuh, option trunk_fuse) against the
layer-by-layer path — correctness on small shapes first (each in its own bounded run), then timing on the deployment window
size, then the per-task trace of CTA 0.  Written after round 1's GPU budget was spent: this is its first hardware run.

    python tools/try_fused.py check                # 1 block: 300x290 untiled, 9 windows of 148x148, 6 windows of 276 wide
    python tools/try_fused.py perf [fuse] [lag]    # 23 blocks, 25 windows of 276x276: trunk ms, layer-by-layer vs fused
    python tools/try_fused.py trace [fuse] [lag]   # clock64 stamps of CTA 0's first 64 tasks of RDB 2
    python tools/try_fused.py fold                 # folded nearest-x2 upsample (csrc/ups_kernel.cuh): bit-identity + tail ms
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import wowsr_b200 as ws  # noqa: E402
from oracle import rrdbnet_ref as R  # noqa: E402


def net(blocks, seed):
    sd = R.calibrate_conv_last(R.random_init_state_dict(seed, blocks), blocks)
    return [sd[k + s].numpy() for k, _, _ in R.conv_specs(blocks) for s in (".weight", ".bias")]


def run(tensors, blocks, img, tile, want_float=False, reps=1, **opts):
    h = ws.Handle(0)
    for k, v in opts.items():
        h.set_option(k, v)
    h.load_rrdbnet(tensors, blocks, precision="bf16")
    out = None
    for rep in range(reps):
        t0 = time.time()
        out = h.enhance_host(img, tile, want_float=want_float)
        print("   ", opts, "rep", rep, "wall ms", round((time.time() - t0) * 1e3, 1), h.timing(), flush=True)
    trace = h.debug_trace() if opts.get("trunk_trace") else None
    h.close()
    return out, trace


def check():
    blocks = 1
    tensors = net(blocks, 4)
    for (H, W, tile) in ((300, 290, 256), (300, 290, 128), (560, 290, 256)):
        img = np.random.default_rng(13).integers(0, 256, (H, W, 3), dtype=np.uint8)
        (u8, f), _ = run(tensors, blocks, img, tile, want_float=True)
        for fuse, lag in ((4, 0), (4, 120), (3, 0), (1, 0)):
            try:
                (u8_d, f_d), _ = run(tensors, blocks, img, tile, want_float=True, trunk_fuse=fuse, trunk_lag=lag)
                print(f"  {H}x{W} tile {tile} fuse {fuse} lag {lag}: float max diff {float(np.abs(f - f_d).max()):.3e}  u8 within1 "
                      f"{float((np.abs(u8.astype(int) - u8_d.astype(int)) <= 1).mean()):.6f}  exact {float((u8 == u8_d).mean()):.6f}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"  {H}x{W} tile {tile} fuse {fuse} lag {lag}: FAILED {e}", flush=True)
                return 1
    return 0


def perf(fuse, lag):
    blocks = 23
    tensors = net(blocks, 0)
    img = np.random.default_rng(1).integers(0, 256, (1044, 1044, 3), dtype=np.uint8)   # 5 x 5 windows of 276 x 276
    print("layer-by-layer")
    a, _ = run(tensors, blocks, img, 256, reps=3)
    for f, l in ((fuse, lag), (fuse, 64), (fuse, 240), (3, lag)):
        print(f"fused tail: convs {f}..5, lag {l}")
        try:
            b, _ = run(tensors, blocks, img, 256, reps=3, trunk_fuse=f, trunk_lag=l)
            print("    u8 within1 vs layer-by-layer", float((np.abs(a.astype(int) - b.astype(int)) <= 1).mean()), flush=True)
        except Exception as e:  # noqa: BLE001
            print("    FAILED", e, flush=True)
            return 1
    return 0


def trace(fuse, lag):
    blocks = 2
    tensors = net(blocks, 0)
    img = np.random.default_rng(1).integers(0, 256, (1044, 1044, 3), dtype=np.uint8)
    _, tr = run(tensors, blocks, img, 256, trunk_fuse=fuse, trunk_lag=lag, trunk_trace=2)
    tr = tr[:64]
    t0 = tr[0, 0]
    print("CTA 0, RDB 2: task  dep_wait_begin  dep_wait_len  mma_begin-dep_end  mma_issue_len   (cycles)")
    for i, (a, b, c, d) in enumerate(tr):
        if d:
            print(f"  {i:3d} {a - t0:12d} {b - a:10d} {c - b:12d} {d - c:12d}")
    return 0


def fold():
    """Folded nearest-x2 upsample (csrc/ups_kernel.cuh, option tail_fold_upsample): same arithmetic as the product tail, so the
    outputs must be bit-identical; `tail` of the timing line is what changes."""
    for blocks, (H, W), tile in ((1, (300, 290), 256), (1, (300, 290), 128), (23, (1044, 1044), 256)):
        tensors = net(blocks, 4)
        img = np.random.default_rng(13).integers(0, 256, (H, W, 3), dtype=np.uint8)
        try:
            (u8, f), _ = run(tensors, blocks, img, tile, want_float=True, reps=2)
            (u8_d, f_d), _ = run(tensors, blocks, img, tile, want_float=True, reps=2, tail_fold_upsample=1)
        except Exception as e:  # noqa: BLE001
            print(f"  {H}x{W} tile {tile} blocks {blocks}: FAILED {e}", flush=True)
            return 1
        print(f"  {H}x{W} tile {tile} blocks {blocks}: identical u8 {bool(np.array_equal(u8, u8_d))}  identical float {bool(np.array_equal(f, f_d))}  "
              f"max float diff {float(np.abs(f - f_d).max()):.3e}", flush=True)
    return 0


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    fuse = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    lag = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    sys.exit({"check": check, "perf": lambda: perf(fuse, lag), "trace": lambda: trace(fuse, lag), "fold": fold}[mode]())
