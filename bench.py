#!/usr/bin/env python
"""bench.py — output Mpix/s of the WOW super-resolution hot path (x4 RRDBNet + WOW post-process).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2|cfg1|post4096]

Workload at N=1 (BASELINE.json configs[1]): one 4096x4096 BGR uint8 image, tile_size=512, tile_pad=10 ->
64 windows of 532x532 through RRDBNet x4plus (23 RRDB, seed-0 default-init weights, bf16 operands / fp32
accumulate) -> 16384x16384 output, then the WOW post-process on that output.  At N>1 (one process per GPU,
launched by torchrun) the scene is N such images stacked vertically (weak scaling: 64 windows per GPU); the
tile rows are sharded across ranks, the CLAHE histograms are all-reduced, seam halo rows are exchanged and
the bands are gathered on rank 0 over NCCL (sentinel2-super-resolution-poc_b200/scene.py).

One JSON line on stdout (rank 0).  `value` = output Mpix/s with the input resident in HBM; `e2e` = same metric
through the public Python API with host buffers (H2D of the input from pinned memory and D2H of the result
inside the timed region).  `roofline` = all tensor-core conv launches of a step (algorithmic FLOPs per
SURVEY 8a / BASELINE.md section 3, timed with CUDA events on the launch stream inside libwowsr).
`cpu_baseline` = the oracle port (torch fp32 + cv2, oracle/) on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_LR_PX = 35_853_696
POST_BYTES_PER_PX = 9


def peaks():
    p = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "src": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(bf16_tflops=m["bf16_tflops"], bf16_tflops_sustained=m.get("bf16_tflops_sustained", m["bf16_tflops"]),
                 hbm_gbs=m["hbm_gbs"], src="measured")
    except Exception:  # noqa: BLE001
        pass
    return p


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        pw.sort()
        # power_w: median board power during the timed region — every full-size run is power-capped, so energy per step
        # (power x ms_per_step) is the quantity to compare between kernel variants
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": pw[len(pw) // 2] if pw else None}


def make_lr_image(H, W, seed=1):
    """Fallback-style Sentinel-2 distribution (up42_client.py:684-690: G~U[80,180), R,B~U[40,120)), blurred sigma=2; BGR."""
    import cv2
    import numpy as np
    rng = np.random.default_rng(seed)
    img = np.empty((H, W, 3), np.float32)
    img[..., 0] = rng.uniform(40, 120, (H, W))
    img[..., 1] = rng.uniform(80, 180, (H, W))
    img[..., 2] = rng.uniform(40, 120, (H, W))
    img = cv2.GaussianBlur(img, (0, 0), 2)
    return np.clip(img, 0, 255).astype(np.uint8)


def workload(name, n_gpus):
    if name == "cfg2":
        return dict(H=4096 * n_gpus, W=4096, tile=512, post=True, label=f"cfg2: {64*n_gpus} windows of 532x532 (4096x{4096*n_gpus} BGR u8, tile_size=512, tile_pad=10), "
                    "RRDBNet x4plus + WOW post-process")
    if name == "cfg2s":
        return dict(H=1200 * n_gpus, W=1200, tile=512, post=True, label="cfg2s (profiling subset): 9 windows of 532x532 per GPU, RRDBNet x4plus + WOW post-process")
    if name == "cfg1":
        return dict(H=128 * n_gpus, W=128, tile=256, post=True, label="cfg1: one 128x128 tile untiled, RRDBNet x4plus + WOW post-process")
    if name == "post4096":
        return dict(H=1024 * n_gpus, W=1024, tile=0, post=True, label="cfg4: WOW post-process only on a 4096x4096 RGB image")
    if name == "scene":
        return dict(H=10980, W=10980, tile=256, post=True, label="cfg5: 10980x10980 scene, 1849 windows of 276x276, RRDBNet x4plus + WOW post-process (strong scaling)")
    raise SystemExit(f"unknown workload {name}")


def window_flops(H, W, tile):
    import wowsr_b200 as ws
    wins = ws._lib.plan_windows(H, W, tile)
    return sum((w.x1 - w.x0) * (w.y1 - w.y0) for w in wins) * FLOP_PER_LR_PX, len(wins)


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------------

def cpu_sample(wl, blocks=23, sample_px=(266, 266)):
    """Times the CPU oracle (torch fp32 RRDBNet + cv2 post-process) on one bounded sample; returns
    (output Mpix/s, description, threads)."""
    import numpy as np
    import torch

    from oracle import rrdbnet_ref as R
    from oracle import wow_cv2
    import cv2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cv2.setNumThreads(cores)
    sd = R.random_init_state_dict(0, blocks)
    img = make_lr_image(sample_px[0], sample_px[1], seed=1)
    t0 = time.perf_counter()
    if wl["tile"] > 0:
        sr = R.enhance(sd, img, blocks, tile_size=max(wl["tile"], 1 << 14))       # one window, untiled
        sr = np.ascontiguousarray(sr[:, :, ::-1])
    else:
        sr = np.ascontiguousarray(np.repeat(np.repeat(img, 4, 0), 4, 1))
    if wl["post"]:
        wow_cv2.enhance_for_crops(sr)
    dt = time.perf_counter() - t0
    mpix = sr.shape[0] * sr.shape[1] / 1e6 / dt
    return mpix, f"one {sample_px[0]}x{sample_px[1]} LR window (of the workload's windows), network + post-process, {dt:.1f} s", cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload, args.gpus)
    vals = []
    times = []
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        # each step = one bounded sample; on slow hosts shrink the sample so the whole run stays within minutes
        px = (448, 448) if time.perf_counter() - t_start < 60 else (128, 128)
        t_s = time.perf_counter()
        v, desc, cores = cpu_sample(wl, sample_px=px if args.workload != "post4096" else (1024, 1024))
        if i >= args.warmup:
            vals.append(v)
            times.append((time.perf_counter() - t_s) * 1e3)
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": "output Mpix/s (x4 RRDBNet + WOW post-process)", "value": value, "unit": "Mpix/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(times) / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "note": "oracle port of the reference CPU path (torch fp32 + cv2); each step = one bounded sample"},
            "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------

def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import wowsr_b200 as ws
    scene = __import__("importlib").import_module("sentinel2-super-resolution-poc_b200.scene")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = workload(args.workload, world)
    H, W, tile = wl["H"], wl["W"], wl["tile"]
    blocks = 23
    # seed-0 PyTorch default init, drawn through the package's own parameter container: it creates its convs in the reference's
    # construction order, so the RNG stream (and every weight) equals the reference class's — pinned by the reference's golden
    # checksums in tests/test_oracle_rrdbnet.py.  Nothing under oracle/ is touched by this arm.
    torch.manual_seed(0)
    sd = ws.app.cnn_super_resolution.RRDBNet(3, 3, 64, blocks, 32, 4).state_dict()
    handle = None
    if args.opt:  # options first: some (trunk_fuse, trunk_dataflow, tc_chunk32) decide how the weights are packed at load time
        handle = ws.Handle(local)
        for kv in args.opt:
            k, v = kv.split("=")
            handle.set_option(k, int(v))
    up = ws.app.cnn_super_resolution.RealESRGAN(scale=4, device=f"cuda:{local}", tile_size=max(tile, 1), state_dict=sd, precision=args.precision,
                                                handle=handle)
    params = ws._lib.post_params("wow")
    backend = scene.GpuBackend(up, params)

    if args.workload == "post4096":
        host_img = make_lr_image(H, W, seed=3)
    else:
        host_img = make_lr_image(H, W, seed=1)
    pinned = torch.from_numpy(host_img).pin_memory()
    dimg = pinned.to(dev)
    OH, OW = 4 * H, 4 * W
    flops, n_windows = window_flops(H, W, tile) if tile > 0 else (0, 0)

    sr_resident = post_out = None
    if tile == 0:  # post-process-only workload: the "SR output" (a nearest-x4 of the input) is resident before the timed region
        sr_resident = dimg.repeat_interleave(4, 0).repeat_interleave(4, 1).contiguous()
        post_out = torch.empty_like(sr_resident)

    def step_device():
        if tile > 0:
            return scene.run_scene(backend, dimg, tile, post=wl["post"], gather=True)
        up._h.post_process_dev(sr_resident.data_ptr(), post_out.data_ptr(), sr_resident.shape[0], sr_resident.shape[1], params,
                               stream=torch.cuda.current_stream().cuda_stream)
        return None, post_out, None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    h = up._h
    launches0 = h.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    conv_ms = []
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
        if tile > 0:
            t = h.timing()
            conv_ms.append(t["trunk"] + t["tail"])
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = (h.launch_count() - launches0) / args.steps
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    out_mpix = OH * OW / 1e6
    value = out_mpix / (ms / 1e3)

    # end-to-end through the public API with host buffers (rank-local scene at N>1 is the same call)
    e2e = None
    if args.no_e2e:
        pass
    elif tile > 0:
        shared = None
        if world > 1:
            try:  # every rank copies its band to the host over its own PCIe link (scene.SharedHostImage)
                shared = scene.SharedHostImage(OH, OW)
            except Exception as e:  # noqa: BLE001  (no /dev/shm space, ...): fall back to gather + rank-0 copy
                print(f"[bench] shared host image unavailable ({e}); using the rank-0 gather path", file=sys.stderr)
                shared = None
            ok = torch.tensor([1 if shared is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0 and shared is not None:
                shared.close()
                shared = None
        out_host = torch.empty((OH, OW, 3), dtype=torch.uint8).pin_memory() if (rank == 0 and shared is None) else None

        def step_e2e():
            if shared is not None:
                scene.run_scene_to_host(backend, pinned, tile, shared, post=wl["post"])   # H2D, pipeline, per-rank D2H
                return
            d = pinned.to(dev, non_blocking=True)                       # H2D of this step's input
            _, _, full = scene.run_scene(backend, d, tile, post=wl["post"], gather=True)
            if rank == 0:
                out_host.copy_(full, non_blocking=True)                   # D2H of the stitched result
            torch.cuda.synchronize()

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            step_e2e()
        barrier()
        e_ms = (time.perf_counter() - t0) / n_e2e * 1e3
        t_e = torch.tensor([e_ms], device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": out_mpix / (float(t_e.item()) / 1e3), "unit": "Mpix/s", "h2d_bytes_per_step": int(H * W * 3),
               "d2h_bytes_per_step": int(OH * OW * 3), "ms_per_step": float(t_e.item()),
               "api": ("RealESRGAN + scene.run_scene_to_host: pinned host input on every rank, every rank copies its band into one "
                       "shared page-locked host image" if shared is not None else
                       "RealESRGAN + scene.run_scene (enhance -> _enhance_for_crops) with pinned host buffers")}
        if shared is not None:
            shared.close()
    else:
        p = ws._lib.post_params("wow")
        src = np.ascontiguousarray(np.repeat(np.repeat(host_img, 4, 0), 4, 1))
        h.post_process_host(src, p)
        t0 = time.perf_counter()
        for _ in range(3):
            h.post_process_host(src, p)
        e_ms = (time.perf_counter() - t0) / 3 * 1e3
        e2e = {"value": out_mpix / (e_ms / 1e3), "unit": "Mpix/s", "h2d_bytes_per_step": int(src.nbytes), "d2h_bytes_per_step": int(src.nbytes),
               "ms_per_step": e_ms, "api": "_enhance_for_crops(np.ndarray)"}

    # The literal reference call sequence with numpy arrays (wow_sr.py:94-110): enhance(host) -> cvtColor -> _enhance_for_crops(host).
    # Two host round trips through pageable memory; reported next to `e2e.value` (which keeps the SR image on the GPU between the
    # two stages, like apply_wow_sr does), never instead of it.  Guarded: a failure here must not cost the bench line.
    if e2e is not None and tile > 0 and world == 1:
        try:
            import cv2

            def step_dropin():
                sr = up.enhance(host_img)
                rgb = cv2.cvtColor(sr, cv2.COLOR_BGR2RGB)
                return ws.app.wow_sr._enhance_for_crops(rgb) if wl["post"] else rgb

            step_dropin()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(2):
                step_dropin()
            d_ms = (time.perf_counter() - t0) / 2 * 1e3
            e2e["dropin_two_calls"] = {"value": out_mpix / (d_ms / 1e3), "unit": "Mpix/s", "ms_per_step": d_ms,
                                       "api": "RealESRGAN.enhance(ndarray) -> cv2.cvtColor -> _enhance_for_crops(ndarray), pageable host arrays"}
        except Exception as e:  # noqa: BLE001
            e2e["dropin_two_calls"] = {"error": str(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    if tile > 0:
        conv_t = sum(conv_ms) / len(conv_ms) / 1e3
        per_rank_flops = flops / world
        ach = per_rank_flops / conv_t / 1e12
        traffic = None
        try:  # dram__bytes_read+write of the 350 launches from the committed ncu pass (9 windows of 532x532), scaled by window pixels
            with open(os.path.join(ROOT, "profiles", "r01_cfg2s_dram_traffic.json")) as f:
                t = json.load(f)
            traffic = t["dram_bytes"] / (t["windows"] * 532 * 532) * (flops / world / FLOP_PER_LR_PX)
        except Exception:  # noqa: BLE001
            pass
        roof = {"bound": "tensor", "kernel": "conv3x3_tc_kernel (all 350 launches of a step, rank 0)", "achieved": ach,
                "peak": pk["bf16_tflops_sustained"], "peak_burst": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops_sustained"], "frac_of_burst": ach / pk["bf16_tflops"], "peak_source": pk["src"],
                "traffic": traffic, "traffic_note": "DRAM bytes per step of these launches (ncu, profiles/r01_cfg2s_dram_traffic.json, scaled per window pixel)",
                "conv_ms_per_step": conv_t * 1e3, "algorithmic_flops_per_step": per_rank_flops}
    else:
        ach = OH * OW * POST_BYTES_PER_PX / (ms / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": "clahe_hist + post_apply", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "peak_source": pk["src"], "traffic": None}
    cpu = None
    if world == 1 and not args.no_cpu:
        v, desc, cores = cpu_sample(wl, sample_px=(448, 448) if tile > 0 else (1024, 1024))
        cpu = {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": desc}
    line = {"metric": "output Mpix/s (x4 RRDBNet + WOW post-process)", "value": value, "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if args.workload == "scene" else "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": wl["label"], "windows": n_windows, "weights": "seed-0 PyTorch default init (random-init, no network)",
                       "l2": "working set (GBs of activations per step) is far larger than the 126 MB L2; no explicit flush",
                       "parallelism": f"contiguous window ranges (near-equal counts) over {world} GPU(s); cut tile rows exchanged P2P"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiler passes only; such a line is not a bench value)")
    ap.add_argument("--opt", action="append", default=[], help="libwowsr option key=value")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
