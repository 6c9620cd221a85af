#!/usr/bin/env python
"""bench.py — output Mpix/s of the WOW super-resolution hot path (x4 RRDBNet + WOW post-process).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload scene|cfg2|cfg1|cfg3|post4096]

Default workload, for every N (BASELINE.json configs[4], the configuration the metric and the north_star target are
quoted on; it fits one GPU): a full 10980x10980 BGR uint8 Sentinel-2 scene, tile_size=256, tile_pad=10 -> 1849 windows
of 276x276 through RRDBNet x4plus (23 RRDB, seed-0 default-init weights, bf16 operands / fp32 accumulate) -> 43920x43920
output, then the WOW post-process on that output.  STRONG scaling: at N>1 (one process per GPU, launched by torchrun)
the same scene is sharded — contiguous window ranges per rank, cut tile rows exchanged point-to-point, the CLAHE
histograms all-reduced, seam halo rows exchanged, bands gathered over NCCL (sentinel2-super-resolution-poc_b200/scene.py).
`--workload cfg2` = BASELINE configs[1] (64 windows of 532x532 per GPU, weak scaling), cfg1 = configs[0], cfg3 = configs[2]
(EDSR-baseline x4 on a 1024x1024 tile), post4096 = configs[3].

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
EDSR_FLOP_PER_LR_PX = 3_966_336
POST_BYTES_PER_PX = 9
TRAFFIC_FILES = {532: "r02_dram_traffic.json", 276: "r02_dram_traffic_276.json"}   # by window side; written by tools/launch_report.py from the committed ncu launch lists


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


def sample_short_region(sampler, step, world, min_samples=3, seconds=1.5):
    """A timed region shorter than a few nvidia-smi periods (cfg1: 0.1 s, post4096: 11 ms) leaves the sampler nothing to read.
    Single-GPU runs then keep the GPU under the SAME steps, untimed, until a few samples are in; the clocks object says so."""
    import torch
    if world != 1 or sampler.proc is None or len(sampler.lines) >= min_samples:
        return None
    t_end = time.time() + seconds
    while time.time() < t_end and len(sampler.lines) < min_samples:
        step()
    torch.cuda.synchronize()
    return "timed region shorter than the sampling period: sampled under an untimed continuation of the same steps"


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
        return dict(H=4096 * n_gpus, W=4096, tile=512, post=True, scaling="weak",
                    label=f"cfg2: {64*n_gpus} windows of 532x532 (4096x{4096*n_gpus} BGR u8, tile_size=512, tile_pad=10), RRDBNet x4plus + WOW post-process")
    if name == "cfg2s":
        return dict(H=1200 * n_gpus, W=1200, tile=512, post=True, scaling="weak",
                    label="cfg2s (profiling subset): 9 windows of 532x532 per GPU, RRDBNet x4plus + WOW post-process")
    if name == "cfg5s":
        return dict(H=1280, W=1280, tile=256, post=True, scaling="strong",
                    label="cfg5s (profiling subset): 25 windows of 276x276, RRDBNet x4plus + WOW post-process")
    if name == "cfg1":
        return dict(H=128 * n_gpus, W=128, tile=256, post=True, scaling="weak",
                    label="cfg1: one 128x128 tile untiled, RRDBNet x4plus + WOW post-process")
    if name == "cfg3":
        return dict(H=1024, W=1024, tile=-1, post=False, scaling="replicas",
                    label="cfg3: EDSR-baseline x4 (16 resblocks, 64 features; parity unpinned) on one 1024x1024 BGR u8 tile, host in / host out")
    if name == "post4096":
        return dict(H=1024 * n_gpus, W=1024, tile=0, post=True, scaling="weak", label="cfg4: WOW post-process only on a 4096x4096 RGB image")
    if name == "scene":
        return dict(H=10980, W=10980, tile=256, post=True, scaling="strong",
                    label="cfg5: 10980x10980 scene, 1849 windows of 276x276 (tile_size=256, tile_pad=10), RRDBNet x4plus + WOW post-process, strong scaling")
    raise SystemExit(f"unknown workload {name}")


def window_flops(H, W, tile):
    import wowsr_b200 as ws
    wins = ws._lib.plan_windows(H, W, tile)
    return sum((w.x1 - w.x0) * (w.y1 - w.y0) for w in wins) * FLOP_PER_LR_PX, len(wins)


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------------

_CPU_STATE = {}


def cpu_sample(wl, blocks=23, n_windows=2, budget_s=None, window_px=None, post_px=None):
    """The reference's CPU path (oracle port: torch fp32 RRDBNet + cv2 post-process) on the host cores, on a bounded sample of
    THIS workload: `n_windows` windows of the workload's own window size through the network (one untimed warm-up window the
    first time), extrapolated linearly to the workload's window count, plus the post-process timed on
    a <= 4096x4096 image and extrapolated per pixel.  Returns (output Mpix/s of the whole job, description, threads).
    `window_px` / `post_px` shrink the sample (tests/test_bench_contract.py only)."""
    import cv2
    import numpy as np
    import torch

    from oracle import rrdbnet_ref as R
    from oracle import wow_cv2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cv2.setNumThreads(cores)
    H, W, tile = wl["H"], wl["W"], wl["tile"]
    OUT = 16.0 * H * W
    if tile == -1:  # EDSR
        from oracle import edsr_ref as E
        sd = _CPU_STATE.setdefault("edsr_sd", E.random_init_state_dict(0))
        img = make_lr_image(256, 256, seed=2)
        if "edsr_warm" not in _CPU_STATE:
            E.upsample(sd, img[:64, :64])
            _CPU_STATE["edsr_warm"] = True
        t0 = time.perf_counter()
        E.upsample(sd, img)
        dt = time.perf_counter() - t0
        total = dt * (H * W) / (256.0 * 256.0)
        return OUT / 1e6 / total, f"one 256x256 LR crop of the 1024x1024 tile through the EDSR restatement ({dt:.1f} s), extrapolated per pixel", cores
    t_net, desc = 0.0, []
    if tile > 0:
        import wowsr_b200 as ws
        wins = ws._lib.plan_windows(H, W, tile)
        wh, ww = window_px or (wins[0].y1 - wins[0].y0, wins[0].x1 - wins[0].x0)
        sd = _CPU_STATE.setdefault(("sd", blocks), R.random_init_state_dict(0, blocks))
        win = make_lr_image(wh, ww, seed=1)
        if "warm" not in _CPU_STATE:  # first call: thread pools, oneDNN primitive caches
            t0 = time.perf_counter()
            R.enhance(sd, win, blocks, tile_size=1 << 14)
            _CPU_STATE["warm"] = time.perf_counter() - t0
        if budget_s is not None and _CPU_STATE["warm"] * n_windows > budget_s:
            n_windows = 1
        t0 = time.perf_counter()
        for _ in range(n_windows):
            R.enhance(sd, win, blocks, tile_size=1 << 14)      # one window, untiled: what _tile_process runs per window
        per_win = (time.perf_counter() - t0) / n_windows
        t_net = per_win * len(wins)
        desc.append(f"{n_windows} window(s) of {wh}x{ww} through the network at {per_win:.2f} s each, extrapolated linearly to {len(wins)} windows")
    t_post = 0.0
    if wl["post"]:
        ph, pw = post_px or (min(4 * H, 4096), min(4 * W, 4096))
        key = ("post", ph, pw)
        if key not in _CPU_STATE:
            lr = make_lr_image((ph + 3) // 4, (pw + 3) // 4, seed=3)
            _CPU_STATE[key] = np.ascontiguousarray(np.repeat(np.repeat(lr, 4, 0), 4, 1)[:ph, :pw])
            wow_cv2.enhance_for_crops(_CPU_STATE[key][:256, :256])
        t0 = time.perf_counter()
        wow_cv2.enhance_for_crops(_CPU_STATE[key])
        dt = time.perf_counter() - t0
        t_post = dt * OUT / (ph * pw)
        desc.append(f"_enhance_for_crops on {ph}x{pw} in {dt:.2f} s" + ("" if ph * pw == OUT else ", extrapolated per pixel"))
    return OUT / 1e6 / (t_net + t_post), "; ".join(desc), cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload, args.gpus)
    vals, times = [], []
    n_steps = args.warmup + args.steps
    for i in range(n_steps):
        t_s = time.perf_counter()
        # each step = one bounded sample of the workload; two windows per step unless that would take the run past ~4 minutes
        v, desc, cores = cpu_sample(wl, n_windows=2, budget_s=240.0 / n_steps)
        if i >= args.warmup:
            vals.append(v)
            times.append((time.perf_counter() - t_s) * 1e3)
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": "output Mpix/s (x4 RRDBNet + WOW post-process)", "value": value, "unit": "Mpix/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(times) / len(times), "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "note": "oracle port of the reference CPU path (torch fp32 + cv2) on the host cores; each step = "
                       "one bounded sample of this workload, value = whole-job output Mpix/s extrapolated from it"},
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
    if args.opt:  # options first: some (tc_chunk32) decide how the weights are packed at load time
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
    conv_ms, enq_ms = [], []
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
        if tile > 0:
            t = h.timing()
            conv_ms.append(t["head"] + t["trunk"] + t["tail"])   # conv_first + RRDB trunk + HR tail: every conv launch of the step
            enq_ms.append(t.get("enqueue_host", 0.0))            # host time of the thread that enqueued them
    ev1.record()
    barrier()
    launches = (h.launch_count() - launches0) / args.steps
    clock_note = sample_short_region(sampler, step_device, world) if rank == 0 else None
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None and clock_note:
        clocks["note"] = clock_note
    ms = ev0.elapsed_time(ev1) / args.steps
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

        up_bytes = [H * W * 3]

        def step_e2e():
            if shared is not None:
                _, up_bytes[0] = scene.run_scene_to_host(backend, pinned, tile, shared, post=wl["post"])   # H2D, pipeline, per-rank D2H
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
        t_up = torch.tensor([float(up_bytes[0])], device=dev, dtype=torch.float64)   # bytes uploaded, summed over the ranks
        if world > 1:
            dist.all_reduce(t_up, op=dist.ReduceOp.SUM)
        e2e = {"value": out_mpix / (float(t_e.item()) / 1e3), "unit": "Mpix/s", "h2d_bytes_per_step": int(t_up.item()),
               "d2h_bytes_per_step": int(OH * OW * 3), "ms_per_step": float(t_e.item()),
               "api": ("RealESRGAN + scene.run_scene_to_host: pinned host input, every rank uploads the LR rows its windows read and copies "
                       "its band into one shared page-locked host image" if shared is not None else
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
        traffic, traffic_note = None, "not measured on this step"
        try:  # dram__bytes_read+write of the conv launches from the committed ncu launch list (a subset of this workload's windows)
            wins0 = ws._lib.plan_windows(H, W, tile)
            TRAFFIC_FILE = TRAFFIC_FILES[wins0[0].x1 - wins0[0].x0]
            with open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)) as f:
                t = json.load(f)
            traffic = t["dram_bytes"] / (t["windows"] * t.get("side", 532) ** 2) * (flops / world / FLOP_PER_LR_PX)
            traffic_note = (f"SCALED, not measured on this step: DRAM bytes of the conv launches of one ncu-profiled step on {t['windows']} windows of "
                            f"{t.get('side', 532)}x{t.get('side', 532)} (profiles/{TRAFFIC_FILE}), per window pixel x this step's window pixels")
        except Exception:  # noqa: BLE001
            pass
        roof = {"bound": "tensor", "kernel": "conv3x3_roll_kernel / conv3x3_tc_ups_kernel (all 350 tensor-core conv launches of a step + conv_first, rank 0)",
                "achieved": ach, "peak": pk["bf16_tflops_sustained"], "peak_burst": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops_sustained"], "frac_of_burst": ach / pk["bf16_tflops"], "peak_source": pk["src"],
                "traffic": traffic, "traffic_note": traffic_note,
                "conv_ms_per_step": conv_t * 1e3, "algorithmic_flops_per_step": per_rank_flops,
                "host_enqueue_ms_per_step": sum(enq_ms) / max(len(enq_ms), 1)}
    else:
        ach = OH * OW * POST_BYTES_PER_PX / (ms / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": "clahe_hist + post_apply", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "peak_source": pk["src"], "traffic": None}
    cpu = None
    if world == 1 and not args.no_cpu:
        v, desc, cores = cpu_sample(wl, n_windows=2)
        cpu = {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": desc}
    line = {"metric": "output Mpix/s (x4 RRDBNet + WOW post-process)", "value": value, "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": wl["label"], "windows": n_windows, "weights": "seed-0 PyTorch default init (random-init, no network)",
                       "l2": "working set (GBs of activations per step) is far larger than the 126 MB L2; no explicit flush",
                       "parallelism": f"contiguous window ranges (near-equal counts) over {world} GPU(s); cut tile rows exchanged P2P"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_edsr(args):
    """cfg3: EDSR-baseline x4 on one 1024x1024 tile (the /api/sr "farm SR" variant, super_resolution.py:196).  One replica per
    rank (the path does not shard: a single tile); value = N x per-rank throughput."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import wowsr_b200 as ws
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = workload("cfg3", world)
    H, W = wl["H"], wl["W"]
    # seeded PyTorch default init in the key order of app.super_resolution.edsr_keys (resblock outputs damped x0.1 like the
    # self-consistency tests); nothing under oracle/ is touched by this arm
    torch.manual_seed(0)
    sd = {}
    keys = ws.app.super_resolution.edsr_keys(16)
    for k in keys:
        cin, cout = (3, 64) if k == "head" else (64, 256) if k in ("up1", "up2") else (64, 3) if k == "tail" else (64, 64)
        conv = torch.nn.Conv2d(cin, cout, 3, 1, 1)
        g = 0.1 if k.endswith("conv2") else 1.0
        sd[k + ".weight"], sd[k + ".bias"] = conv.weight.detach() * g, conv.bias.detach() * g
    sr = ws.app.super_resolution.EdsrSuperRes(sd, device=local, precision=args.precision)
    h = sr._h
    for kv in args.opt:
        k, v = kv.split("=")
        h.set_option(k, int(v))
    host_img = make_lr_image(H, W, seed=2)
    dimg = torch.from_numpy(host_img).to(dev)
    out = torch.empty((4 * H, 4 * W, 3), dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        h.edsr_upsample_dev(dimg.data_ptr(), H, W, out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = h.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    conv_ms = []
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
        conv_ms.append(h.timing()["total"])
    ev1.record()
    barrier()
    launches = (h.launch_count() - launches0) / args.steps
    clock_note = sample_short_region(sampler, step, world) if rank == 0 else None
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None and clock_note:
        clocks["note"] = clock_note
    ms = ev0.elapsed_time(ev1) / args.steps
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    out_mpix = 16.0 * H * W / 1e6 * world
    e2e = None
    if not args.no_e2e:
        sr.upsample(host_img)
        barrier()
        t0 = time.perf_counter()
        n_e = max(1, min(args.steps, 5))
        for _ in range(n_e):
            sr.upsample(host_img)                                  # the call the reference makes: host BGR in, host BGR out
        barrier()
        t_e = torch.tensor([(time.perf_counter() - t0) / n_e * 1e3], device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": out_mpix / (float(t_e.item()) / 1e3), "unit": "Mpix/s", "h2d_bytes_per_step": int(H * W * 3) * world,
               "d2h_bytes_per_step": int(48 * H * W) * world, "ms_per_step": float(t_e.item()), "api": "create_sr_model(...)[0].upsample(ndarray) (pageable host arrays)"}
    if rank == 0:
        pk = peaks()
        conv_t = sum(conv_ms) / len(conv_ms) / 1e3
        ach = EDSR_FLOP_PER_LR_PX * H * W / conv_t / 1e12
        cpu = None
        if world == 1 and not args.no_cpu:
            v, desc, cores = cpu_sample(wl)
            cpu = {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": desc}
        line = {"metric": "output Mpix/s (EDSR-baseline x4)", "value": out_mpix / (ms / 1e3), "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
                "data": "synthetic",
                "config": {"workload": wl["label"], "weights": "seed-0 PyTorch default init, resblock outputs x0.1 (random-init; EDSR_x4.pb is not available offline)",
                           "parallelism": "replicas only: one tile per rank, no exchange", "l2": "HR activations (2 GB per tensor) exceed the L2; no explicit flush"},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "roofline": {"bound": "tensor", "kernel": "all 43 tensor-core conv launches of one EDSR forward + the head conv", "achieved": ach,
                             "peak": pk["bf16_tflops"], "peak_sustained": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                             "peak_source": pk["src"] + " (burst: a forward lasts milliseconds)", "traffic": None,
                             "conv_ms_per_step": conv_t * 1e3, "algorithmic_flops_per_step": EDSR_FLOP_PER_LR_PX * H * W},
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="scene")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiler passes only; such a line is not a bench value)")
    ap.add_argument("--opt", action="append", default=[], help="libwowsr option key=value")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg3":
        run_edsr(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
