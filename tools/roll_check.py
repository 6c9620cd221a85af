"""GPU bring-up of the rolling conv kernel (csrc/roll_kernel.cuh): one process per mode so a fault in one mode cannot hide the
others.    python tools/roll_check.py layer|net|perf <mode>      mode: tile | roll1 (single CTA) | roll2 (CTA pairs)"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wowsr_b200 as ws  # noqa: E402
from oracle import rrdbnet_ref as R  # noqa: E402

MODES = {"tile": dict(roll=0), "roll1": dict(roll=1, roll_pair=0), "roll2": dict(roll=1, roll_pair=1)}


def handle(mode, **extra):
    h = ws.Handle(0)
    for k, v in {**MODES[mode], **extra}.items():
        h.set_option(k, v)
    return h


def layer(mode):
    h = handle(mode)
    bad = 0
    for (cin, cout, act, prec) in [(64, 32, 1, "bf16"), (96, 32, 1, "bf16"), (128, 32, 0, "fp16"), (160, 32, 1, "bf16"), (192, 64, 0, "bf16"),
                                   (64, 64, 1, "fp16"), (64, 3, 0, "bf16")]:
        for (n, hh, w) in [(1, 8, 128), (2, 11, 150), (1, 40, 276), (2, 300, 148), (1, 130, 20), (3, 276, 276), (1, 532, 532)]:
            rng = np.random.default_rng(cin * 7 + cout)
            x = rng.standard_normal((n, hh, w, cin)).astype(np.float32)
            wt = (rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
            b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
            dt = torch.bfloat16 if prec == "bf16" else torch.float16
            ref = F.conv2d(torch.from_numpy(x).to(dt).double().permute(0, 3, 1, 2), torch.from_numpy(wt).to(dt).double(),
                           torch.from_numpy(b).double(), padding=1)
            if act:
                ref = F.leaky_relu(ref, 0.2)
            ref = ref.permute(0, 2, 3, 1).numpy()
            try:
                out = h.conv3x3_host(x, wt, b, act=act, precision=prec)
                err = float(np.abs(out - ref).max())
                tol = 2e-5 * max(1.0, float(np.abs(ref).max()))
                ok = err < tol
                where = ""
                if not ok:
                    d = np.abs(out - ref)
                    idx = np.unravel_index(np.argmax(d), d.shape)
                    rows_bad = np.unique(np.argwhere(d > tol)[:, 1])
                    where = f" worst at {idx}; bad rows {rows_bad[:12].tolist()}... ({len(rows_bad)} rows, {int((d > tol).sum())} values)"
                print(f"  {mode} conv {cin}->{cout} act {act} {prec} shape {(n, hh, w)}: max err {err:.3e} (tol {tol:.1e}) {'ok' if ok else 'MISMATCH' + where}",
                      flush=True)
                bad += not ok
            except Exception as e:  # noqa: BLE001
                print(f"  {mode} conv {cin}->{cout} shape {(n, hh, w)}: EXCEPTION {e}", flush=True)
                bad += 1
    print(f"{mode} layer check: {bad} failures")
    return bad


def net(mode):
    blocks = 2
    sd = R.calibrate_conv_last(R.random_init_state_dict(4, blocks), blocks)
    tensors = [sd[k + s].numpy() for k, _, _ in R.conv_specs(blocks) for s in (".weight", ".bias")]
    bad = 0
    for shape, tile in (((40, 48), 256), ((150, 276), 256), ((560, 290), 256), ((300, 290), 128)):
        img = np.random.default_rng(13).integers(0, 256, shape + (3,), dtype=np.uint8)
        ref_f = R.enhance_float(sd, img, blocks, tile)
        h = handle(mode)
        h.load_rrdbnet(tensors, blocks, precision="bf16")
        try:
            u8, f = h.enhance_host(img, tile, want_float=True)
            ref = R.quantise(ref_f)
            w1 = float((np.abs(u8.astype(int) - ref.astype(int)) <= 1).mean())
            err = float(np.abs(f - ref_f).max())
            ok = w1 >= 0.999 and err < 0.02 * max(1.0, float(np.abs(ref_f).max()))
            print(f"  {mode} net {shape} tile {tile}: within1 {w1:.6f} float max err {err:.3e} {'ok' if ok else 'MISMATCH'}", flush=True)
            bad += not ok
        except Exception as e:  # noqa: BLE001
            print(f"  {mode} net {shape} tile {tile}: EXCEPTION {e}", flush=True)
            bad += 1
        h.close()
    print(f"{mode} net check: {bad} failures")
    return bad


def perf(mode, nwin=25, size=276):
    blocks = 23
    sd = R.random_init_state_dict(0, blocks)
    tensors = [sd[k + s].numpy() for k, _, _ in R.conv_specs(blocks) for s in (".weight", ".bias")]
    side = int(round(nwin ** 0.5))
    tile = size - 20
    img = np.random.default_rng(1).integers(0, 256, (tile * side, tile * side, 3), dtype=np.uint8)
    h = handle(mode)
    h.load_rrdbnet(tensors, blocks, precision="bf16")
    d = torch.from_numpy(img).cuda()
    out = torch.empty((img.shape[0] * 4, img.shape[1] * 4, 3), dtype=torch.uint8, device="cuda")
    for rep in range(4):
        t0 = time.perf_counter()
        h.enhance_dev(d.data_ptr(), img.shape[0], img.shape[1], tile, out.data_ptr())
        torch.cuda.synchronize()
        print(f"  {mode} perf {side * side} windows of {size}: rep {rep} wall {1e3 * (time.perf_counter() - t0):.1f} ms {h.timing()}", flush=True)
    return 0


if __name__ == "__main__":
    what, mode = sys.argv[1], sys.argv[2]
    extra = [int(a) for a in sys.argv[3:]]
    sys.exit({"layer": layer, "net": net, "perf": perf}[what](mode, *extra))
