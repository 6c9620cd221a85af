"""Bring-up probe for the tensor-core conv kernel (run on the GPU box):
each variant runs in its own process so a faulting variant cannot poison the others."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(cfg):
    import numpy as np
    import torch
    import torch.nn.functional as F
    import wowsr_b200 as ws
    h = ws.Handle(0)
    for k, v in cfg.get("opts", {}).items():
        h.set_option(k, v)
    rng = np.random.default_rng(cfg.get("seed", 0))
    n, hh, ww, cin, cout = cfg["n"], cfg["h"], cfg["w"], cfg["cin"], cfg["cout"]
    x = rng.standard_normal((n, hh, ww, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
    mask = np.zeros((3, 3), np.float32)
    for (ky, kx) in cfg["taps"]:
        mask[ky, kx] = 1
    w *= mask
    b = rng.standard_normal(cout).astype(np.float32) * 0.1
    prec = cfg.get("prec", "bf16")
    dt = torch.bfloat16 if prec == "bf16" else torch.float16
    xr = torch.from_numpy(x).to(dt).float().permute(0, 3, 1, 2)
    wr = torch.from_numpy(w).to(dt).float()
    ref = F.conv2d(xr.double(), wr.double(), torch.from_numpy(b).double(), padding=1)
    if cfg.get("act"):
        ref = F.leaky_relu(ref, 0.2)
    ref = ref.permute(0, 2, 3, 1).numpy()
    out = h.conv3x3_host(x, w, b, act=cfg.get("act", 0), precision=prec)
    err = np.abs(out - ref)
    bad = np.argwhere(err > 1e-3)
    res = {"max_err": float(err.max()), "mean_err": float(err.mean()), "ref_absmax": float(np.abs(ref).max()),
           "n_bad": int(len(bad)), "first_bad": bad[:6].tolist()}
    print("RESULT " + json.dumps(res))


def main():
    all_taps = [(ky, kx) for ky in range(3) for kx in range(3)]
    base = dict(n=2, h=11, w=150, cin=64, cout=32)
    variants = []
    # simple (CUDA-core) kernel first: validates the harness itself
    variants.append(("simple N32", dict(base, taps=all_taps, opts={"conv_impl": 1})))
    for flags, fname in [(0, "unstacked/bo0"), (2, "unstacked/bo1"), (1, "stacked/bo0"), (3, "stacked/bo1")]:
        for taps, tname in [([(1, 0)], "tap(1,0) aligned"), ([(1, 1)], "tap(1,1) +128B"), ([(1, 2)], "tap(1,2) +256B"),
                            ([(0, 0)], "tap(0,0)"), ([(2, 0)], "tap(2,0)"), (all_taps, "all taps")]:
            variants.append((f"tc {fname} {tname}", dict(base, taps=taps, opts={"tc_flags": flags})))
    for flags in (1, 3, 0, 2):
        variants.append((f"tc flags{flags} cin96 N32", dict(base, cin=96, taps=all_taps, opts={"tc_flags": flags})))
        variants.append((f"tc flags{flags} cin192 N64 act", dict(base, cin=192, cout=64, act=1, taps=all_taps, opts={"tc_flags": flags})))
        variants.append((f"tc flags{flags} cin64 N16(cout3)", dict(base, cin=64, cout=3, taps=all_taps, opts={"tc_flags": flags})))
        variants.append((f"tc flags{flags} cin160 N32 fp16 big", dict(base, n=3, h=37, w=276, cin=160, prec="fp16", taps=all_taps,
                                                                   opts={"tc_flags": flags})))
        variants.append((f"tc flags{flags} stream-weights cin128", dict(base, cin=128, taps=all_taps,
                                                                        opts={"tc_flags": flags, "tc_force_stream": 1})))
    results = []
    for name, cfg in variants:
        try:
            p = subprocess.run([sys.executable, __file__, "--one", json.dumps(cfg)], capture_output=True, text=True, timeout=180)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                r = json.loads(line[-1][7:])
                status = "OK " if r["max_err"] < 1e-3 else "BAD"
                msg = f"{status} {name}: max_err={r['max_err']:.3e} mean={r['mean_err']:.3e} refmax={r['ref_absmax']:.2f} n_bad={r['n_bad']} first_bad={r['first_bad'][:3]}"
            else:
                msg = f"ERR {name}: rc={p.returncode} {p.stderr.strip().splitlines()[-1:] if p.stderr.strip() else ''}"
        except subprocess.TimeoutExpired:
            msg = f"TIMEOUT {name}"
        print(msg, flush=True)
        results.append(msg)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.txt"), "w") as f:
        f.write("\n".join(results) + "\n")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        one(json.loads(sys.argv[2]))
    else:
        main()
