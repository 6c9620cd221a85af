"""End-to-end RRDBNet check on the GPU box: libwowsr (simple / tensor-core kernels) vs the CPU fp32 oracle."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import wowsr_b200 as ws  # noqa: E402
from oracle import rrdbnet_ref as R  # noqa: E402


def report(tag, got_u8, got_f, ref_f):
    ref_u8 = R.quantise(ref_f)
    d = np.abs(got_u8.astype(int) - ref_u8.astype(int))
    mse = float(((got_u8.astype(np.float64) - ref_u8) ** 2).mean())
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    print(f"{tag}: within1={100*(d<=1).mean():.4f}% exact={100*(d==0).mean():.3f}% max={d.max()} psnr={psnr:.1f}dB "
          f"float_maxerr={np.abs(got_f-ref_f).max():.3e} (ref absmax {np.abs(ref_f).max():.2f}) nonzero={100*(ref_u8>0).mean():.1f}%", flush=True)


def main():
    blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    tile = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (size, size + 8, 3), dtype=np.uint8)
    sd0 = R.random_init_state_dict(0, blocks)
    for wname, sd in (("default-init", sd0), ("calibrated", R.calibrate_conv_last(sd0, blocks))):
        t0 = time.time()
        ref_f = R.enhance_float(sd, img, blocks, tile)
        print(f"[{wname}] oracle {time.time()-t0:.1f}s", flush=True)
        for prec in ("bf16", "fp16"):
            for impl in (1, 0):
                try:
                    h = ws.Handle(0)
                    h.set_option("conv_impl", impl)
                    if len(sys.argv) > 4:
                        h.set_option("tc_flags", int(sys.argv[4]))
                    tensors = [sd[k + s].numpy() for k, _, _ in R.conv_specs(blocks) for s in (".weight", ".bias")]
                    h.load_rrdbnet(tensors, blocks, precision=prec)
                    t0 = time.time()
                    u8, f = h.enhance_host(img, tile, want_float=True)
                    dt = time.time() - t0
                    report(f"[{wname}] {prec} {'simple' if impl else 'tc'} ({dt*1e3:.0f} ms, {h.timing()})", u8, f, ref_f)
                except Exception as e:  # noqa: BLE001
                    print(f"[{wname}] {prec} impl={impl}: FAILED {e}", flush=True)


if __name__ == "__main__":
    main()
