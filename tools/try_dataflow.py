"""One-shot check of the experimental dataflow trunk (csrc/trunk_kernel.cuh) against the layer-by-layer path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import wowsr_b200 as ws
from oracle import rrdbnet_ref as R

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (300, 290)
sd = R.calibrate_conv_last(R.random_init_state_dict(4, blocks), blocks)
tensors = [sd[k + s].numpy() for k, _, _ in R.conv_specs(blocks) for s in (".weight", ".bias")]
img = np.random.default_rng(13).integers(0, 256, (H, W, 3), dtype=np.uint8)


def run(tile, **opts):
    h = ws.Handle(0)
    for k, v in opts.items():
        h.set_option(k, v)
    h.load_rrdbnet(tensors, blocks, precision="bf16")
    t0 = time.time()
    out = h.enhance_host(img, tile, want_float=True)
    print(tile, opts, "ms", round((time.time() - t0) * 1e3, 1), h.timing(), flush=True)
    h.close()
    return out


for tile, dbg, reps in ((256, 0, 3), (128, 0, 3), (128, 1, 1), (256, 1, 1)):
    u8, f = run(tile)
    for rep in range(reps):
        try:
            u8_d, f_d = run(tile, trunk_dataflow=1, trunk_debug=dbg)
            d = np.abs(f - f_d)
            print("  dataflow vs layer-by-layer: float max diff", float(d.max()), "u8 within1",
                  float((np.abs(u8.astype(int) - u8_d.astype(int)) <= 1).mean()), "exact", float((u8 == u8_d).mean()), flush=True)
        except Exception as e:  # noqa: BLE001
            print("  dataflow FAILED:", e, flush=True)
