"""First timing of the experimental dataflow trunk on the deployment window size (276 x 276, 23 blocks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import wowsr_b200 as ws
from oracle import rrdbnet_ref as R
blocks = 23
sd = R.calibrate_conv_last(R.random_init_state_dict(0, blocks), blocks)
tensors = [sd[k + s].numpy() for k, _, _ in R.conv_specs(blocks) for s in (".weight", ".bias")]
img = np.random.default_rng(1).integers(0, 256, (1044, 1044, 3), dtype=np.uint8)   # 5 x 5 windows of 276 x 276
outs = {}
for name, opts in (("layer-by-layer", {}), ("dataflow G=2", {"trunk_dataflow": 1}), ("dataflow G=4", {"trunk_dataflow": 1, "trunk_group": 4})):
    h = ws.Handle(0)
    for k, v in opts.items():
        h.set_option(k, v)
    h.load_rrdbnet(tensors, blocks, precision="bf16")
    for rep in range(3):
        try:
            outs[name] = h.enhance_host(img, 256)
            print(name, "rep", rep, h.timing(), flush=True)
        except Exception as e:  # noqa: BLE001
            print(name, "FAILED", e, flush=True)
            break
    h.close()
a = outs.get("layer-by-layer")
for k, v in outs.items():
    if a is not None and k != "layer-by-layer":
        print(k, "u8 within1 vs layer-by-layer", float((np.abs(a.astype(int) - v.astype(int)) <= 1).mean()), flush=True)
