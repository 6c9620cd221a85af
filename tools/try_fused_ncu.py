"""One forward of a 2-block RRDBNet on 25 windows of 276x276 for the ncu DRAM-traffic pass of tools/experiments/r2_queue.sh:
argument 0 = layer-by-layer, f > 0 = fused tail from conv f (option trunk_fuse)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import wowsr_b200 as ws  # noqa: E402
from oracle import rrdbnet_ref as R  # noqa: E402

fuse = int(sys.argv[1]) if len(sys.argv) > 1 else 0
blocks = 2
sd = R.random_init_state_dict(0, blocks)
tensors = [sd[k + s].numpy() for k, _, _ in R.conv_specs(blocks) for s in (".weight", ".bias")]
img = np.random.default_rng(1).integers(0, 256, (1044, 1044, 3), dtype=np.uint8)
h = ws.Handle(0)
if fuse:
    h.set_option("trunk_fuse", fuse)
h.load_rrdbnet(tensors, blocks, precision="bf16")
h.enhance_host(img, 256)
print(h.timing())
