"""Cycle attribution of CTA 0 of selected rolling-kernel launches (csrc/roll_kernel.cuh, P.trace) of one forward.
usage: python tools/roll_trace.py [cfg2s|cfg5s|cfg1] [key=value ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import wowsr_b200 as ws  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2s"
opts = dict(kv.split("=") for kv in sys.argv[2:])
torch.manual_seed(0)
sd = ws.app.cnn_super_resolution.RRDBNet(3, 3, 64, 23, 32, 4).state_dict()
h = ws.Handle(0)
for k, v in opts.items():
    h.set_option(k, int(v))
tile, side = {"cfg2s": (512, 1200), "cfg1": (256, 128)}.get(which, (256, 1280))
up = ws.app.cnn_super_resolution.RealESRGAN(device="cuda", tile_size=tile, state_dict=sd, handle=h)
img = torch.from_numpy(bench.make_lr_image(side, side)).cuda()
up.enhance_cuda(img)
torch.cuda.synchronize()
print(which, opts)
for layer, name in [(11, "rdb.conv1"), (12, "rdb.conv2"), (14, "rdb.conv4"), (10, "rdb2.conv5 (lo in, lo out)"), (15, "rdb3.conv5 (lo in, fp32 res2, fp32 out)"),
                    (20, "rdb1.conv5 (fp32 res1, lo out)"), (346, "conv_body"), (347, "conv_up1"), (348, "conv_up2"), (349, "conv_hr"), (350, "conv_last")]:
    h.set_option("tc_trace_layer", layer)
    up.enhance_cuda(img)
    torch.cuda.synchronize()
    t = h.debug_trace().reshape(-1)
    h.set_option("tc_trace_layer", 0)
    prod, issue, epi = t[0:4], t[4:8], t[8:15]
    if issue[3] == 0:
        print(f"== {name} (launch {layer}): no rolling trace (tile kernel?)")
        continue
    g = float(issue[3])
    print(f"== {name} (launch {layer}): {int(g)} groups (2 input rows each), {int(prod[3])} stages")
    print(f"   issuer   : {issue[0] / g:8.0f} cyc/group total; waiting for TMA data {issue[1] / g:7.0f}, waiting for a free accumulator pair {issue[2] / g:7.0f}")
    print(f"   producer : {prod[0] / g:8.0f} cyc/group total; waiting for a free stage {prod[1] / g:7.0f}  ({int(prod[2])} boxes)")
    print(f"   epilogue : {epi[0] / g:8.0f} cyc/group total; waiting for accumulators {epi[1] / g:7.0f}, TMEM read + clear + hand-back {epi[2] / g:7.0f}, "
          f"arithmetic + stores after the hand-back {epi[3] / g:7.0f}  ({int(epi[4])} pairs)")
    if len(t) > 20 and t[16]:
        ph = [int(t[i] - t[16]) for i in range(17, 21)]
        print(f"   phases of CTA 0 (ns after entry): barriers + TMEM ready {ph[0]}, weights resident {ph[1]}, roles done {ph[2]}, exit {ph[3]}")
    if epi[5] or epi[6]:
        print(f"              plain N = 32 epilogue of the traced warp: bias / activation / pack {epi[5] / g:7.0f}, transpose + stores {epi[6] / g:7.0f} cyc/pair")
