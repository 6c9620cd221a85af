"""Cycle attribution of CTA 0 of selected rolling-kernel launches (csrc/roll_kernel.cuh, P.trace) of one forward.
usage: python tools/roll_trace.py [cfg2s|cfg5s|cfg5b|cfg1] [key=value ...]   (cfg5b: one full batch of 144 windows of 276^2)"""
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
tile, side = {"cfg2s": (512, 1200), "cfg1": (256, 128), "cfg5b": (256, 3072)}.get(which, (256, 1280))
up = ws.app.cnn_super_resolution.RealESRGAN(device="cuda", tile_size=tile, state_dict=sd, handle=h)
img = torch.from_numpy(bench.make_lr_image(side, side)).cuda()
for _ in range(5):          # with roll_adapt=1 the work lists adapt to the measured unit speeds over the first four batches (conv.cu, balance_update)
    up.enhance_cuda(img)
torch.cuda.synchronize()
print(which, opts)
_win = {"cfg2s": (9, 532, 512), "cfg1": (1, 128, 128), "cfg5b": (144, 276, 256)}.get(which, (25, 276, 256))
_info = ws._lib.roll_plan(_win[0], _win[1], _win[1], _win[2], True, 74)[2]
print(f"LR-resolution work list ({_win[0]} windows of {_win[1]}^2): {_info['units']} units, the first {_info['units_h']} horizontal, the others vertical (remainder strip)")
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
    if len(t) > 64 + 296 and t[64]:
        st, en = t[64:64 + 148].astype(np.float64), t[64 + 148:64 + 296].astype(np.float64)
        live = en > 0
        if live.any():
            t_first = st[live].min()
            dur = (en[live] - t_first) / 1e3
            # pair mode: CTAs 2k, 2k+1 form unit k; units_h is not exported here, so print the sorted finish times per unit
            per_unit = dur[::2] if live.sum() > 74 else dur
            q = np.percentile(per_unit, [0, 25, 50, 75, 100])
            print(f"   finish time of the units (us after the first role start): min {q[0]:.1f}  q25 {q[1]:.1f}  median {q[2]:.1f}  q75 {q[3]:.1f}  max {q[4]:.1f}"
                  f"  -> mean / max = {per_unit.mean() / q[4]:.3f}")
            print("   per unit: " + " ".join(f"{v:.0f}" for v in per_unit))
            if len(t) >= 64 + 444:
                sm = t[64 + 296:64 + 444][live]
                sm_u = sm[::2] if live.sum() > 74 else sm
                print("   SM of the unit's first CTA: " + " ".join(str(int(v)) for v in sm_u))
                order = np.argsort(sm_u)
                print("   finish time by SM id:       " + " ".join(f"{per_unit[i]:.0f}" for i in order))
    if epi[5] or epi[6]:
        print(f"              plain N = 32 epilogue of the traced warp: bias / activation / pack {epi[5] / g:7.0f}, transpose + stores {epi[6] / g:7.0f} cyc/pair")
