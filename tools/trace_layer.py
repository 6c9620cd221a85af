"""Per-tile timeline (clock64) of CTA 0 for selected conv launches of one cfg2s forward — ROUND-1 TILE KERNEL (option roll=0);
the rolling kernel's per-role cycle attribution is tools/roll_trace.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wowsr_b200 as ws
from oracle import rrdbnet_ref as R
sys.path.insert(0, ROOT)
import bench

sd = R.random_init_state_dict(0, 23)
up = ws.app.cnn_super_resolution.RealESRGAN(device="cuda", tile_size=512, state_dict=sd)
img = torch.from_numpy(bench.make_lr_image(1200, 1200)).cuda()
up._h.set_option("roll", 0)
up.enhance_cuda(img); torch.cuda.synchronize()
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 1
up._h.set_option("tc_flags", flags)
print("tc_flags", flags)
for layer, name in [(11, "rdb.conv1"), (14, "rdb.conv4"), (15, "rdb.conv5"), (349, "conv_hr")]:
    up._h.set_option("tc_trace_layer", layer)
    up.enhance_cuda(img); torch.cuda.synchronize()
    t_all = up._h.debug_trace()
    up._h.set_option("tc_trace_layer", 0)
    eb = t_all[64:128]
    t = t_all[:64]
    t = t[(t != 0).all(1)]
    eb = eb[eb[:, 2] > 0]
    if len(eb):
        it = eb[:, 2].astype(float)
        print(f"   epilogue breakdown (warp 0, per 32-channel iteration): tmem_ld+wait {float((eb[:,0]/it).mean()):.0f} cyc, "
              f"math+loads+stores {float((eb[:,1]/it).mean()):.0f} cyc, iterations/tile {it.mean():.1f}, wait for accumulators {float(eb[:,3].mean()):.0f} cyc/tile")
    t0 = t[0, 0]
    print(f"== {name} (launch {layer}): tile  mma_start  mma_issue_len  epi_start-mma_issued  epi_len  | next_mma_start - mma_start")
    for i in range(min(len(t), 4)):
        ms, mi, es, ee = t[i]
        nxt = t[i + 1, 0] - ms if i + 1 < len(t) else 0
        print(f"   {i:3d} {ms - t0:9d} {mi - ms:9d} {es - mi:9d} {ee - es:9d} | {nxt:9d}")
    print("   mean tile period", float(np.diff(t[:, 0]).mean()), "mean mma issue len", float((t[:, 1] - t[:, 0]).mean()), "mean epi len", float((t[:, 3] - t[:, 2]).mean()))
