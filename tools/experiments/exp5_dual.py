"""Experiment: two half-batches on two streams, each conv launch limited to half of the SMs, so that DRAM-bound
layers (rdb.conv5) of one stream overlap tensor-bound layers (rdb.conv1-4) of the other."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wowsr_b200 as ws
from oracle import rrdbnet_ref as R
import bench

sd = R.random_init_state_dict(0, 23)
img = torch.from_numpy(bench.make_lr_image(4096, 4096)).cuda()
halves = [img[:2048].contiguous(), img[2048:].contiguous()]

def make(grid):
    up = ws.app.cnn_super_resolution.RealESRGAN(device="cuda", tile_size=512, state_dict=sd)
    if grid:
        up._h.set_option("tc_grid", grid)
    return up

def run_single(up):
    torch.cuda.synchronize(); t0 = time.time()
    for h in halves:
        up.enhance_cuda(h)
    torch.cuda.synchronize(); return time.time() - t0

def run_dual(ups):
    streams = [torch.cuda.Stream() for _ in ups]
    def work(i):
        with torch.cuda.stream(streams[i]):
            ups[i].enhance_cuda(halves[i])
    torch.cuda.synchronize(); t0 = time.time()
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize(); return time.time() - t0

single = make(0)
for rep in range(3):
    print("single handle, 148 CTAs, 2 x 32 windows sequential: %.1f ms" % (1e3 * run_single(single)), flush=True)
for g in (74, 80, 100, 148):
    ups = [make(g), make(g)]
    for rep in range(3):
        print("dual handles, tc_grid=%d each, 32 windows each concurrent: %.1f ms" % (g, 1e3 * run_dual(ups)), flush=True)
    del ups
