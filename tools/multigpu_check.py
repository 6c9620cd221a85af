"""Run under torchrun on N GPUs: the sharded scene (scene.run_scene over NCCL) must equal the single-GPU result
bit for bit (same kernels, same windows; only the decomposition differs)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import wowsr_b200 as ws  # noqa: E402
from oracle import rrdbnet_ref as R  # noqa: E402

scene = importlib.import_module("sentinel2-super-resolution-poc_b200.scene")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    blocks = 2
    sd = R.random_init_state_dict(0, blocks)
    sd = R.calibrate_conv_last(sd, blocks)
    up = ws.app.cnn_super_resolution.RealESRGAN(device=f"cuda:{local}", tile_size=256, state_dict=sd, model_name="realesrgan_anime") \
        if False else None
    cnn = ws.app.cnn_super_resolution
    cnn.MODELS["test2"] = dict(cnn.MODELS["realesrgan_anime"], blocks=blocks)
    up = cnn.RealESRGAN(device=f"cuda:{local}", tile_size=256, state_dict=sd, model_name="test2")
    ok_all = True
    for (H, W, kind) in [(600, 700, "wow"), (1100, 530, "farm"), (300, 280, "wow")]:
        img = np.random.default_rng(H).integers(0, 256, (H, W, 3), dtype=np.uint8)
        d = torch.from_numpy(img).to(dev)
        backend = scene.GpuBackend(up, ws._lib.post_params(kind))
        plan, band, full = scene.run_scene(backend, d, 256, post=True, gather=True)
        torch.cuda.synchronize()
        if rank == 0:
            sr = up.enhance_cuda(d)
            fn = ws.app.wow_sr.enhance_for_crops_cuda if kind == "wow" else ws.app.farm_sr.farm_post_cuda
            want = fn(sr.contiguous())
            same = bool(torch.equal(full, want))
            ok_all &= same
            print(f"[multigpu_check] {H}x{W} {kind} world={world} bands={plan.bands} pieces={len(plan.pieces)} "
                  f"equal_to_single_gpu={same}", flush=True)
        dist.barrier()
        # host-to-host variant: every rank copies its band into one shared page-locked host image
        shared = scene.SharedHostImage(4 * H, 4 * W)
        if rank == 0:   # a failing registration (same range twice) must not poison later CUDA calls
            shared.pin_rows(0, 8)
            rc = torch.cuda.cudart().cudaHostRegister(shared._range[0], shared._range[1] - shared._range[0], 0)
            assert int(rc) != 0
            scene._clear_cuda_error()
            torch.zeros(4, device=dev).sum().item()
        scene.run_scene_to_host(backend, torch.from_numpy(img).pin_memory(), 256, shared)
        if rank == 0:
            same = bool(torch.equal(shared.array, want.cpu()))
            ok_all &= same
            print(f"[multigpu_check] {H}x{W} {kind} shared host image (pinned={shared.pinned}) equal_to_single_gpu={same}", flush=True)
        shared.close()
    if rank == 0:
        print("[multigpu_check] " + ("PASS" if ok_all else "FAIL"), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
