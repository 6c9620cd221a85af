"""Generates the committed golden fixtures by running the UNMODIFIED reference
(/root/reference/server/app, imported via oracle/refload.py) in the build container.

    python tests/golden/make_golden.py

Outputs (np.savez_compressed, all small):
  post_wow_96x128.npz, post_farm_101x77.npz  : input RGB + reference _enhance_for_crops / farm trio output
  rrdb2_tiled_50x70.npz                      : 2-block RRDBNet, tile_size=16 (4x5 windows), input + u8 + float out
  rrdb23_cfg1_64.npz                         : 23-block x4plus, 64x64 crop of BASELINE config 1 input, u8 + float
  rrdb23_cfg1_128_u8.npz                     : BASELINE config 1 (128x128, seed 0), reference uint8 output
  init_checksums.npz                         : float64 sums of the seed-0 default-init weights (RNG stream pin)
  green_mask_u8_90x121.npz, green_mask_u16_64x80.npz : compute_green_mask_hsv (vector_extraction.py:222-270) of the unmodified
                                               reference, its raster served by a stub rasterio.open (`python make_golden.py green`)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402
from tests.conftest import image_like  # noqa: E402


def main():
    cnn, wow, farm = refload.load()
    torch.set_num_threads(os.cpu_count())

    img = image_like(96, 128, seed=11)
    np.savez_compressed(os.path.join(HERE, "post_wow_96x128.npz"), img=img, out=wow._enhance_for_crops(img))
    img = image_like(101, 77, seed=12)
    out = farm.enhance_vegetation(farm.apply_unsharp_mask(farm.enhance_local_contrast(img, clip_limit=2.5, grid_size=8),
                                                          strength=1.2, radius=1.5))
    np.savez_compressed(os.path.join(HERE, "post_farm_101x77.npz"), img=img, out=out)

    def run(blocks, img, tile, seed=0, want_float=True):
        torch.manual_seed(seed)
        model = cnn.RRDBNet(3, 3, 64, blocks, 32, 4).eval()
        up = refload.make_upsampler(cnn, model, tile_size=tile)
        u8 = up.enhance(img)
        f = None
        if want_float:
            x = torch.from_numpy(img.astype(np.float32) / 255.0).permute(2, 0, 1).unsqueeze(0)
            with torch.no_grad():
                h, w = x.shape[2:]
                y = up._tile_process(x) if h * w > tile * tile * 4 else model(x)
            f = y.squeeze(0).permute(1, 2, 0).numpy()
        return u8, f, model

    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (50, 70, 3), dtype=np.uint8)
    u8, f, _ = run(2, img, 16)
    np.savez_compressed(os.path.join(HERE, "rrdb2_tiled_50x70.npz"), img=img, u8=u8, f32=f, blocks=2, tile=16, seed=0)

    rng = np.random.default_rng(0)
    img128 = rng.integers(0, 256, (128, 128, 3), dtype=np.uint8)
    u8, f, model = run(23, img128[:64, :64].copy(), 256)
    np.savez_compressed(os.path.join(HERE, "rrdb23_cfg1_64.npz"), img=img128[:64, :64], u8=u8, f32=f, blocks=23, tile=256, seed=0)
    u8, _, _ = run(23, img128, 256, want_float=False)
    np.savez_compressed(os.path.join(HERE, "rrdb23_cfg1_128_u8.npz"), img=img128, u8=u8, blocks=23, tile=256, seed=0)
    sd = model.state_dict()
    names = ["conv_first.weight", "body.0.rdb1.conv1.weight", "body.11.rdb2.conv5.bias", "body.22.rdb3.conv5.weight", "conv_last.weight"]
    np.savez_compressed(os.path.join(HERE, "init_checksums.npz"), names=np.array(names),
                        sums=np.array([float(sd[n].double().sum()) for n in names]),
                        abssums=np.array([float(sd[n].double().abs().sum()) for n in names]))
    print("golden fixtures written")


def main_green():
    """Fixtures of the HSV vegetation mask: uniform colour noise (every hue / saturation / value box edge is hit) with an
    image-like half, default and non-default ExtractionConfig; a uint16 raster for the max-scaling branch (:245-247)."""
    store = {}
    ve = refload.load_vector_extraction(lambda p: store[str(p)])
    rng = np.random.default_rng(31)
    img = rng.integers(0, 256, (90, 121, 3), dtype=np.uint8)
    img[45:] = (img[45:].astype(np.int32) * 3 // 8 + image_like(45, 121, seed=32).astype(np.int32) * 5 // 8).astype(np.uint8)
    store["u8"] = [img[..., c] for c in range(3)]
    cfg2 = dict(hsv_green_hue_range=(30, 90), hsv_saturation_min=20, hsv_value_min=50)
    np.savez_compressed(os.path.join(HERE, "green_mask_u8_90x121.npz"), img=img,
                        mask_default=ve.compute_green_mask_hsv("u8", ve.ExtractionConfig()),
                        mask_cfg2=ve.compute_green_mask_hsv("u8", ve.ExtractionConfig(**cfg2)),
                        cfg2_hue=np.array(cfg2["hsv_green_hue_range"]), cfg2_sat=cfg2["hsv_saturation_min"], cfg2_val=cfg2["hsv_value_min"])
    raster = rng.integers(0, 9000, (64, 80, 3)).astype(np.uint16)
    store["u16"] = [raster[..., c] for c in range(3)]
    np.savez_compressed(os.path.join(HERE, "green_mask_u16_64x80.npz"), raster=raster,
                        mask_default=ve.compute_green_mask_hsv("u16", ve.ExtractionConfig()))
    print("green-mask fixtures written")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "green":
        main_green()
    else:
        main()
        main_green()
