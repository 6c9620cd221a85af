"""CPU: pins the torch-fp32 RRDBNet oracle (oracle/rrdbnet_ref.py) against the committed goldens produced
by the unmodified reference, and against the reference classes themselves when the tree is present."""
import os

import numpy as np
import pytest
import torch

from oracle import refload
from oracle import rrdbnet_ref as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_init_rng_stream_matches_reference_checksums():
    g = np.load(os.path.join(GOLD, "init_checksums.npz"))
    sd = R.random_init_state_dict(0, 23)
    assert len(sd) == 702
    for n, s, a in zip(g["names"], g["sums"], g["abssums"]):
        t = sd[str(n)].double()
        assert float(t.sum()) == pytest.approx(float(s), abs=1e-9)
        assert float(t.abs().sum()) == pytest.approx(float(a), abs=1e-9)
    assert sum(v.numel() for v in sd.values()) == 16697987


def test_tiled_golden_2_blocks():
    g = np.load(os.path.join(GOLD, "rrdb2_tiled_50x70.npz"))
    sd = R.random_init_state_dict(int(g["seed"]), int(g["blocks"]))
    f = R.enhance_float(sd, g["img"], int(g["blocks"]), int(g["tile"]))
    assert np.abs(f - g["f32"]).max() < 1e-5
    assert np.array_equal(R.quantise(f), g["u8"]) or (np.abs(R.quantise(f).astype(int) - g["u8"]).max() <= 1)


def test_cfg1_crop_golden_23_blocks():
    g = np.load(os.path.join(GOLD, "rrdb23_cfg1_64.npz"))
    sd = R.random_init_state_dict(0, 23)
    torch.set_num_threads(os.cpu_count())
    f = R.enhance_float(sd, g["img"], 23, 256)
    assert np.abs(f - g["f32"]).max() < 1e-4
    d = np.abs(R.quantise(f).astype(int) - g["u8"].astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.9999


def test_window_table_values():
    # SURVEY Appendix B window tables
    t = R.plan_axis(10980, 256)
    assert len(t) == 43 and t[0] == (0, 276, 0, 266) and t[1] == (256, 532, 266, 522) and t[-1] == (10704, 10980, 10714, 10980)
    t = R.plan_axis(4096, 512)
    assert len(t) == 8 and t[-1] == (3564, 4096, 3574, 4096)
    assert len(R.plan_axis(10980, 512)) == 22


def test_tiling_with_identity_model():
    """The fake-model trick: an identity-x4 model makes the stitched output np.repeat(img)."""
    fake = lambda t: torch.nn.functional.interpolate(t, scale_factor=4, mode="nearest")
    for (h, w, T) in [(513, 512, 256), (300, 2000, 256), (70, 50, 16), (257, 257, 64)]:
        x = torch.rand(1, 3, h, w)
        y = R.tile_process(fake, x, T)
        assert torch.equal(y, fake(x))


@pytest.mark.skipif(not refload.available(), reason="reference tree not present")
def test_against_unmodified_reference_classes():
    cnn, _, _ = refload.load()
    torch.manual_seed(0)
    m = cnn.RRDBNet(3, 3, 64, 2, 32, 4).eval()
    sd = R.random_init_state_dict(0, 2)
    assert all(torch.equal(sd[k], v) for k, v in m.state_dict().items())
    x = torch.rand(1, 3, 40, 36)
    with torch.no_grad():
        assert float((m(x) - R.rrdbnet_forward(sd, x, 2)).abs().max()) == 0.0
    up = refload.make_upsampler(cnn, m, tile_size=16)
    img = np.random.default_rng(1).integers(0, 256, (50, 70, 3), dtype=np.uint8)
    assert np.array_equal(up.enhance(img), R.enhance(sd, img, 2, 16))


def test_split_trunk_representation_bound():
    """the residual trunk is stored as hi + lo with hi = bf16(x) and lo = bf16(x - hi).  Host-side
    restatement of that arithmetic: x is reproduced to 2^-17 relative (two 8-bit significands plus the sign of lo), x - hi is exact in fp32,
    and 5 * hi (the identity K-step's product, 1/0.2) is exact in fp32."""
    import torch
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1 << 16, generator=g) * torch.logspace(-3, 3, 1 << 16)
    hi = x.to(torch.bfloat16).float()
    d = x - hi
    assert torch.equal(d.double(), x.double() - hi.double())                 # Sterbenz-style exactness of the difference
    lo = d.to(torch.bfloat16).float()
    err = (x.double() - (hi.double() + lo.double())).abs()
    assert float((err / x.double().abs()).max()) <= 2.0 ** -17
    assert torch.equal((5.0 * hi).double(), 5.0 * hi.double())               # 8-bit significand x 3-bit constant
    # (acc + 5*hi) * 0.2f reproduces hi to fp32 rounding: 5 * fl(0.2) = 1 + 1.5e-8
    back = (5.0 * hi) * torch.tensor(0.2, dtype=torch.float32)
    assert float(((back - hi).abs() / hi.abs().clamp_min(1e-30)).max()) <= 2.0 ** -23


def test_product_parameter_container_draws_the_reference_init_stream(ws):
    """bench.py draws its seed-0 weights through the package's own RRDBNet container (nothing under oracle/ in the measured
    arm): same construction order as the reference class, hence the same RNG stream — pinned by the reference's checksums."""
    g = np.load(os.path.join(GOLD, "init_checksums.npz"))
    torch.manual_seed(0)
    sd = ws.app.cnn_super_resolution.RRDBNet(3, 3, 64, 23, 32, 4).state_dict()
    assert len(sd) == 702 and list(sd) == list(R.random_init_state_dict(0, 23))
    for n, s, a in zip(g["names"], g["sums"], g["abssums"]):
        t = sd[str(n)].double()
        assert float(t.sum()) == pytest.approx(float(s), abs=1e-9)
        assert float(t.abs().sum()) == pytest.approx(float(a), abs=1e-9)
