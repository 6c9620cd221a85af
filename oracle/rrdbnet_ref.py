"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference's RRDBNet path.

Imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs; the product package never touches it.

Restates, with plain torch fp32 functional ops on the CPU (the reference itself *is* torch fp32 on
the CPU — ``Dockerfile:34`` installs CPU-only torch):

* ``ResidualDenseBlock.forward``  server/app/cnn_super_resolution.py:85-91
* ``RRDB.forward``                server/app/cnn_super_resolution.py:103-107
* ``RRDBNet.__init__/forward``    server/app/cnn_super_resolution.py:122-158
* ``RealESRGAN.enhance``          server/app/cnn_super_resolution.py:217-234
* ``RealESRGAN._tile_process``    server/app/cnn_super_resolution.py:236-280

Pinned against the unmodified reference classes in ``tests/test_oracle_rrdbnet.py`` (runs when
/root/reference is present) and through the committed fixtures under ``tests/golden/``.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

TILE_PAD = 10  # cnn_super_resolution.py:172


def conv_specs(num_block=23, num_feat=64, num_grow=32, num_in=3, num_out=3):
    """(state-dict prefix, Cin, Cout) in the reference's construction order (:122-136, :78-82, :99-101)."""
    specs = [("conv_first", num_in, num_feat)]
    for b in range(num_block):
        for r in (1, 2, 3):
            for k in range(1, 6):
                cin = num_feat + (k - 1) * num_grow
                cout = num_grow if k < 5 else num_feat
                specs.append((f"body.{b}.rdb{r}.conv{k}", cin, cout))
    specs += [("conv_body", num_feat, num_feat), ("conv_up1", num_feat, num_feat),
              ("conv_up2", num_feat, num_feat), ("conv_hr", num_feat, num_feat),
              ("conv_last", num_feat, num_out)]
    return specs


def random_init_state_dict(seed=0, num_block=23, num_feat=64, num_grow=32):
    """PyTorch default Conv2d init drawn in the reference's construction order under manual_seed(seed).

    Equals ``torch.manual_seed(seed); RRDBNet(3,3,num_feat,num_block,num_grow,4).state_dict()`` of
    the reference class bit for bit (checked in tests/test_oracle_rrdbnet.py).
    """
    torch.manual_seed(seed)
    sd = OrderedDict()
    for name, cin, cout in conv_specs(num_block, num_feat, num_grow):
        conv = torch.nn.Conv2d(cin, cout, 3, 1, 1)
        sd[name + ".weight"] = conv.weight.detach().clone()
        sd[name + ".bias"] = conv.bias.detach().clone()
    return sd


def calibrate_conv_last(sd, num_block, probe_seed=1234, size=48):
    """Second weight set of SURVEY.md 8(d): affinely rescale conv_last so fp32 outputs span ~[0,1]."""
    rng = np.random.default_rng(probe_seed)
    x = torch.from_numpy(rng.random((1, 3, size, size), dtype=np.float32))
    with torch.no_grad():
        y = rrdbnet_forward(sd, x, num_block)
    m = float(y.mean())
    s = float(y.std())
    g = 0.2 / max(s, 1e-12)
    sd = OrderedDict((k, v.clone()) for k, v in sd.items())
    sd["conv_last.weight"] = sd["conv_last.weight"] * g
    sd["conv_last.bias"] = (sd["conv_last.bias"] - m) * g + 0.5
    return sd


def _conv(sd, name, x):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=1)


def _lrelu(x):
    return F.leaky_relu(x, 0.2)


def rdb_forward(sd, prefix, x):
    x1 = _lrelu(_conv(sd, prefix + ".conv1", x))
    x2 = _lrelu(_conv(sd, prefix + ".conv2", torch.cat([x, x1], 1)))
    x3 = _lrelu(_conv(sd, prefix + ".conv3", torch.cat([x, x1, x2], 1)))
    x4 = _lrelu(_conv(sd, prefix + ".conv4", torch.cat([x, x1, x2, x3], 1)))
    x5 = _conv(sd, prefix + ".conv5", torch.cat([x, x1, x2, x3, x4], 1))
    return x5 * 0.2 + x


def rrdb_forward(sd, prefix, x):
    out = rdb_forward(sd, prefix + ".rdb1", x)
    out = rdb_forward(sd, prefix + ".rdb2", out)
    out = rdb_forward(sd, prefix + ".rdb3", out)
    return out * 0.2 + x


@torch.no_grad()
def rrdbnet_forward(sd, x, num_block=23, scale=4):
    feat = _conv(sd, "conv_first", x)
    body = feat
    for b in range(num_block):
        body = rrdb_forward(sd, f"body.{b}", body)
    feat = feat + _conv(sd, "conv_body", body)
    feat = _lrelu(_conv(sd, "conv_up1", F.interpolate(feat, scale_factor=2, mode="nearest")))
    if scale == 4:
        feat = _lrelu(_conv(sd, "conv_up2", F.interpolate(feat, scale_factor=2, mode="nearest")))
    feat = _lrelu(_conv(sd, "conv_hr", feat))
    return _conv(sd, "conv_last", feat)


def plan_axis(L: int, T: int, P: int = TILE_PAD):
    """Window table along one axis: list of (a, b, keep_lo, keep_hi) in LR pixels (:244-277)."""
    n = (L + T - 1) // T
    out = []
    for i in range(n):
        a = i * T
        b = min(a + T + 2 * P, L)
        a = max(b - T - 2 * P, 0)
        lo = a + (P if i > 0 else 0)
        hi = b - (P if i < n - 1 else 0)
        out.append((a, b, lo, hi))
    return out


@torch.no_grad()
def tile_process(model_fn, img: torch.Tensor, tile_size: int, scale: int = 4) -> torch.Tensor:
    """RealESRGAN._tile_process: row-major windows, later writes win (:236-280)."""
    _, c, H, W = img.shape
    out = torch.zeros((1, c, H * scale, W * scale))
    for (y1, y2, ylo, yhi) in plan_axis(H, tile_size):
        for (x1, x2, xlo, xhi) in plan_axis(W, tile_size):
            t = model_fn(img[:, :, y1:y2, x1:x2])
            t = t[:, :, (ylo - y1) * scale:(yhi - y1) * scale, (xlo - x1) * scale:(xhi - x1) * scale]
            out[:, :, ylo * scale:yhi * scale, xlo * scale:xhi * scale] = t
    return out


@torch.no_grad()
def enhance_float(sd, img: np.ndarray, num_block=23, tile_size=256, scale=4) -> np.ndarray:
    """enhance() up to, but excluding, the uint8 quantisation: float32 HxWx3 (:220-231)."""
    x = torch.from_numpy(img.astype(np.float32) / 255.0).permute(2, 0, 1).unsqueeze(0)
    h, w = x.shape[2:]
    fn = lambda t: rrdbnet_forward(sd, t, num_block, scale)
    y = tile_process(fn, x, tile_size, scale) if h * w > tile_size * tile_size * 4 else fn(x)
    return y.squeeze(0).permute(1, 2, 0).numpy()


def quantise(out_f: np.ndarray) -> np.ndarray:
    """(out*255).clip(0,255).astype(uint8) — truncation (:232)."""
    return (out_f * 255.0).clip(0, 255).astype(np.uint8)


def enhance(sd, img: np.ndarray, num_block=23, tile_size=256, scale=4) -> np.ndarray:
    return quantise(enhance_float(sd, img, num_block, tile_size, scale))
