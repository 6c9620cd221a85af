"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of EDSR-baseline x4 ("farm SR" variant).

PARITY UNPINNED.  The reference reaches EDSR only through ``cv2.dnn_superres.DnnSuperResImpl`` with an external
TensorFlow graph (``EDSR_x4.pb`` from github.com/Saafke/EDSR_Tensorflow, master, unpinned URL —
server/app/super_resolution.py:31-34,92-124,196).  Neither the contrib module (``hasattr(cv2,'dnn_superres')`` is
False in this image) nor the weights are available offline and the reference has no test for it, so this file
restates the published EDSR-baseline definition (Lim et al. 2017, as exported by that repository) and is checked
only for self-consistency: the CUDA path must match THIS restatement on seeded random weights.

Layer list (B = 16 resblocks, F = 64 features, no batch norm, res_scale = 1):
    x = BGR uint8 (0..255) - mean_BGR                        mean = (103.1545782, 111.561547, 114.35629928)
    h = head(x)                       conv3x3 3->F
    r = h; for b in range(B): r = r + res_scale * conv2(relu(conv1(r)))
    r = body_end(r) + h               conv3x3 F->F, global skip
    u = depth_to_space(up1(r), 2)     conv3x3 F->4F, PixelShuffle(2)
    u = depth_to_space(up2(u), 2)
    y = tail(u) + mean_BGR            conv3x3 F->3
    out = uint8(clip(rint(y), 0, 255))                        (cv2 convertTo/saturate_cast rounding)
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

MEAN_BGR = (103.1545782, 111.561547, 114.35629928)


def conv_specs(num_block=16, nf=64):
    specs = [("head", 3, nf)]
    for b in range(num_block):
        specs += [(f"body.{b}.conv1", nf, nf), (f"body.{b}.conv2", nf, nf)]
    specs += [("body_end", nf, nf), ("up1", nf, 4 * nf), ("up2", nf, 4 * nf), ("tail", nf, 3)]
    return specs


def random_init_state_dict(seed=0, num_block=16, nf=64, body_gain=0.1):
    """Seeded PyTorch default init; resblock outputs damped by `body_gain` so 16 blocks stay well conditioned."""
    torch.manual_seed(seed)
    sd = OrderedDict()
    for name, cin, cout in conv_specs(num_block, nf):
        conv = torch.nn.Conv2d(cin, cout, 3, 1, 1)
        w, b = conv.weight.detach().clone(), conv.bias.detach().clone()
        if name.endswith("conv2"):
            w, b = w * body_gain, b * body_gain
        sd[name + ".weight"], sd[name + ".bias"] = w, b
    return sd


@torch.no_grad()
def forward_float(sd, img: np.ndarray, num_block=16, res_scale=1.0) -> np.ndarray:
    mean = torch.tensor(MEAN_BGR, dtype=torch.float32).view(1, 3, 1, 1)
    x = torch.from_numpy(img.astype(np.float32)).permute(2, 0, 1).unsqueeze(0) - mean
    c = lambda n, t: F.conv2d(t, sd[n + ".weight"], sd[n + ".bias"], padding=1)
    h = c("head", x)
    r = h
    for b in range(num_block):
        r = r + res_scale * c(f"body.{b}.conv2", F.relu(c(f"body.{b}.conv1", r)))
    r = c("body_end", r) + h
    u = F.pixel_shuffle(c("up1", r), 2)
    u = F.pixel_shuffle(c("up2", u), 2)
    y = c("tail", u) + mean
    return y.squeeze(0).permute(1, 2, 0).numpy()


def quantise(y: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(y), 0, 255).astype(np.uint8)


def upsample(sd, img, num_block=16, res_scale=1.0):
    return quantise(forward_float(sd, img, num_block, res_scale))
