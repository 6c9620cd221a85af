"""Fixture generator (build container; needs cv2 with the dnn module and torch): pins the ARITHMETIC ENGINE of the EDSR path.

The reference runs EDSR through ``cv2.dnn_superres.DnnSuperResImpl`` (server/app/super_resolution.py:92-124,196), a thin
wrapper around ``cv2.dnn`` that feeds the float BGR image to the network and converts the output with a rounding saturate
cast.  ``cv2.dnn_superres`` (contrib) and ``EDSR_x4.pb`` are absent here, but ``cv2.dnn`` itself is present: this script
exports the layer list of oracle/edsr_ref.py (seeded weights) to ONNX, runs it through ``cv2.dnn`` and stores input, float
output and uint8 output.  What this pins: our restatement and the CUDA path agree with OpenCV's dnn engine on the EDSR-baseline
graph as we restate it.  What stays UNPINNED: that the restated graph equals the TensorFlow graph inside EDSR_x4.pb.

torch's TorchScript ONNX exporter post-processes the serialized model with the ``onnx`` package (not installed) only to
attach onnxscript functions, of which this graph has none; that step is bypassed.

    python tests/golden/make_golden_edsr.py        -> tests/golden/edsr_cv2dnn_64x80.npz
"""
import os
import sys
import tempfile

import cv2
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import edsr_ref as E  # noqa: E402


class Net(nn.Module):
    def __init__(self, sd, nb=16):
        super().__init__()
        self.nb = nb
        for name, cin, cout in E.conv_specs(nb):
            c = nn.Conv2d(cin, cout, 3, 1, 1)
            c.weight.data.copy_(sd[name + ".weight"])
            c.bias.data.copy_(sd[name + ".bias"])
            setattr(self, name.replace(".", "_"), c)
        self.register_buffer("mean", torch.tensor(E.MEAN_BGR, dtype=torch.float32).view(1, 3, 1, 1))

    def forward(self, x):
        x = x - self.mean
        h = self.head(x)
        r = h
        for b in range(self.nb):
            r = r + getattr(self, f"body_{b}_conv2")(F.relu(getattr(self, f"body_{b}_conv1")(r)))
        r = self.body_end(r) + h
        u = F.pixel_shuffle(self.up1(r), 2)
        u = F.pixel_shuffle(self.up2(u), 2)
        return self.tail(u) + self.mean


def main():
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    onnx_proto_utils._add_onnxscript_fn = lambda proto, custom_opsets: proto
    seed, shape = 0, (64, 80)
    sd = E.random_init_state_dict(seed, 16)
    net = Net(sd).eval()
    path = os.path.join(tempfile.mkdtemp(), "edsr.onnx")
    torch.onnx.export(net, torch.zeros(1, 3, *shape), path, opset_version=13, input_names=["x"], output_names=["y"], dynamo=False)
    engine = cv2.dnn.readNetFromONNX(path)
    img = np.random.default_rng(12).integers(0, 256, shape + (3,), dtype=np.uint8)
    engine.setInput(np.ascontiguousarray(img.astype(np.float32).transpose(2, 0, 1)[None]))
    y = engine.forward()[0].transpose(1, 2, 0).copy()
    # dnn_superres converts its float result with Mat::convertTo(CV_8U), a rounding (half-to-even) saturate cast
    u8 = cv2.convertScaleAbs(np.maximum(y, 0.0))
    assert np.array_equal(u8, np.clip(np.rint(y), 0, 255).astype(np.uint8))
    ref = E.forward_float(sd, img, 16)
    print("cv2", cv2.__version__, "dnn vs torch restatement: max abs diff", float(np.abs(y - ref).max()), "of", float(np.abs(ref).max()))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "edsr_cv2dnn_64x80.npz")
    np.savez_compressed(out, seed=seed, img=img, f32=y.astype(np.float32), u8=u8, cv2_version=cv2.__version__)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
