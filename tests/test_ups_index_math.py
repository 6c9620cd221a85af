"""CPU: index algebra of the EXPERIMENTAL folded-upsample kernel (csrc/ups_kernel.cuh), restated in numpy.

The kernel never sees an upsampled tensor: per stage row it loads the box (rep = 0..1, source pixels xs0 .. xs0 + 65) of source row
``row_up >> 1`` with ``xs0 = (u0 >> 1) - 1`` (out-of-bounds elements zero filled), which lands in shared memory in pixel order
``2 * (xs - xs0) + rep`` (each row its own box on a 1024-byte boundary); the MMA of output pixel ``m`` and run-axis tap ``kx`` reads stage pixel ``1 + m + kx`` (descriptor start
+128 B, tap shift kx * 128 B).  This test replays exactly that addressing for every tile of small images and compares with
``conv3x3(nearest_x2(src))`` with zero padding — the reference's ``conv(F.interpolate(x, scale_factor=2, mode="nearest"))``
(cnn_super_resolution.py:150-153)."""
import numpy as np
import pytest


def _stage_row(src, xs0, src_row):
    """The 132 'pixels' one TMA box delivers: zero outside the source image (rows and columns)."""
    H, W = src.shape
    out = np.zeros(132)
    if 0 <= src_row < H:
        for p in range(132):
            xs = xs0 + p // 2
            if 0 <= xs < W:
                out[p] = src[src_row, xs]
    return out


@pytest.mark.parametrize("H,W", [(5, 7), (9, 70), (3, 129), (66, 64)])
def test_folded_upsample_addressing_equals_conv_of_nearest_upsample(H, W):
    rng = np.random.default_rng(H * 131 + W)
    src = rng.standard_normal((H, W))
    w = rng.standard_normal((3, 3))
    up = np.repeat(np.repeat(src, 2, 0), 2, 1)
    padded = np.pad(up, 1)
    want = sum(w[ky, kx] * padded[ky:ky + 2 * H, kx:kx + 2 * W] for ky in range(3) for kx in range(3))
    got = np.zeros((2 * H, 2 * W))
    R = 4
    for v0 in range(0, 2 * H, R):                 # tile rows of the UPSAMPLED image
        for u0 in range(0, 2 * W, 128):           # 128-pixel runs
            xs0 = (u0 >> 1) - 1
            rows = {}
            for sp in range((R + 2) // 2):        # the producer's loop: two rows per stage
                for half in range(2):
                    row_up = v0 - 1 + 2 * sp + half
                    rows[row_up] = _stage_row(src, xs0, row_up >> 1)      # arithmetic shift: (-1) >> 1 == -1 -> zero row
            for r in range(R):
                y = v0 + r
                if y >= 2 * H:
                    break
                for m in range(128):
                    x = u0 + m
                    if x >= 2 * W:
                        break
                    got[y, x] = sum(w[ky, kx] * rows[y + ky - 1][1 + m + kx] for ky in range(3) for kx in range(3))
    assert np.allclose(got, want, atol=1e-12)
