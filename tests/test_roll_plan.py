"""Work list of the rolling conv kernel (csrc/roll_kernel.cuh, wowsr_debug_roll_plan): every (column, row) of a launch is
covered exactly once, pair partners walk segments of the same length, and the units are balanced."""
import numpy as np
import pytest

import wowsr_b200 as ws


def _cover(n_win, h, w, strip_x0, pair, units):
    tasks, off, info = ws._lib.roll_plan(n_win, h, w, strip_x0, pair, units)
    assert len(off) == info["units"] + 1 and off[0] == 0 and off[-1] == len(tasks) and np.all(np.diff(off) >= 0)
    assert 1 <= info["units"] <= units
    seen_h = np.zeros((n_win, (strip_x0 + 127) // 128, h), np.int32)
    rem = w - strip_x0
    seen_v = np.zeros((n_win, (h + 127) // 128 if rem else 0, max(rem, 1)), np.int32)
    load = np.zeros(info["units"], np.int64)
    for u in range(info["units"]):
        vert = u >= info["units_h"]
        for t in tasks[off[u]:off[u + 1]]:
            n0, n1, u0, u1, v0, rows = [int(x) for x in t[:6]]
            assert rows >= 1
            load[u] += rows + 4
            for n, uu in ((n0, u0), (n1, u1)) if pair else ((n0, u0),):
                if n < 0:
                    continue
                assert uu % 128 == 0
                if vert:
                    assert strip_x0 <= v0 and v0 + rows <= w
                    seen_v[n, uu // 128, v0 - strip_x0:v0 - strip_x0 + rows] += 1
                else:
                    assert 0 <= v0 and v0 + rows <= h
                    seen_h[n, uu // 128, v0:v0 + rows] += 1
            if not pair:
                assert n0 >= 0
    assert np.all(seen_h == 1)
    if rem:
        assert np.all(seen_v == 1)
    return tasks, off, info, load


@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("shape", [(64, 532, 532, 512), (231, 276, 276, 256), (1, 128, 128, 128), (1, 40, 48, 48), (3, 150, 276, 256),
                                   (9, 532, 532, 512), (1, 300, 290, 256), (5, 1104, 1104, 1024), (2, 70, 33, 33), (1, 2128, 2128, 2048)])
def test_exact_cover_and_balance(shape, pair):
    n_win, h, w, sx = shape
    units = 74 if pair else 148
    tasks, off, info, load = _cover(n_win, h, w, sx, pair, units)
    total_rows = n_win * ((sx + 127) // 128) * h
    if total_rows >= 64 * units:  # enough work: no unit carries more than its share + one column's fixed cost + a minimum segment
        assert load.max() <= load.mean() * 1.10 + 16, (load.max(), load.mean())


def test_small_unit_counts():
    for units in (2, 3, 7):
        _cover(4, 276, 276, 256, True, units)
        _cover(4, 276, 276, 256, False, units)
    _cover(4, 276, 276, 276, True, 1)   # one unit cannot serve both orientations: the caller drops the strip
    _cover(4, 276, 276, 276, False, 1)
    with pytest.raises(ValueError):
        ws._lib.roll_plan(4, 276, 276, 256, True, 1)
