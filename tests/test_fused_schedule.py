"""CPU: the host-built task list of the EXPERIMENTAL fused-tail launch (csrc/sched_plan.h, option trunk_fuse — round-2 work
in progress, not on the product path) through the C ABI: tile cover, dependency cover, band targets, producers before
consumers, no deadlock for CTAs that walk the list in order, and the skew that hides the publish -> poll latency."""
import numpy as np
import pytest

K, VERT, WIN, U0, V0, DEP0, DEPN, PUB0, PUBN = range(9)


def _decode(arr):
    t = np.empty((len(arr), 9), dtype=np.int64)
    t[:, K] = arr[:, 0] & 255
    t[:, VERT] = arr[:, 0] >> 8
    t[:, WIN:] = arr[:, 1:]
    return t


def _out_rect(t, h, w):
    """(y0, y1, x0, x1) of the output pixels a tile owns (exclusive ends), clipped to the window."""
    R = 4 if t[K] == 4 else 8
    if t[VERT]:
        return t[U0], min(t[U0] + 128, h), t[V0], min(t[V0] + R, w)
    return t[V0], min(t[V0] + R, h), t[U0], min(t[U0] + 128, w)


SHAPES = [(276, 276, 8, 4, 0), (276, 276, 2, 1, 0), (276, 276, 3, 4, 0), (276, 276, 3, 4, 120), (276, 276, 2, 3, 0), (276, 276, 2, 2, 0), (532, 532, 2, 4, 120), (532, 532, 3, 4, 0),
          (148, 148, 9, 4, 0), (40, 48, 1, 4, 0), (300, 290, 1, 4, 0), (150, 276, 2, 4, 0), (64, 128, 2, 4, 0), (9, 1000, 1, 2, 0)]


@pytest.mark.parametrize("h,w,n_win,first,lag", SHAPES)
def test_schedule_is_complete_and_safe(ws, h, w, n_win, first, lag):
    _check_schedule(ws, h, w, n_win, first, lag, check_slack=True)


def test_schedule_random_shapes(ws):
    """Property test over random window shapes (ragged widths with and without a remainder strip, partial last bands)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(h=st.integers(8, 300), w=st.integers(1, 600), n_win=st.integers(1, 3), first=st.integers(1, 4), lag=st.sampled_from([0, -1]))
    def run(h, w, n_win, first, lag):
        _check_schedule(ws, h, w, n_win, first, lag)

    run()


def _check_schedule(ws, h, w, n_win, first, lag, check_slack=False):
    arr, info = ws._lib.fused_schedule(h, w, n_win, first, lag)
    t = _decode(arr)
    nb, br, x0 = info["n_bands"], info["band_rows"], info["strip_x0"]
    assert nb == (h + br - 1) // br and info["n_layers"] == 6 - first
    # (1) per conv and window the tiles partition the window: every output pixel is owned exactly once
    for k in range(first - 1, 5):
        for win in range(n_win):
            cover = np.zeros((h, w), dtype=np.int32)
            for row in t[(t[:, K] == k) & (t[:, WIN] == win)]:
                y0, y1, xa, xb = _out_rect(row, h, w)
                assert y1 > y0 and xb > xa
                assert (xa >= x0) == bool(row[VERT])              # vertical tiles exactly on the remainder strip
                cover[y0:y1, xa:xb] += 1
            assert (cover == 1).all(), (k, win)
    # (2) bands: publications per band == the target the kernel waits for; only convs before conv5 publish; the first fused
    #     conv waits for nothing (its inputs come from earlier launches)
    pubs = np.zeros((5, n_win, nb), dtype=np.int64)
    for row in t:
        assert (row[PUBN] == 0) == (row[K] == 4)
        assert (row[DEPN] == 0) == (row[K] == first - 1)
        y0, y1, _, _ = _out_rect(row, h, w)
        if row[PUBN]:
            assert (row[PUB0], row[PUB0] + row[PUBN] - 1) == (y0 // br, (y1 - 1) // br)     # exactly the bands it covers
            pubs[row[K], row[WIN], row[PUB0]:row[PUB0] + row[PUBN]] += 1
        if row[DEPN]:
            lo, hi = max(y0 - 1, 0), min(y1, h - 1)                                          # halo rows of a 3x3 conv
            assert (row[DEP0], row[DEP0] + row[DEPN] - 1) == (lo // br, hi // br)
    for k in range(first - 1, 4):
        assert (pubs[k] == info["band_target_tiles"]).all()
    # (3) producers precede consumers: replay the list, a dependency must be complete when its consumer is reached
    done = np.zeros((5, n_win, nb), dtype=np.int64)
    last_pub_index = np.full((5, n_win, nb), -1, dtype=np.int64)
    slack = []
    for i, row in enumerate(t):
        if row[DEPN]:
            sl = slice(row[DEP0], row[DEP0] + row[DEPN])
            assert (done[row[K] - 1, row[WIN], sl] == info["band_target_tiles"]).all(), i
            slack.append(i - last_pub_index[row[K] - 1, row[WIN], sl].max())
        if row[PUBN]:
            sl = slice(row[PUB0], row[PUB0] + row[PUBN])
            done[row[K], row[WIN], sl] += 1
            last_pub_index[row[K], row[WIN], sl] = i
    assert min(slack) >= 1
    if check_slack and ((lag >= 120 and w <= 276) or lag == 0) and len(t) > 1500:
        # the point of the skew: a consumer's last producer is more than one full machine of tasks (148 CTAs) behind it —
        # everywhere but in the fill phase at the head of the list, where only the first conv has tiles to interleave
        slack = np.array(slack)
        assert np.median(slack) >= 240 and (slack <= 148).mean() < 0.02, (np.median(slack), (slack <= 148).mean())


@pytest.mark.parametrize("grid", [1, 3, 7, 148])
def test_in_order_ctas_never_deadlock(ws, grid):
    """CTA b walks tasks b, b + grid, ... in order and blocks on unfinished dependencies; with random task durations the
    whole list must still drain (event simulation with the kernel's static assignment)."""
    arr, info = ws._lib.fused_schedule(276, 276, 2, 3, 0)
    t = _decode(arr)
    n, nb = len(t), info["n_bands"]
    rng = np.random.default_rng(grid)
    dur = rng.integers(1, 50, n)
    done = np.zeros((5, 2, nb), dtype=np.int64)
    pos = list(range(min(grid, n)))              # next task of every CTA
    busy_until = np.zeros(len(pos), dtype=np.int64)
    running = [None] * len(pos)
    clock, finished = 0, 0
    while finished < n:
        progressed = False
        for c in range(len(pos)):
            if running[c] is not None and busy_until[c] <= clock:
                row = t[running[c]]
                if row[PUBN]:
                    done[row[K], row[WIN], row[PUB0]:row[PUB0] + row[PUBN]] += 1
                running[c] = None
                finished += 1
                progressed = True
            if running[c] is None and pos[c] < n:
                row = t[pos[c]]
                if not row[DEPN] or (done[row[K] - 1, row[WIN], row[DEP0]:row[DEP0] + row[DEPN]] == info["band_target_tiles"]).all():
                    running[c] = pos[c]
                    busy_until[c] = clock + dur[pos[c]]
                    pos[c] += grid
                    progressed = True
        if not progressed:
            nxt = [busy_until[c] for c in range(len(pos)) if running[c] is not None]
            assert nxt, f"deadlock at clock {clock}: {finished} of {n} tasks finished"
            clock = min(nxt)


def test_unschedulable_shapes_and_lags_are_refused(ws):
    with pytest.raises(ValueError):
        ws._lib.fused_schedule(4, 64, 1, 4, 0)        # no 8-row band structure
    with pytest.raises(ValueError):
        ws._lib.fused_schedule(276, 276, 25, 4, 8)    # explicit lag shorter than the reach of a strip tile's halo
    for first in (0, 5):                              # conv5 alone is not a fusion; there is no conv0
        with pytest.raises(ValueError):
            ws._lib.fused_schedule(276, 276, 1, first, 0)
    legal_min = ws._lib.fused_schedule(276, 276, 25, 4, -1)[1]["lag"]
    auto = ws._lib.fused_schedule(276, 276, 25, 4, 0)[1]["lag"]
    assert legal_min % 8 == 0 and legal_min >= 48
    assert auto == legal_min + 80                      # 240 tasks of slack at 3 tasks per step (conv4 tile + two conv5 tiles)
