"""CPU: design check for the next version of the device corner selection.

cv::goodFeaturesToTrack keeps a local maximum when no STRONGER kept corner lies closer than the minimum distance (greedy,
strongest first, oracle/detector.c).  k_det_select does that with one warp over sorted candidates; the planned kernel
decides all candidates in parallel rounds instead:  a candidate is KEPT as soon as every stronger candidate within the
distance is decided and none of them is kept, REJECTED as soon as one stronger neighbour is kept.  This test shows on real
min-eigenvalue maps that the rounds reach the same set as the greedy scan (and how few rounds it takes), so the kernel can
be written against a known-equivalent formulation."""
import numpy as np
import pytest

import rvio_b200  # noqa: F401
from rvio_b200 import synth
from oracle import oracle as orc


def _candidates(eig, quality):
    thr = np.float32(float(eig.max()) * quality)
    h, w = eig.shape
    c = eig[1:-1, 1:-1]
    is_max = c > thr
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            is_max &= ~(eig[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx] > c)
    ys, xs = np.nonzero(is_max)
    ys += 1; xs += 1
    order = np.lexsort((-(ys * w + xs), -eig[ys, xs]))            # value desc, index desc (greaterThanPtr)
    return xs[order], ys[order]


def _parallel_rounds(xs, ys, d):
    n = len(xs)
    cell = int(round(d))
    cx, cy = xs // cell, ys // cell
    buckets = {}
    for i in range(n):
        buckets.setdefault((cx[i], cy[i]), []).append(i)
    stronger = []                                                 # stronger candidates within the distance, per candidate
    d2 = d * d
    for i in range(n):
        nb = []
        for yy in (cy[i] - 1, cy[i], cy[i] + 1):
            for xx in (cx[i] - 1, cx[i], cx[i] + 1):
                for j in buckets.get((xx, yy), ()):
                    if j < i and (xs[i] - xs[j]) ** 2 + (ys[i] - ys[j]) ** 2 < d2:
                        nb.append(j)
        stronger.append(np.array(nb, np.int64))
    state = np.zeros(n, np.int8)                                  # 0 undecided, 1 kept, 2 rejected
    rounds = 0
    while (state == 0).any():
        rounds += 1
        snap = state.copy()                                       # synchronous rounds: decisions use the previous round's states
        for i in np.nonzero(snap == 0)[0]:
            s = snap[stronger[i]]
            if (s == 1).any():
                state[i] = 2
            elif not (s == 0).any():
                state[i] = 1
    return np.nonzero(state == 1)[0], rounds


@pytest.mark.parametrize("s", [1, 2])
def test_parallel_rounds_equal_greedy(s):
    import cv2
    cfg = synth.baseline_config(1)
    cfg.width, cfg.height = 376, 240                              # quarter-size frame keeps the pure-Python part short
    st = synth.Stream(cfg, 2, 20261003, t_static=0.05)
    img = orc.clahe(st.frames[1])
    q = float(np.float32(cfg.qual_lvl)); d = s * float(np.float32(cfg.min_dist))
    xs, ys = _candidates(orc.min_eig_map(img), q)
    kept, rounds = _parallel_rounds(xs, ys, d)
    got = np.stack([xs[kept], ys[kept]], 1).astype(np.float32)   # rank order == strongest first
    want = orc.good_features(img, 0, q, d)                        # greedy scan, no corner limit
    assert len(xs) > 2000 and len(want) > 30
    assert np.array_equal(got, want), (len(got), len(want))
    # truncating the rank-ordered result is what the corner limit does in the greedy scan
    lim = orc.good_features(img, 20, q, d)
    assert np.array_equal(got[:20], lim)
    print(f"s={s}: {len(xs)} local maxima -> {len(want)} corners in {rounds} rounds")
    assert rounds <= 24
