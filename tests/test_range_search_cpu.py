"""Host restatement of the warp-wide 32-ary lower bound of csrc/binning.cu::tile_ranges_views_kernel (32 probes per round:
`pos = min(lo + (lane + 1) * step - 1, hi - 1)`, first probe with key >= t bounds the answer from above, the probe before it
from below) against numpy's searchsorted, and of the ranges it produces against the boundary-detection definition of the
reference's identifyTileRanges: empty tiles (0, 0), otherwise [first, last + 1)."""
import numpy as np


def warp_lower_bound(keys, lo, hi, t):
    rounds = 0
    while hi > lo:
        step = (hi - lo + 31) >> 5
        pos = [min(lo + (lane + 1) * step - 1, hi - 1) for lane in range(32)]
        ge = [keys[p] >= t for p in pos]
        rounds += 1
        if not any(ge):
            return hi, rounds
        f = ge.index(True)
        pf = pos[f]
        pprev = lo - 1 if f == 0 else min(lo + f * step - 1, hi - 1)
        lo, hi = pprev + 1, pf
    return lo, rounds


def ranges_by_search(keys, G):
    N = len(keys)
    out = np.zeros((G, 2), np.uint32)
    for t in range(G):
        first, _ = warp_lower_bound(keys, 0, N, t)
        last, _ = warp_lower_bound(keys, first, N, t + 1)
        if last > first:
            out[t] = (first, last)
    return out


def ranges_by_boundaries(keys, G):
    out = np.zeros((G, 2), np.uint32)
    N = len(keys)
    for i in range(N):
        if i == 0 or keys[i] != keys[i - 1]:
            out[keys[i], 0] = i
            if i:
                out[keys[i - 1], 1] = i
    if N:
        out[keys[-1], 1] = N
    return out


def test_lower_bound_equals_searchsorted():
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 31, 32, 33, 1000, 1025, 40000):
        for G in (1, 7, 300):
            keys = np.sort(rng.integers(0, G, n)).astype(np.uint32)
            for t in list(range(0, G + 1, max(1, G // 13))) + [G, G + 5]:
                got, rounds = warp_lower_bound(keys, 0, n, t)
                assert got == int(np.searchsorted(keys, t, side="left")), (n, G, t)
                assert rounds <= 4      # 32^3 > 40000: three rounds narrow to one element, a fourth confirms it


def test_ranges_equal_the_boundary_definition():
    rng = np.random.default_rng(2)
    for n, G in ((0, 5), (1, 5), (500, 6), (5000, 97), (3000, 2000)):
        keys = np.sort(rng.integers(0, G, n)).astype(np.uint32)
        np.testing.assert_array_equal(ranges_by_search(keys, G), ranges_by_boundaries(keys, G))
    keys = np.full(777, 3, np.uint32)                      # one tile holds everything
    np.testing.assert_array_equal(ranges_by_search(keys, 6), ranges_by_boundaries(keys, 6))
