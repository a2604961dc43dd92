"""Host-side restatement of K1's `tile_span` (csrc/preprocess.cu): the tile-index range [t0, t1) an interval [lo, hi] of pixel
coordinates reaches must be EXACTLY the set of tiles for which the blend kernels' staging test -- `hi >= tile0 && lo <= tile0 + 15`
with tile0 = 16 t, the union of the eight sub-tile tests of blend_common.cuh::stage_entry -- holds, in fp32, for every input
including interval ends on tile boundaries, negative coordinates, sub-pixel intervals and values the image never sees.
That identity is what makes tight binning drop precisely the instances whose eight sub-tile bits are clear."""
import numpy as np

f32 = np.float32


def tile_span(lo, hi):
    """the kernel's arithmetic, op for op, in float32"""
    lo, hi = f32(lo), f32(hi)
    u = f32(lo * f32(0.0625))
    fu = np.floor(u)
    t0 = int(fu) + (1 if f32(u - fu) > f32(0.9375) else 0)
    t1 = min(int(np.floor(f32(hi * f32(0.0625)))), 1 << 24) + 1
    return t0, t1


def reached(lo, hi, t):
    lo, hi = f32(lo), f32(hi)
    tile0 = f32(16 * t)
    return bool(hi >= tile0 and lo <= f32(tile0 + f32(15.0)))


def _check(lo, hi, t_range):
    t0, t1 = tile_span(lo, hi)
    for t in t_range:
        assert reached(lo, hi, t) == (t0 <= t < t1), (lo, hi, t, t0, t1)


def test_tile_span_edges():
    T = range(-4, 12)
    for lo, hi in [(0.0, 0.0), (15.0, 15.0), (15.0, 16.0), (16.0, 16.0), (15.999, 16.001), (31.0, 32.0), (14.9999, 15.0001),
                   (-0.5, 0.5), (-16.0, -1.0), (-17.0, -16.0), (-1.0, -1.0), (-0.0001, 0.0), (47.0, 47.0), (47.00001, 47.5),
                   (3.25, 100.75), (np.nextafter(f32(15), f32(16)), 20.0), (np.nextafter(f32(15), f32(0)), 15.0),
                   (np.nextafter(f32(31), f32(32)), 31.5), (5.0, np.nextafter(f32(16), f32(0))), (5.0, 16.0)]:
        _check(lo, hi, T)


def test_tile_span_random():
    rng = np.random.default_rng(0)
    centres = rng.uniform(-200, 4200, 4000).astype(f32)
    halves = np.abs(rng.standard_cauchy(4000)).astype(f32) * f32(3.0) + f32(0.011)
    for c, h in zip(centres, halves):
        lo, hi = f32(c - h), f32(c + h)
        t_lo, t_hi = int(np.floor(lo / 16)) - 2, int(np.floor(hi / 16)) + 3
        _check(lo, hi, range(max(t_lo, -20), min(t_hi, 300)))
    # interval ends snapped onto / next to multiples of 16 and onto 16 t + 15
    for base in rng.integers(-3, 260, 300):
        for d_lo in (f32(15.0), np.nextafter(f32(15), f32(16)), np.nextafter(f32(15), f32(0)), f32(0.0), f32(16.0)):
            for d_hi in (f32(0.0), np.nextafter(f32(16), f32(0)), f32(16.0), f32(31.0)):
                lo, hi = f32(16 * base + d_lo), f32(16 * base + d_hi)
                _check(lo, hi, range(int(base) - 2, int(base) + 5))


def test_tile_span_extremes():
    """what the kernel relies on at the extremes: the span is intersected with the reference rect, whose tile indices fit
    16 bits, so only the behaviour below 2^16 matters -- and an infinite / huge upper end must not wrap (the clamp to 2^24
    before the + 1), a huge negative lower end must not either (the saturating conversion; here Python ints)"""
    assert tile_span(0.0, 3.0e38)[1] == (1 << 24) + 1
    t0, t1 = tile_span(-1e30, 1e30)
    assert t0 < 0 and t1 == (1 << 24) + 1
    for t in (0, 1, 65535):
        assert reached(-1e30, 1e30, t) and t0 <= t < t1
