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


# ---- the claim itself, on the oracle: instances whose tile the {alpha >= 1/255} bounding box misses contribute nothing ----
def _tight_keep(f, slack=0.05):
    """float32 restatement of K1's cull extents (csrc/preprocess.cu) and of the tile-level union of the staging tests"""
    G = f.ranges.shape[0]
    gx = (f.W + 15) // 16
    lens = (f.ranges[:, 1] - f.ranges[:, 0]).astype(np.int64)
    tile_of = np.repeat(np.arange(G, dtype=np.int64), lens)
    gid = f.point_list.astype(np.int64)
    co = f.conic_opacity.astype(np.float64)
    a, b, c, op = co[:, 0], co[:, 1], co[:, 2], co[:, 3]
    det = a * c - b * b
    with np.errstate(divide="ignore", invalid="ignore"):
        cov_xx, cov_yy = (c / det).astype(f32), (a / det).astype(f32)      # the 2-D covariance the conic was inverted from
        tau = np.log((f32(255.0) * op.astype(f32)).astype(f32)).astype(f32)
        taus = (tau + f32(slack)).astype(f32)
        hx = (np.sqrt((f32(2.0) * taus * cov_xx).astype(f32)) * f32(1.001) + f32(0.01)).astype(f32)
        hy = (np.sqrt((f32(2.0) * taus * cov_yy).astype(f32)) * f32(1.001) + f32(0.01)).astype(f32)
    dead = ~((tau > 0) & (op > 0))
    hx[dead], hy[dead] = f32(-1.0), f32(-1.0)
    x, y = f.means2D[:, 0][gid], f.means2D[:, 1][gid]
    ex, ey = hx[gid], hy[gid]
    tx0, ty0 = ((tile_of % gx) * 16).astype(f32), ((tile_of // gx) * 16).astype(f32)
    keep = (ex >= 0) & ((x + ex).astype(f32) >= tx0) & ((x - ex).astype(f32) <= tx0 + f32(15)) & \
           ((y + ey).astype(f32) >= ty0) & ((y - ey).astype(f32) <= ty0 + f32(15))
    return keep, tile_of


def _scenes():
    from tests.util import small_scene
    yield "small", small_scene(4000, 160, 96, 3, 5, 7.0)
    yield "large_splats", small_scene(1500, 128, 128, 1, 9, 30.0)
    sc = small_scene(5000, 192, 112, 2, 13, 10.0)
    sc["opacities"] = (sc["opacities"] * 0.05).contiguous()
    yield "faint", sc


import pytest  # noqa: E402


@pytest.mark.parametrize("name", ["small", "large_splats", "faint"])
def test_oracle_image_and_gradients_do_not_see_the_dropped_instances(oracle, name):
    """The C oracle blends the literal lists and the lists without the instances tight binning drops: colour, depth, final_T
    bit-identical, n_contrib pointing at the same instance, gradients equal -- the property K1's shrunk rect relies on,
    established on the restatement of the REFERENCE algorithm, with no device involved."""
    import copy
    from multiview_inpaint_b200 import scenes as S
    from tests.util import oracle_forward
    sc = dict(_scenes())[name]
    f = oracle_forward(oracle, sc)
    keep, tile_of = _tight_keep(f)
    assert 0 < keep.sum() < keep.size
    G = f.ranges.shape[0]
    cnt = np.bincount(tile_of[keep], minlength=G)
    end = np.cumsum(cnt)
    g = copy.copy(f)
    g.inputs = dict(f.inputs)
    g.point_list = f.point_list[keep].copy()
    g.ranges = np.stack([end - cnt, end], 1).astype(np.uint32)
    g.ranges[cnt == 0] = 0
    g.num_rendered = int(keep.sum())
    oracle.blend(g, sc["bg"].numpy())
    np.testing.assert_array_equal(g.color.view(np.uint32), f.color.view(np.uint32))
    np.testing.assert_array_equal(g.depth.view(np.uint32), f.depth.view(np.uint32))
    np.testing.assert_array_equal(g.final_T.view(np.uint32), f.final_T.view(np.uint32))
    # n_contrib: the same last contributor, at its position in the shorter list
    gx = (f.W + 15) // 16
    py, px = np.mgrid[0:f.H, 0:f.W]
    t = (py // 16) * gx + px // 16
    nl, nt = f.n_contrib.astype(np.int64), g.n_contrib.astype(np.int64)
    has = nl > 0
    assert ((nt > 0) == has).all()
    lit_idx = f.ranges[:, 0].astype(np.int64)[t] + nl - 1
    assert keep[lit_idx[has]].all()
    kept_before = np.cumsum(keep) - keep
    np.testing.assert_array_equal(kept_before[lit_idx[has]], (g.ranges[:, 0].astype(np.int64)[t] + nt - 1)[has])
    # gradients: the same pairs contribute
    wt = S.loss_weights(f.W, f.H, 3).numpy()
    ga, gb = oracle.backward(f, wt), oracle.backward(g, wt)
    for k in ga:
        scale = np.abs(ga[k]).max() + 1e-30
        assert np.abs(ga[k] - gb[k]).max() / scale < 1e-6, k
