"""Size-independent properties of the CPU oracle (the checker of the GPU parity tests), on seeded random scenes:
what must hold for ANY correct implementation of SURVEY.md Appendix A, so a slip in the restatement shows up even
where no golden vector exists ("parity unpinned")."""
import numpy as np
import pytest
import torch

from multiview_inpaint_b200 import scenes as S
from tests.util import oracle_forward, small_scene


def _permuted(sc, perm):
    out = dict(sc)
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        out[k] = sc[k][perm].contiguous()
    return out


@pytest.mark.parametrize("seed", [3, 4])
def test_gaussian_order_does_not_matter(oracle, seed):
    """The image depends on the SET of Gaussians: lists are ordered by (depth, index), so with distinct depths a
    permutation of the inputs permutes radii / point ids and leaves every pixel bit-identical."""
    sc = small_scene(1500, 96, 64, 1, seed, 7.0)
    f = oracle_forward(oracle, sc)
    d = f.depths[f.radii > 0]
    assert len(np.unique(d.view(np.uint32))) == len(d), "scene has depth ties: pick another seed"
    perm = torch.from_numpy(np.random.default_rng(seed).permutation(sc["P"]))
    g = oracle_forward(oracle, _permuted(sc, perm))
    np.testing.assert_array_equal(g.radii, f.radii[perm.numpy()])
    assert g.num_rendered == f.num_rendered
    np.testing.assert_array_equal(g.color.view(np.uint32), f.color.view(np.uint32))
    np.testing.assert_array_equal(g.depth.view(np.uint32), f.depth.view(np.uint32))
    np.testing.assert_array_equal(g.n_contrib, f.n_contrib)
    np.testing.assert_array_equal(perm.numpy()[g.point_list], f.point_list)          # same Gaussians in the same list slots


def test_background_enters_linearly_through_final_T(oracle):
    sc = small_scene(1200, 80, 48, 0, 9, 6.0)
    f0 = oracle_forward(oracle, sc, bg=np.zeros(3, np.float32))
    bg = np.array([0.25, 0.5, 1.0], np.float32)
    f1 = oracle_forward(oracle, sc, bg=bg)
    np.testing.assert_array_equal(f1.final_T, f0.final_T)
    np.testing.assert_array_equal(f1.n_contrib, f0.n_contrib)
    np.testing.assert_allclose(f1.color, f0.color + f0.final_T[None] * bg[:, None, None], rtol=0, atol=1e-6)
    assert ((f0.final_T >= 0) & (f0.final_T <= 1)).all()


def test_invisible_gaussians_change_nothing(oracle):
    """Gaussians behind the camera, beyond the frustum guard band or with alpha below 1/255 everywhere may be added
    or removed freely: radii 0 (or no contribution), identical image."""
    sc = small_scene(800, 64, 64, 1, 21, 6.0)
    f = oracle_forward(oracle, sc)
    extra = 50
    g = torch.Generator().manual_seed(1)
    add = dict(means3D=torch.cat([sc["means3D"], torch.tensor([[0.0, 0.0, -3.0]]).repeat(extra, 1) + torch.randn(extra, 3, generator=g) * 0.1]),
               scales=torch.cat([sc["scales"], torch.full((extra, 3), 0.05)]),
               rotations=torch.cat([sc["rotations"], torch.tensor([[1.0, 0, 0, 0]]).repeat(extra, 1)]),
               opacities=torch.cat([sc["opacities"], torch.full((extra, 1), 0.9)]),
               shs=torch.cat([sc["shs"], torch.zeros(extra, sc["shs"].shape[1], 3)]))
    sc2 = dict(sc, **add, P=sc["P"] + extra)
    f2 = oracle_forward(oracle, sc2)
    assert (f2.radii[sc["P"]:] == 0).all() and f2.num_rendered == f.num_rendered
    np.testing.assert_array_equal(f2.color.view(np.uint32), f.color.view(np.uint32))
    np.testing.assert_array_equal(f2.point_list, f.point_list)


def test_zero_loss_weights_give_zero_gradients_and_linearity(oracle):
    """The backward is linear in dL/dcolor: g(a w1 + b w2) = a g(w1) + b g(w2) (up to fp32 summation), g(0) = 0."""
    sc = small_scene(600, 64, 48, 1, 33, 6.0)
    f = oracle_forward(oracle, sc)
    w1 = S.loss_weights(64, 48, 1).numpy()
    w2 = S.loss_weights(64, 48, 2).numpy()
    g0 = oracle.backward(f, np.zeros_like(w1))
    for k, v in g0.items():
        assert not np.any(v), k
    g1, g2, g12 = oracle.backward(f, w1), oracle.backward(f, w2), oracle.backward(f, (2.0 * w1 - 0.5 * w2).astype(np.float32))
    for k in ("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dmeans2D"):
        want = 2.0 * g1[k].astype(np.float64) - 0.5 * g2[k].astype(np.float64)
        scale = np.abs(want).max() + 1e-30
        assert np.abs(g12[k] - want).max() <= 2e-5 * scale, k


def test_depth_ties_break_by_gaussian_index(oracle):
    """The domain's collisions: Gaussians at exactly the same depth (clones made by densification sit on top of their
    originals, gaussian_model.py:456-464) have equal sort keys; a stable sort keeps them in index order inside every
    tile list, and the two-level scheme (depth sort, emission in that order, stable tile sort) must give the very
    same lists as the single stable (tile | depth) sort."""
    sc = small_scene(600, 64, 48, 1, 13, 8.0)
    # every Gaussian three times: same position / shape, different colours and opacities
    rep = lambda t: torch.cat([t, t, t]).contiguous()
    g = torch.Generator().manual_seed(2)
    sc3 = dict(sc, means3D=rep(sc["means3D"]), scales=rep(sc["scales"]), rotations=rep(sc["rotations"]),
               opacities=torch.rand(3 * sc["P"], 1, generator=g) * 0.6 + 0.05,
               shs=torch.randn(3 * sc["P"], sc["shs"].shape[1], 3, generator=g) * 0.3, P=3 * sc["P"])
    f = oracle_forward(oracle, sc3)
    P = sc["P"]
    assert (f.depths[:P] == f.depths[P:2 * P]).all() and (f.radii[:P] == f.radii[2 * P:]).all()
    depth_bits = f.depths.view(np.uint32)[f.point_list].astype(np.int64)
    ids = f.point_list.astype(np.int64)
    n_ties = 0
    for t in range(f.ranges.shape[0]):
        a, b = int(f.ranges[t, 0]), int(f.ranges[t, 1])
        d, i = depth_bits[a:b], ids[a:b]
        assert (np.diff(d) >= 0).all()
        tie = np.diff(d) == 0
        assert (np.diff(i)[tie] > 0).all(), f"tile {t}: equal depths not in index order"
        n_ties += int(tie.sum())
    assert n_ties >= 2 * (f.radii[:P] > 0).sum()                   # each visible triple ties twice in every tile it touches
    # two-level == single sort (what DESIGN.md section 4 claims), with ties
    dkeys = np.where(f.radii > 0, f.depths.view(np.uint32), np.uint32(0xFFFFFFFF))
    order = np.argsort(dkeys, kind="stable")
    tiles, vals = [], []
    for gid in order:
        if f.radii[gid] <= 0:
            continue
        sel = f.values_unsorted == gid
        tiles.append((f.keys_unsorted[sel] >> 32).astype(np.int64))
        vals.append(f.values_unsorted[sel])
    tiles, vals = np.concatenate(tiles), np.concatenate(vals)
    np.testing.assert_array_equal(vals[np.argsort(tiles, kind="stable")], f.point_list)
    # the image only depends on the order through the blend: finite and bounded
    assert np.isfinite(f.color).all() and (f.final_T >= 0).all()


@pytest.mark.parametrize("seed,W,H,rad", [(41, 96, 64, 7.0), (42, 200, 120, 25.0), (43, 40, 24, 4.0)])
def test_radius_rect_and_keys_against_an_independent_numpy_restatement(oracle, seed, W, H, rad):
    """SURVEY Appendix A.2 steps 6-9 and A.3 restated a second time in numpy from the oracle's OWN float64 torch twin
    (oracle/torch_golden.py gives the 2-D covariance): radius = ceil(3 sqrt(lambda_max)) with the 0.1 floor, the tile
    rect with C truncation and clamping, tiles_touched, and the (tile << 32 | depth bits) keys in row-major rect order.
    Integers must agree exactly wherever the float32 / float64 radius does not sit on a ceil() boundary."""
    from oracle import torch_golden as TG
    sc = small_scene(1500, W, H, 1, seed, rad)
    cam = sc["camera"]
    f = oracle_forward(oracle, sc)
    dd = torch.float64
    P = sc["P"]
    pix, conic, opac, rgb, tz = TG.preprocess(sc["means3D"].to(dd), torch.zeros(P, 3, dtype=dd), sc["opacities"].to(dd),
                                              sc["scales"].to(dd), sc["rotations"].to(dd), sc["shs"].to(dd), 1,
                                              cam.world_view_transform, cam.full_proj_transform, cam.camera_center, W, H,
                                              cam.tanfovx, cam.tanfovy)
    cn = conic.numpy()
    det_c = cn[:, 0] * cn[:, 2] - cn[:, 1] ** 2                      # conic = cov^-1: cov = adj(conic) / det(conic)
    a, c, b = cn[:, 2] / det_c, cn[:, 0] / det_c, -cn[:, 1] / det_c
    mid = 0.5 * (a + c)
    lam = mid + np.sqrt(np.maximum(0.1, mid * mid - (a * c - b * b)))
    r64 = 3.0 * np.sqrt(lam)
    radius = np.ceil(r64)
    safe = (np.abs(r64 - np.round(r64)) > 1e-4) & (np.abs(tz.numpy() - 0.2) > 1e-6)   # away from ceil() and near-plane boundaries
    safe |= tz.numpy() < 0.2 - 1e-6                                                  # culled: everything is 0 whatever the rest
    gx, gy = (W + 15) // 16, (H + 15) // 16
    px, py = pix.numpy()[:, 0], pix.numpy()[:, 1]
    trunc = lambda v: np.trunc(v).astype(np.int64)
    x0 = np.clip(trunc((px - radius) / 16), 0, gx)
    x1 = np.clip(trunc((px + radius + 15) / 16), 0, gx)
    y0 = np.clip(trunc((py - radius) / 16), 0, gy)
    y1 = np.clip(trunc((py + radius + 15) / 16), 0, gy)
    # pixel positions a hair away from a multiple of 16 could truncate differently in float32: exclude those too
    frac = lambda v: np.abs(v / 16 - np.round(v / 16))
    safe &= (frac(px - radius) > 1e-5) & (frac(px + radius + 15) > 1e-5) & (frac(py - radius) > 1e-5) & (frac(py + radius + 15) > 1e-5)
    tiles = (x1 - x0) * (y1 - y0)
    culled = tz.numpy() <= 0.2
    want_tiles = np.where(culled, 0, tiles)
    want_radii = np.where(culled | (tiles == 0), 0, radius).astype(np.int64)
    assert safe.mean() > 0.95
    np.testing.assert_array_equal(f.tiles_touched[safe].astype(np.int64), want_tiles[safe])
    np.testing.assert_array_equal(f.radii[safe].astype(np.int64), want_radii[safe])
    # keys: row-major over the rect, tile id in the high word, the float32 depth's bit pattern in the low word
    offs = np.concatenate([[0], np.cumsum(f.tiles_touched.astype(np.int64))])
    np.testing.assert_array_equal(f.point_offsets.astype(np.int64), offs[1:])
    checked = 0
    for i in np.nonzero(safe & (f.radii > 0))[0][:300]:
        ks = f.keys_unsorted[offs[i]:offs[i + 1]]
        want = [((y * gx + x) << 32) | int(f.depths[i:i + 1].view(np.uint32)[0]) for y in range(y0[i], y1[i]) for x in range(x0[i], x1[i])]
        np.testing.assert_array_equal(ks.astype(np.uint64), np.array(want, dtype=np.uint64))
        assert (f.values_unsorted[offs[i]:offs[i + 1]] == i).all()
        checked += 1
    assert checked > 100


@pytest.mark.parametrize("seed,W,H,rad,P", [(51, 64, 48, 9.0, 900), (52, 40, 24, 5.0, 40)])
def test_blend_outputs_against_an_independent_per_pixel_restatement(oracle, seed, W, H, rad, P):
    """SURVEY Appendix A.5 a second time, pixel by pixel in float64 numpy from the oracle's own geometry state and
    lists: colour, final_T, n_contrib (index of the last contributor) and the w-depth fork's MEDIAN depth (depth of
    the Gaussian whose blending takes T across 0.5; 15.0 where none does -- gen_seq.py:50).  The float32 oracle may
    differ from float64 only where a threshold (1/255, 1e-4, 0.5) is hit within rounding: a handful of pixels."""
    sc = small_scene(P, W, H, 1, seed, rad)
    bg = np.array([0.3, 0.6, 0.9], np.float32)
    f = oracle_forward(oracle, sc, bg=bg)
    gx = (W + 15) // 16
    xy, co, rgb, dep = f.means2D.astype(np.float64), f.conic_opacity.astype(np.float64), f.rgb.astype(np.float64), f.depths
    color = np.zeros((3, H, W)); final_T = np.ones((H, W)); n_contrib = np.zeros((H, W), np.int64)
    depth = np.full((H, W), 15.0, np.float32)
    fragile = np.zeros((H, W), bool)                       # a threshold was closer than float32 rounding
    for y in range(H):
        for x in range(W):
            t = (y // 16) * gx + x // 16
            T, C, last = 1.0, np.zeros(3), 0
            for k, j in enumerate(f.point_list[f.ranges[t, 0]:f.ranges[t, 1]], start=1):
                dx, dy = xy[j, 0] - x, xy[j, 1] - y
                power = -0.5 * (co[j, 0] * dx * dx + co[j, 2] * dy * dy) - co[j, 1] * dx * dy
                if power > 0:
                    continue
                alpha = min(0.99, co[j, 3] * np.exp(power))
                if abs(alpha - 1 / 255) < 1e-6:
                    fragile[y, x] = True
                if alpha < 1 / 255:
                    continue
                test_T = T * (1 - alpha)
                if abs(test_T - 1e-4) < 1e-8 or abs(test_T - 0.5) < 1e-6:
                    fragile[y, x] = True
                if test_T < 1e-4:
                    break
                C += rgb[j] * alpha * T
                if T > 0.5 and test_T < 0.5:
                    depth[y, x] = dep[j]
                T, last = test_T, k
            color[:, y, x] = C + T * bg
            final_T[y, x], n_contrib[y, x] = T, last
    ok = ~fragile
    assert ok.mean() > 0.98
    assert np.abs(color - f.color)[:, ok].max() < 2e-6
    assert np.abs(final_T - f.final_T)[ok].max() < 1e-6
    np.testing.assert_array_equal(n_contrib[ok], f.n_contrib[ok])
    np.testing.assert_array_equal(depth[ok], f.depth[0][ok])
    assert (f.depth[0] != np.float32(15.0)).any()
    if P < 100:
        assert (f.depth[0] == np.float32(15.0)).any()        # the sparse scene leaves pixels no Gaussian takes across T = 0.5
