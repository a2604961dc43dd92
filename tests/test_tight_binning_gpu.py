"""GSR_FLAG_TIGHT_BINNING (include/gsrast_b200.h): K1 stores, instead of the tile rect of the reference's 3-sigma square
(SURVEY Appendix A.2 step 9), its intersection with the tiles the Gaussian's {alpha >= 1/255} bounding box reaches.

What the flag may and may not change, each pinned here on the device through the C ABI:

  * radii, colour, depth, final_T ............ bit-identical to the literal lists (flags = 0), which the parity tests
                                               (tests/test_parity_gpu.py) pin to the oracle
  * the point list ............................ exactly the literal list minus the instances whose tile the box misses,
                                               same order; the expected sub-list is recomputed on the host from the
                                               literal list and the cull extents K1 stored (fp32, the kernel's op order)
  * n_contrib ................................. positions in the shorter list: maps onto the literal one
  * gradients ................................. equal to the literal path's within the atomics' reordering noise
  * batched forward == per-view forward, speculative capacity == exact size, as for the literal lists.
"""
import numpy as np
import pytest
import torch

from tests.util import cuda_backward, cuda_forward, rel_err, small_scene
from multiview_inpaint_b200 import scenes as S

pytestmark = pytest.mark.gpu
TIGHT, PRECISE = 32, 2


def _state(out, sc, cam, flags):
    from multiview_inpaint_b200 import _C
    n, color, radii, geom, binning, img, depth = out
    st = _C.unpack_state(sc["P"], cam.image_width, cam.image_height, n, geom, binning, img, flags)
    return {k: v.cpu().numpy() for k, v in st.items()}


def _expected_keep(st_lit, W):
    """Which instances of the literal list survive: the tile-level union of stage_entry's sub-tile tests."""
    gx = (W + 15) // 16
    r = st_lit["ranges"].view(np.uint32).astype(np.int64)
    lens = r[:, 1] - r[:, 0]
    # instances are stored tile after tile (ranges of non-empty tiles are contiguous and ascending)
    tile_of = np.repeat(np.arange(r.shape[0], dtype=np.int64), lens)
    gid = st_lit["point_list"].view(np.uint32).astype(np.int64)
    x, y = st_lit["means2D"][:, 0][gid], st_lit["means2D"][:, 1][gid]
    hx, hy = st_lit["cull"][:, 0][gid], st_lit["cull"][:, 1][gid]
    tx0 = ((tile_of % gx) * 16).astype(np.float32)
    ty0 = ((tile_of // gx) * 16).astype(np.float32)
    f32 = np.float32
    xlo, xhi, ylo, yhi = (x - hx).astype(f32), (x + hx).astype(f32), (y - hy).astype(f32), (y + hy).astype(f32)
    keep = (hx >= 0) & (xhi >= tx0) & (xlo <= tx0 + f32(15)) & (yhi >= ty0) & (ylo <= ty0 + f32(15))
    return keep, tile_of


def _scenes():
    # anisotropic, mixed-opacity splats of a few pixels; larger splats over a small image; many faint Gaussians
    yield "small", small_scene(4000, 160, 96, 3, 5, 7.0)
    yield "large_splats", small_scene(1500, 128, 128, 1, 9, 30.0)
    sc = small_scene(5000, 192, 112, 2, 13, 10.0)
    sc["opacities"] = (sc["opacities"] * 0.05).contiguous()          # most below 1/255 .. 0.05: small or empty boxes
    yield "faint", sc


@pytest.mark.parametrize("precise", [0, PRECISE])
@pytest.mark.parametrize("name,sc", list(_scenes()), ids=[n for n, _ in _scenes()])
def test_tight_lists_are_the_filtered_literal_lists(name, sc, precise):
    lit, d, cam, bg = cuda_forward(sc, flags=precise)
    tig, _, _, _ = cuda_forward(sc, flags=precise | TIGHT)
    W, H = cam.image_width, cam.image_height
    assert torch.equal(lit[2], tig[2]), "radii"
    assert torch.equal(lit[1], tig[1]), "colour must be bit-identical"
    assert torch.equal(lit[6], tig[6]), "depth must be bit-identical"
    a, b = _state(lit, sc, cam, precise), _state(tig, sc, cam, precise | TIGHT)
    np.testing.assert_array_equal(a["final_T"].view(np.uint32), b["final_T"].view(np.uint32))
    keep, tile_of = _expected_keep(a, W)
    assert 0 < keep.sum() < keep.size, "the scene should exercise the cull"
    assert tig[0] == int(keep.sum()), (tig[0], int(keep.sum()), lit[0])
    np.testing.assert_array_equal(b["point_list"].view(np.uint32), a["point_list"].view(np.uint32)[keep])
    G = a["ranges"].shape[0]
    cnt = np.bincount(tile_of[keep], minlength=G)
    end = np.cumsum(cnt)
    exp = np.stack([end - cnt, end], 1).astype(np.uint32)
    exp[cnt == 0] = 0
    np.testing.assert_array_equal(b["ranges"].view(np.uint32), exp)
    # n_contrib: position of the last contributor, 1-based, in the tile's list -- the same instance in both lists
    gx = (W + 15) // 16
    pos_lit = np.arange(keep.size) - np.repeat(a["ranges"].view(np.uint32)[:, 0].astype(np.int64),
                                               (a["ranges"].view(np.uint32)[:, 1] - a["ranges"].view(np.uint32)[:, 0]).astype(np.int64))
    kept_before = np.cumsum(keep) - keep            # kept instances in front of each literal instance (global)
    py, px = np.mgrid[0:H, 0:W]
    t = (py // 16) * gx + px // 16
    nl = a["n_contrib"].view(np.uint32).astype(np.int64)
    nt = b["n_contrib"].view(np.uint32).astype(np.int64)
    has = nl > 0
    assert ((nt > 0) == has).all()
    gl = a["ranges"].view(np.uint32)[:, 0].astype(np.int64)[t] + nl - 1          # literal global index of the last contributor
    assert keep[gl[has]].all(), "a contributing instance was dropped"
    gt = exp[:, 0].astype(np.int64)[t] + nt - 1
    np.testing.assert_array_equal(kept_before[gl[has]], gt[has])
    assert (pos_lit[gl[has]] == nl[has] - 1).all()


@pytest.mark.parametrize("name,sc", list(_scenes()), ids=[n for n, _ in _scenes()])
def test_tight_gradients_match_literal(name, sc):
    W, H = sc["W"], sc["H"]
    wt = S.loss_weights(W, H, 3)
    lit, d, cam, bg = cuda_forward(sc, flags=0)
    tig, _, _, _ = cuda_forward(sc, flags=TIGHT)
    ga = cuda_backward(lit, d, cam, bg, sc, wt, flags=0)
    gb = cuda_backward(tig, d, cam, bg, sc, wt, flags=TIGHT)
    for k in ga:
        a, b = ga[k].cpu().numpy(), gb[k].cpu().numpy()
        assert np.isfinite(b).all(), k
        assert rel_err(b, a) < 5e-4, (k, rel_err(b, a))   # same pairs, same kernels: the order of the sums differs (batch boundaries, REDs)


def test_tight_batched_equals_single_and_speculative():
    from multiview_inpaint_b200 import _C
    from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings
    sc = small_scene(6000, 160, 96, 3, 21, 7.0)
    dev = "cuda"
    g = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    cams = [c.to(dev) for c in S.orbit_cameras(3, 160, 96, max_deg=12.0)]
    e = torch.empty(0, device=dev)
    bg = torch.zeros(3, device=dev)
    rss = [GaussianRasterizationSettings(image_height=c.image_height, image_width=c.image_width, tanfovx=c.tanfovx, tanfovy=c.tanfovy,
                                         bg=bg, scale_modifier=1.0, viewmatrix=c.world_view_transform, projmatrix=c.full_proj_transform,
                                         sh_degree=3, campos=c.camera_center, prefiltered=False) for c in cams]
    singles = [_C.rasterize_gaussians(rs.bg, g["means3D"], e, g["opacities"], g["scales"], g["rotations"], 1.0, e, rs.viewmatrix,
                                      rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, g["shs"], 3,
                                      rs.campos, False, flags=TIGHT, capacity=0) for rs in rss]
    lits = [_C.rasterize_gaussians(rs.bg, g["means3D"], e, g["opacities"], g["scales"], g["rotations"], 1.0, e, rs.viewmatrix,
                                   rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, g["shs"], 3,
                                   rs.campos, False, flags=0, capacity=0) for rs in rss]
    assert all(s[0] < l[0] for s, l in zip(singles, lits))
    res = torch.zeros(len(rss), 2, dtype=torch.int64).pin_memory()
    caps = [int(s[0] * 1.25) + 64 for s in singles]
    outs = _C.forward_views(bg, g["means3D"], e, g["opacities"], g["scales"], g["rotations"], 1.0, e, rss, g["shs"], 3, False, caps,
                            [res[k] for k in range(len(rss))], flags=TIGHT)
    torch.cuda.synchronize()
    for k, (s, o) in enumerate(zip(singles, outs)):
        assert int(res[k, 0]) == s[0] and int(res[k, 1]) == 0
        assert torch.equal(s[1], o[1]) and torch.equal(s[2], o[2]) and torch.equal(s[6], o[6])
        a = _C.unpack_state(sc["P"], 160, 96, s[0], s[3], s[4], s[5], TIGHT)
        b = _C.unpack_state(sc["P"], 160, 96, s[0], o[3], o[4], o[5], TIGHT)
        for key in ("point_list", "ranges", "n_contrib", "final_T", "tiles_touched"):
            assert torch.equal(a[key], b[key]), key
        # speculative capacity (one call, capacity > N) gives the same as the exact-size path
        sp = _C.rasterize_gaussians(rss[k].bg, g["means3D"], e, g["opacities"], g["scales"], g["rotations"], 1.0, e, rss[k].viewmatrix,
                                    rss[k].projmatrix, rss[k].tanfovx, rss[k].tanfovy, 96, 160, g["shs"], 3, rss[k].campos, False,
                                    flags=TIGHT, capacity=caps[k])
        assert sp[0] == s[0] and torch.equal(sp[1], s[1])


def test_default_flags_of_the_python_layers_are_tight():
    from multiview_inpaint_b200 import _C
    import os
    if "GSR_FLAGS" not in os.environ:
        assert _C.DEFAULT_FLAGS == _C.FLAG_TIGHT_BINNING == TIGHT
    sc = small_scene(3000, 96, 80, 3, 11, 6.0)
    a, _, _, _ = cuda_forward(sc, flags=None)
    b, _, _, _ = cuda_forward(sc, flags=0)
    assert a[0] < b[0] and torch.equal(a[1], b[1])
