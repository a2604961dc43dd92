"""Batched multi-view forward / blend backward (gsr_forward_views, gsr_backward_blend_views): ONE launch per pipeline
stage for all views of a step.  The per-view entry point gsr_forward runs the same kernels with one view, and THAT path
is pinned bit-exactly to the oracle by tests/test_parity_gpu.py; here the batched launches must reproduce it bit for bit
on every integer output and on colour / depth (same arithmetic, same order), for views that share a size and for views
that do not (different tile grids, different tile-id widths in one launch), and with a capacity that is too small
(overflow reported per view, no out-of-bounds write)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from multiview_inpaint_b200 import scenes as S
from tests.util import rel_err, small_scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _settings(cam, bg, deg):
    from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings
    return GaussianRasterizationSettings(image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx,
                                         tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
                                         projmatrix=cam.full_proj_transform, sh_degree=deg, campos=cam.camera_center,
                                         prefiltered=False)


def _single(g, rs, flags=0):
    from multiview_inpaint_b200 import _C
    e = torch.empty(0, device=DEV)
    return _C.rasterize_gaussians(rs.bg, g["means3D"], e, g["opacities"], g["scales"], g["rotations"], rs.scale_modifier, e,
                                  rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width,
                                  g["shs"], rs.sh_degree, rs.campos, rs.prefiltered, flags=flags, capacity=0)


def _batched(g, rss, caps, flags=0, workspaces=None):
    from multiview_inpaint_b200 import _C
    e = torch.empty(0, device=DEV)
    res = torch.zeros(len(rss), 2, dtype=torch.int64).pin_memory()
    outs = _C.forward_views(rss[0].bg, g["means3D"], e, g["opacities"], g["scales"], g["rotations"], 1.0, e, rss, g["shs"],
                            rss[0].sh_degree, False, caps, [res[k] for k in range(len(rss))], workspaces=workspaces, flags=flags)
    torch.cuda.synchronize()
    return outs, res


def _scene(P=6000, W=160, H=96, deg=3, seed=21):
    sc = small_scene(P, W, H, deg, seed, 7.0)
    g = {k: sc[k].to(DEV) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    return sc, g


def _compare(P, rs, ref, out, n):
    from multiview_inpaint_b200 import _C
    W, H = rs.image_width, rs.image_height
    assert torch.equal(ref[2], out[2]), "radii"
    assert torch.equal(ref[1], out[1]), "colour"
    assert torch.equal(ref[6], out[6]), "depth"
    a = _C.unpack_state(P, W, H, n, ref[3], ref[4], ref[5], 0)
    b = _C.unpack_state(P, W, H, n, out[3], out[4], out[5], 0)
    V = int((ref[2] > 0).sum())
    for k in ("point_list", "ranges", "final_T", "n_contrib", "tiles_touched"):
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(a["order"][:V], b["order"][:V]) and torch.equal(a["point_offsets"][:V], b["point_offsets"][:V])
    vis = (ref[2] > 0)
    for k in ("means2D", "conic_opacity", "rgb", "depths", "clamped"):
        assert torch.equal(a[k][vis], b[k][vis]), k


@pytest.mark.parametrize("flags", [0, 2])
@pytest.mark.parametrize("nv", [1, 3, 4, 8, 11])
def test_batched_forward_is_bit_identical_to_the_per_view_forward(nv, flags):
    sc, g = _scene()
    bg = torch.tensor([0.1, 0.2, 0.3], device=DEV)
    cams = [c.to(DEV) for c in S.orbit_cameras(nv, sc["W"], sc["H"], max_deg=12.0)]
    rss = [_settings(c, bg, sc["sh_degree"]) for c in cams]
    refs = [_single(g, rs, flags) for rs in rss]
    caps = [int(r[0] * 1.3) + 100 for r in refs]
    outs, res = _batched(g, rss, caps, flags)
    for k in range(nv):
        assert int(res[k, 0]) == refs[k][0] and int(res[k, 1]) == 0
        _compare(sc["P"], rss[k], refs[k], outs[k], refs[k][0])


def test_batched_forward_with_views_of_different_sizes():
    """one launch, three tile grids: 10x6 (6-bit tile ids), 40x23 (10 bits: two tile-sort passes), 3x2 partial tiles"""
    sc, g = _scene(P=5000, W=160, H=96, deg=1, seed=5)
    bg = torch.zeros(3, device=DEV)
    sizes = [(160, 96), (640, 360), (40, 24)]
    rss = [_settings(S.orbit_cameras(3, W, H, max_deg=8.0)[k].to(DEV), bg, 1) for k, (W, H) in enumerate(sizes)]
    refs = [_single(g, rs) for rs in rss]
    outs, res = _batched(g, rss, [int(r[0] * 1.2) + 64 for r in refs])
    for k in range(3):
        assert int(res[k, 0]) == refs[k][0] and int(res[k, 1]) == 0
        _compare(sc["P"], rss[k], refs[k], outs[k], refs[k][0])


def test_batched_forward_reports_capacity_overflow_per_view_and_stays_in_bounds():
    from multiview_inpaint_b200 import _C
    sc, g = _scene(P=8000, seed=9)
    bg = torch.zeros(3, device=DEV)
    rss = [_settings(c.to(DEV), bg, 3) for c in S.orbit_cameras(3, sc["W"], sc["H"], max_deg=10.0)]
    refs = [_single(g, rs) for rs in rss]
    caps = [refs[0][0] + 10, max(refs[1][0] // 3, 1), refs[2][0] + 10]      # the middle view is too small
    wss = [_C.Workspace(DEV) for _ in rss]
    for w in wss:                                                            # recycled, dirty memory behind every buffer
        for name in ("geom", "binning", "image"):
            w.bytes(name, 1 << 22).fill_(0xAB)
    outs, res = _batched(g, rss, caps, workspaces=wss)
    assert [int(res[k, 0]) for k in range(3)] == [r[0] for r in refs]
    assert (int(res[1, 1]) >> 32) != 0 or int(res[1, 0]) > caps[1]
    for k in (0, 2):
        assert int(res[k, 1]) == 0
        _compare(sc["P"], rss[k], refs[k], outs[k], refs[k][0])
    for k, w in enumerate(wss):   # nothing was written past the logical end of the binning buffer
        lay = _C.get_layout(sc["P"], sc["W"], sc["H"], caps[k], 0)
        tail = w.bufs["binning"][lay.binning_bytes:]
        assert bool((tail == 0xAB).all())


@pytest.mark.parametrize("nv", [2, 4, 5])
def test_batched_step_gradients_match_the_per_view_step(nv):
    """whole multi-view step: batched front end + one blend-backward launch + batched K8+K9 against the view-by-view
    accumulate path (same kernels, different launch shape: only the order of the fp32 atomics differs)"""
    from multiview_inpaint_b200 import multiview as mv
    sc, g = _scene(P=7000, seed=33)
    P, M = sc["P"], g["shs"].shape[1]
    bg = torch.zeros(3, device=DEV)
    cams = [c.to(DEV) for c in S.orbit_cameras(nv, sc["W"], sc["H"], max_deg=10.0)]
    rss = [_settings(c, bg, sc["sh_degree"]) for c in cams]
    wts = [S.loss_weights(sc["W"], sc["H"], 40 + k).to(DEV) for k in range(nv)]
    fns = [lambda c, w=w: w for w in wts]
    a_ref = mv.GradArena(P, M, DEV)
    for rs, fn in zip(rss, fns):
        mv.cuda_view_fwd_bwd(g, rs, fn, a_ref, capacity=0)
    av = mv.AsyncViews(nv)
    for k, rs in enumerate(rss):
        av.learn(k, _single(g, rs)[0])
    a_new = mv.GradArena(P, M, DEV)
    a_new.flat.fill_(7.0)   # the batched K8+K9 writes the arena: stale contents must not leak
    states = mv.cuda_views_fwd_bwd(g, rss, fns, a_new, capacities=[av.capacity(k) for k in range(nv)],
                                   async_results=[av.slot(k) for k in range(nv)], batched=True)
    torch.cuda.synchronize()
    assert not av.check(range(nv))
    assert len(states) == nv
    for name in a_ref.views:
        assert rel_err(a_new.views[name].cpu().numpy(), a_ref.views[name].cpu().numpy()) < 1e-4, name
    assert torch.equal(a_new.visible_count, a_ref.visible_count) and torch.equal(a_new.max_radii, a_ref.max_radii)
    assert rel_err(a_new.grad_norm_accum.cpu().numpy(), a_ref.grad_norm_accum.cpu().numpy()) < 1e-4


def test_forward_clears_the_backward_accumulators():
    """gsr_view_forward.backward_scratch + GSR_FLAG_SCRATCH_CLEARED: K1 clears each view's 48-byte accumulator rows on the
    way, so the blend backward starts without a memset of its own.  The buffers are poisoned first (NaN bit pattern) and
    P is not a multiple of K1's 256-Gaussian CTAs; the accumulators must equal those of the path that clears them itself."""
    from multiview_inpaint_b200 import _C
    sc, g = _scene(P=7013, seed=17)
    P = sc["P"]
    bg = torch.zeros(3, device=DEV)
    nv = 3
    cams = [c.to(DEV) for c in S.orbit_cameras(nv, sc["W"], sc["H"], max_deg=10.0)]
    rss = [_settings(c, bg, sc["sh_degree"]) for c in cams]
    wts = [S.loss_weights(sc["W"], sc["H"], 60 + k).to(DEV) for k in range(nv)]
    caps = [int(_single(g, rs, _C.DEFAULT_FLAGS)[0] * 1.3) + 100 for rs in rss]
    e = torch.empty(0, device=DEV)

    def run(clear_in_forward):
        res = torch.zeros(nv, 2, dtype=torch.int64).pin_memory()
        wss = [_C.Workspace(DEV) for _ in range(nv)]
        nb = int(_C._lib.gsr_backward_scratch_bytes(P))
        for w in wss:
            w.bytes("scratch", nb).fill_(0xFF)
        cleared = [] if clear_in_forward else None
        outs = _C.forward_views(bg, g["means3D"], e, g["opacities"], g["scales"], g["rotations"], 1.0, e, rss, g["shs"],
                                sc["sh_degree"], False, caps, [res[k] for k in range(nv)], workspaces=wss, scratches_out=cleared)
        scr = _C.backward_blend_views(bg, wts, [o[3] for o in outs], [o[4] for o in outs], [o[5] for o in outs], P,
                                      workspaces=wss, scratches=cleared)
        torch.cuda.synchronize()
        assert int(res[:, 1].sum()) == 0
        if clear_in_forward:
            assert all(a.data_ptr() == b.data_ptr() for a, b in zip(scr, cleared))
        return [s.view(torch.float32)[:12 * P].clone() for s in scr]

    ref, new = run(False), run(True)
    for a, b in zip(ref, new):
        assert torch.isfinite(b).all()
        assert rel_err(b.cpu().numpy(), a.cpu().numpy()) < 1e-5
        assert torch.equal(a == 0, b == 0)   # rows nobody blended stay exactly zero
