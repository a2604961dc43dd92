"""GPU parity of the densification re-pack, through the C ABI: gsr_gather_rows (csrc/densify.cu) and
GaussianParamArena.densify_and_prune / prune_points / reset_opacity on top of it.

Checked against (1) the golden vectors produced by the reference's own GaussianModel (tests/golden/densify.npz, see
make_densify_golden.py): the arenas after the ONE gather launch must equal, bit for bit, the six parameter tensors and
their Adam moments the reference holds after its four cat / mask rounds; (2) torch indexing on fresh seeded inputs
including ragged sizes around the 64-row CTA tile and every SH width; (3) at the headline model size, a checksum
property.  Pure data movement: every comparison is exact.  reset_opacity is floating point (sigmoid, log on the
device vs the golden's CPU torch): 2e-6 absolute."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda"
GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
SLICES = ("_xyz", "_features", "_opacity", "_scaling", "_rotation")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "densify.npz"))


@pytest.fixture(scope="module")
def C():
    from multiview_inpaint_b200 import _C
    return _C


def arena_from_gold(gold, prefix):
    """GaussianParamArena holding the golden model `prefix` (parameters and both Adam moments)."""
    from multiview_inpaint_b200.trainstep import GaussianParamArena
    t = lambda k: torch.from_numpy(gold[f"{prefix}_{k}"].copy()).to(DEV)
    pa = GaussianParamArena.from_tensors(*(t(g) for g in GROUPS))
    for mom, flat in (("exp_avg", pa.exp_avg), ("exp_avg_sq", pa.exp_avg_sq)):
        pa._slice(flat, "_xyz").copy_(t(f"xyz_{mom}"))
        f = pa._slice(flat, "_features")
        f[:, :1].copy_(t(f"f_dc_{mom}"))
        if pa.M > 1:
            f[:, 1:].copy_(t(f"f_rest_{mom}"))
        pa._slice(flat, "_opacity").copy_(t(f"opacity_{mom}"))
        pa._slice(flat, "_scaling").copy_(t(f"scaling_{mom}"))
        pa._slice(flat, "_rotation").copy_(t(f"rotation_{mom}"))
    pa.step_count = 2
    return pa


def assert_arena_equals_gold(pa, gold, prefix, skip_rows=None):
    """bit-exact; `skip_rows` (bool mask over rows) excludes the split children where the samples differ"""
    keep = slice(None) if skip_rows is None else ~skip_rows
    want = lambda k: torch.from_numpy(gold[f"{prefix}_{k}"].copy()).to(DEV)
    assert pa.P == gold[f"{prefix}_xyz"].shape[0]
    views = {"": pa.param, "_exp_avg": pa.exp_avg, "_exp_avg_sq": pa.exp_avg_sq}
    for suffix, flat in views.items():
        got = dict(xyz=pa._slice(flat, "_xyz"), f_dc=pa._slice(flat, "_features")[:, :1],
                   f_rest=pa._slice(flat, "_features")[:, 1:], opacity=pa._slice(flat, "_opacity"),
                   scaling=pa._slice(flat, "_scaling"), rotation=pa._slice(flat, "_rotation"))
        for g in GROUPS:
            w = want(g + suffix)
            assert got[g].shape == w.shape, (g, suffix)
            assert torch.equal(got[g][keep], w[keep]), f"{prefix}: {g}{suffix} differs from the reference"


def cpu_plan(gold, case):
    """the plan made on the CPU with the golden's seed: the same torch.normal samples as the reference drew"""
    from multiview_inpaint_b200 import densify
    t = lambda k: torch.from_numpy(gold[f"{case}_in_{k}"].copy())
    max_grad, min_op, extent, mss, percent_dense, nseed = gold[f"{case}_args"].tolist()
    torch.manual_seed(int(nseed))
    return densify.plan_densify_and_prune(t("xyz"), t("scaling"), t("rotation"), t("opacity"), t("xyz_gradient_accum"),
                                          t("denom"), max_grad, min_op, extent, None if mss < 0 else int(mss),
                                          percent_dense=percent_dense)


# ------------------------------------------------------------------------------------------------ reference golden
@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_apply_plan_equals_reference_densify_and_prune(gold, case):
    from multiview_inpaint_b200 import densify
    pa = arena_from_gold(gold, f"{case}_in")
    densify.apply_plan(pa, cpu_plan(gold, case))
    torch.cuda.synchronize()
    assert_arena_equals_gold(pa, gold, f"{case}_out")
    assert pa.step_count == 2                     # the optimizer's step counter survives (state dict is re-keyed, :399-401)
    # activated buffers were re-made for the new size
    assert pa.scales.shape == (pa.P, 3) and pa.rotations.shape == (pa.P, 4) and pa.opacities.shape == (pa.P, 1)
    g = pa.activate()
    assert g["shs"].shape == (pa.P, pa.M, 3) and bool(torch.isfinite(g["scales"]).all())


@pytest.mark.parametrize("case", ["a", "c"])
def test_densify_and_prune_on_device(gold, case):
    """The whole call on the GPU (masks, nonzero, samples from the CUDA generator): everything but the sampled
    children must equal the reference; the children must be N = 2 samples around their parents with the parents'
    features and 1/1.6 of their scale."""
    from multiview_inpaint_b200.multiview import GradArena
    pa = arena_from_gold(gold, f"{case}_in")
    stats = GradArena(pa.P, pa.M, DEV)
    stats.grad_norm_accum.copy_(torch.from_numpy(gold[f"{case}_in_xyz_gradient_accum"].reshape(-1).copy()))
    stats.visible_count.copy_(torch.from_numpy(gold[f"{case}_in_denom"].reshape(-1).astype(np.int32)))
    max_grad, min_op, extent, mss, percent_dense, _ = gold[f"{case}_args"].tolist()
    plan = pa.densify_and_prune(stats, max_grad, min_op, extent, None if mss < 0 else int(mss), percent_dense=percent_dense)
    torch.cuda.synchronize()
    ref = cpu_plan(gold, case)
    assert plan.counts == ref.counts and plan.n_keep_state == ref.n_keep_state
    assert torch.equal(plan.src_row.cpu(), ref.src_row)
    child = torch.zeros(pa.P, dtype=torch.bool, device=DEV)
    child[plan.child_rows] = True
    want_xyz = torch.from_numpy(gold[f"{case}_out_xyz"].copy()).to(DEV)
    # children: same features / opacity / rotation as the reference's (copies of the parents), own positions
    pa_xyz = pa._xyz.clone()
    pa._xyz[child] = want_xyz[child]
    sc_gpu = pa._scaling[child].clone()
    pa._scaling[child] = torch.from_numpy(gold[f"{case}_out_scaling"].copy()).to(DEV)[child]
    assert_arena_equals_gold(pa, gold, f"{case}_out")
    assert torch.allclose(sc_gpu, torch.from_numpy(gold[f"{case}_out_scaling"].copy()).to(DEV)[child], atol=2e-6)
    # a child sits within a few standard deviations of its parent (|R s| = |s|, s ~ N(0, scale))
    parent_xyz = torch.from_numpy(gold[f"{case}_in_xyz"].copy()).to(DEV)[plan.src_row.long()[child]]
    parent_sc = torch.exp(torch.from_numpy(gold[f"{case}_in_scaling"].copy()).to(DEV))[plan.src_row.long()[child]]
    d = (pa_xyz[child] - parent_xyz).norm(dim=1) / parent_sc.norm(dim=1)
    assert float(d.max()) < 6.0 and float(d.mean()) > 0.2
    s2 = stats.resized(pa.P)
    assert s2.P == pa.P and float(s2.storage.abs().max()) == 0.0


def test_prune_points_equals_reference(gold):
    from multiview_inpaint_b200.multiview import GradArena
    pa = arena_from_gold(gold, "p_in")
    stats = GradArena(pa.P, pa.M, DEV)
    stats.grad_norm_accum.copy_(torch.from_numpy(gold["p_in_xyz_gradient_accum"].reshape(-1).copy()))
    stats.visible_count.copy_(torch.from_numpy(gold["p_in_denom"].reshape(-1).astype(np.int32)))
    mask = torch.from_numpy(gold["p_mask"].copy()).to(DEV)
    pa.prune_points(mask)
    assert_arena_equals_gold(pa, gold, "p_out")
    s2 = stats.pruned(mask)
    assert torch.equal(s2.grad_norm_accum.cpu(), torch.from_numpy(gold["p_out_xyz_gradient_accum"].reshape(-1).copy()))
    assert torch.equal(s2.visible_count.cpu(), torch.from_numpy(gold["p_out_denom"].reshape(-1).astype(np.int32)))


@pytest.mark.parametrize("case", ["a", "b"])
def test_reset_opacity_equals_reference(gold, case):
    pa = arena_from_gold(gold, f"{case}_out")
    pa.reset_opacity()
    want = torch.from_numpy(gold[f"{case}_reset_opacity"].copy()).to(DEV)
    assert float((pa._opacity - want).abs().max()) <= 2e-6 * float(want.abs().max())
    m, v = pa.moments("_opacity")
    assert float(m.abs().max()) == 0.0 and float(v.abs().max()) == 0.0
    m, v = pa.moments("_xyz")
    assert float(m.abs().max()) > 0.0            # the other groups keep their state


# ------------------------------------------------------------------------------------------------ the gather kernel
@pytest.mark.parametrize("n_src,n_dst", [(1, 1), (5, 63), (200, 64), (77, 65), (1000, 4097), (4096, 129), (100003, 250001)])
@pytest.mark.parametrize("M", [1, 4, 16])
@pytest.mark.parametrize("bulk", [1, 0])
def test_gather_rows_equals_torch_indexing(C, n_src, n_dst, M, bulk):
    """bulk = 1 (default): the SH segments (48- / 192-byte rows at M = 4 / 16) move by cp.async.bulk -- 64 row loads on one
    mbarrier, one 12 KB store per tile; bulk = 0: every segment through thread loads / stores.  Same bits either way."""
    C.debug_set(5, bulk)
    try:
        _gather_rows_case(C, n_src, n_dst, M)
    finally:
        C.debug_set(5, 1)


def _gather_rows_case(C, n_src, n_dst, M):
    g = torch.Generator(device=DEV).manual_seed(n_src * 31 + n_dst + M)
    widths = (3, 3 * M, 1, 3, 4)
    src = [torch.randn(n_src, w, device=DEV, generator=g) for w in widths]
    mom = [torch.randn(n_src, w, device=DEV, generator=g) for w in widths]
    idx = torch.randint(0, n_src, (n_dst,), device=DEV, generator=g, dtype=torch.int32)
    n_keep = n_dst * 2 // 3
    dst = [torch.full((n_dst, w), float("nan"), device=DEV) for w in widths]
    dmom = [torch.full((n_dst, w), float("nan"), device=DEV) for w in widths]
    segs = [dict(src=s, dst=d, zero_new=False) for s, d in zip(src, dst)] + \
           [dict(src=s, dst=d, zero_new=True) for s, d in zip(mom, dmom)]
    l0 = C.kernel_launches()
    C.gather_rows(idx, n_src, segs, n_keep_state=n_keep)
    assert C.kernel_launches() - l0 == 1           # ten segments, ONE launch
    torch.cuda.synchronize()
    for s, d in zip(src, dst):
        assert torch.equal(d, s[idx.long()])
    for s, d in zip(mom, dmom):
        w = s[idx.long()]
        w[n_keep:] = 0
        assert torch.equal(d, w)


def test_gather_rows_empty_and_errors(C):
    x = torch.randn(10, 3, device=DEV)
    idx = torch.zeros(0, dtype=torch.int32, device=DEV)
    C.gather_rows(idx, 10, [dict(src=x, dst=torch.empty(0, 3, device=DEV))])       # nothing to do
    idx = torch.arange(10, dtype=torch.int32, device=DEV)
    with pytest.raises(RuntimeError):
        C.gather_rows(idx, 10, [dict(src=x, dst=x)])                                # aliasing
    with pytest.raises(RuntimeError):
        C.gather_rows(idx, 10, [dict(src=x, dst=torch.empty(10, 4, device=DEV))])   # row length mismatch
    with pytest.raises(RuntimeError):
        C.gather_rows(idx.long(), 10, [dict(src=x, dst=torch.empty(10, 3, device=DEV))])   # int64 indices
    with pytest.raises(RuntimeError):
        C.gather_rows(idx.cpu(), 10, [dict(src=x, dst=torch.empty(10, 3, device=DEV))])    # no CPU path
    with pytest.raises(RuntimeError):
        C.gather_rows(idx, 10, [dict(src=x, dst=torch.empty(10, 3, device=DEV))] * 17)     # > GSR_GATHER_MAX_SEGS


def test_full_size_repack_checksum(C):
    """Headline model size (3 M Gaussians, M = 16, BASELINE.json): a permutation-with-drops must preserve per-row
    checksums (size-independent property), and a second densify with nothing selected is the identity."""
    from multiview_inpaint_b200 import densify
    from multiview_inpaint_b200.trainstep import GaussianParamArena
    P, M = 3_000_000, 16
    pa = GaussianParamArena(P, M, DEV)
    g = torch.Generator(device=DEV).manual_seed(3)
    pa.param.normal_(generator=g)
    pa.exp_avg.normal_(generator=g)
    pa.exp_avg_sq.uniform_(generator=g)
    # per-row checksum over the raw bit patterns (integer sums: exact and order-independent)
    bits = lambda a, flat, name: a._slice(flat, name).view(torch.int32).reshape(a.P, -1).long().sum(1)
    row_sum = lambda a, flat: sum(bits(a, flat, name) * (k + 1) for k, name in enumerate(SLICES))
    before = {k: row_sum(pa, getattr(pa, k)) for k in ("param", "exp_avg", "exp_avg_sq")}
    mask = torch.rand(P, device=DEV, generator=g) < 0.25
    plan = densify.plan_prune(mask)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    new = GaussianParamArena(plan.n_dst, M, DEV)
    segs = []
    for name in SLICES:
        segs.append(dict(src=getattr(pa, name), dst=getattr(new, name)))
        for a, b in zip(pa.moments(name), new.moments(name)):
            segs.append(dict(src=a, dst=b, zero_new=True))
    gb = plan.n_dst * (4 + 2 * 3 * 59 * 4) / 1e9
    for bulk in (0, 1, 0, 1):      # thread path, bulk (TMA) path -- twice: the first of each pays the cold caches
        C.debug_set(5, bulk)
        e0.record()
        C.gather_rows(plan.src_row, P, segs, n_keep_state=plan.n_keep_state)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"gather_rows (bulk={bulk}): {plan.n_dst} rows x 59 floats x 3 arenas in {ms:.3f} ms = {gb / ms * 1e3:.0f} GB/s")
    for k in before:
        assert torch.equal(row_sum(new, getattr(new, k)), before[k][~mask]), k
    assert ms < 5.0


def test_training_loop_through_a_densification():
    """train.py:86-128 with the fused step: optimise, densify_and_prune on the accumulated statistics, swap the
    statistics arena for one of the new size, keep optimising.  The model grows, nothing sized for the old P
    is touched, the loss stays finite and the surviving originals keep their Adam state."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    from multiview_inpaint_b200 import scenes as S
    from multiview_inpaint_b200.multiview import GradArena
    from multiview_inpaint_b200.trainstep import GaussianParamArena, ViewLoss, fused_train_step
    from tests.util import small_scene
    W, H, deg = 112, 80, 1
    sc = small_scene(P=2000, W=W, H=H, deg=deg, seed=77)
    raw = [sc["means3D"], sc["shs"][:, :1].contiguous(), sc["shs"][:, 1:].contiguous(),
           torch.logit(sc["opacities"].clamp(1e-4, 1 - 1e-4)).reshape(-1, 1), torch.log(sc["scales"]), sc["rotations"]]
    pa = GaussianParamArena.from_tensors(*(t.to(DEV) for t in raw))
    cams = [c.to(DEV) for c in S.orbit_cameras(2, W, H, max_deg=6.0)]
    bg = torch.zeros(3, device=DEV)
    settings = [GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                                              scale_modifier=1.0, viewmatrix=c.world_view_transform,
                                              projmatrix=c.full_proj_transform, sh_degree=deg, campos=c.camera_center,
                                              prefiltered=False) for c in cams]
    torch.manual_seed(3)
    losses = [ViewLoss(torch.rand(3, H, W, device=DEV), 0.2) for _ in cams]
    lrs = dict(xyz=0.00016, f_dc=0.0025, f_rest=0.0025 / 20, opacity=0.05, scaling=0.005, rotation=0.001)
    from multiview_inpaint_b200.densify import DensificationStats
    arena = GradArena(pa.P, pa.M, DEV)
    stats = DensificationStats(pa.P, DEV)
    hist = []
    for it in range(3):
        fused_train_step(pa, settings, losses, arena, lrs)
        stats.add_step(arena)                                   # train.py:115-116
        hist.append(sum(float(l.out3[2]) for l in losses))
    assert 0 < int(arena.visible_count.max()) <= len(cams)      # the arena holds ONE step's statistics ...
    assert int(stats.visible_count.max()) == 3 * int(arena.visible_count.max())   # ... the accumulator all three
    P0 = pa.P
    xyz_m_before = pa.moments("_xyz")[0].clone()
    # thresholds chosen so that the small test scene both clones and splits
    g = stats.grad_norm_accum / stats.visible_count.clamp(min=1)
    thr = float(g[stats.visible_count > 0].median())
    extent = float(torch.exp(pa._scaling).max(dim=1).values.median()) / 0.01
    plan = pa.densify_and_prune(stats, thr, 0.005, extent, None)
    assert plan.counts["cloned"] > 0 and plan.counts["split"] > 0
    assert pa.P == P0 + plan.counts["cloned"] + plan.counts["split"] - plan.counts["pruned"] and pa.P != P0
    kept = plan.src_row[:plan.n_keep_state].long()
    assert torch.equal(pa.moments("_xyz")[0][:plan.n_keep_state], xyz_m_before[kept])
    assert float(pa.moments("_xyz")[0][plan.n_keep_state:].abs().max()) == 0.0
    arena, stats = arena.resized(pa.P), stats.resized(pa.P)
    for it in range(2):
        fused_train_step(pa, settings, losses, arena, lrs)
        stats.add_step(arena)
        hist.append(sum(float(l.out3[2]) for l in losses))
    torch.cuda.synchronize()
    assert all(np.isfinite(h) for h in hist), hist
    assert pa.step_count == 5 and bool(torch.isfinite(pa.param).all())
    assert 0 < int(stats.visible_count.max()) <= 2 * len(cams)  # statistics restarted after the densification
    pa.reset_opacity()
    fused_train_step(pa, settings, losses, arena, lrs)
    assert bool(torch.isfinite(pa.param).all()) and float(torch.sigmoid(pa._opacity).max()) < 0.2
