"""Host logic of the densification re-pack (multiview_inpaint_b200/densify.py) against golden vectors produced by the
reference's own GaussianModel (tests/golden/make_densify_golden.py runs gs-simp/scene/gaussian_model.py:467-480,
:365-383, :263-266 unmodified): the composed index map must reproduce, bit for bit, the six parameter tensors and
both Adam moments the reference ends up with after clone -> split -> prune parents -> final prune.  The plan is
applied here with plain torch indexing (the checker); the CUDA gather that applies it in the product is compared
with the same vectors in tests/test_densify_gpu.py."""
import os

import numpy as np
import pytest
import torch

from multiview_inpaint_b200 import densify

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "densify.npz"))
GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
CASES = ("a", "b", "c", "d")


def gold_model(prefix):
    t = lambda k: torch.from_numpy(GOLD[f"{prefix}_{k}"].copy())
    m = {g: t(g) for g in GROUPS}
    m.update({f"{g}_exp_avg": t(f"{g}_exp_avg") for g in GROUPS})
    m.update({f"{g}_exp_avg_sq": t(f"{g}_exp_avg_sq") for g in GROUPS})
    for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
        m[k] = t(k)
    return m


def make_plan(case):
    m = gold_model(f"{case}_in")
    max_grad, min_op, extent, mss, percent_dense, nseed = GOLD[f"{case}_args"].tolist()
    torch.manual_seed(int(nseed))
    plan = densify.plan_densify_and_prune(m["xyz"], m["scaling"], m["rotation"], m["opacity"], m["xyz_gradient_accum"],
                                          m["denom"], max_grad, min_op, extent, None if mss < 0 else int(mss),
                                          percent_dense=percent_dense)
    return m, plan


def apply_with_torch(m, plan):
    """the checker: what gsr_gather_rows + the child overwrite must produce"""
    idx = plan.src_row.long()
    out = {}
    for g in GROUPS:
        out[g] = m[g][idx].clone()
        for mom in ("exp_avg", "exp_avg_sq"):
            v = m[f"{g}_{mom}"][idx].clone()
            v[plan.n_keep_state:] = 0
            out[f"{g}_{mom}"] = v
    out["xyz"][plan.child_rows] = plan.child_xyz
    out["scaling"][plan.child_rows] = plan.child_scaling
    return out


@pytest.mark.parametrize("case", CASES)
def test_plan_reproduces_reference_densify_and_prune(case):
    m, plan = make_plan(case)
    want = gold_model(f"{case}_out")
    got = apply_with_torch(m, plan)
    assert plan.n_dst == want["xyz"].shape[0]
    for g in GROUPS:
        for k in (g, f"{g}_exp_avg", f"{g}_exp_avg_sq"):
            assert got[k].shape == want[k].shape, k
            assert torch.equal(got[k], want[k]), f"{case}:{k} differs from the reference"
    # densification_postfix leaves zeroed statistics of the new size (gaussian_model.py:423-425)
    for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
        assert want[k].shape[0] == plan.n_dst and float(want[k].abs().max()) == 0.0


@pytest.mark.parametrize("case", CASES)
def test_plan_structure(case):
    m, plan = make_plan(case)
    P = m["xyz"].shape[0]
    s = plan.src_row.long()
    assert plan.src_row.dtype == torch.int32 and plan.n_src == P
    assert int(s.min()) >= 0 and int(s.max()) < P
    # surviving originals first, strictly ascending, each at most once; they alone keep optimizer state
    head = s[:plan.n_keep_state]
    assert bool((head[1:] > head[:-1]).all())
    # children are the tail of the model and come in N = 2 copies of the parents' order
    if plan.child_rows.numel():
        assert int(plan.child_rows.min()) >= plan.n_keep_state
        assert bool((plan.child_rows[1:] > plan.child_rows[:-1]).all())
        assert int(plan.child_rows.max()) == plan.n_dst - 1
    c = plan.counts
    assert plan.n_dst == P + c["cloned"] + c["split"] * 2 - c["split"] - c["pruned"]
    if case == "d":
        assert c == dict(cloned=0, split=0, pruned=0) and torch.equal(s, torch.arange(P))
    else:
        assert c["cloned"] > 0 and c["split"] > 0 and c["pruned"] > 0


def test_prune_points_matches_reference():
    m, want = gold_model("p_in"), gold_model("p_out")
    mask = torch.from_numpy(GOLD["p_mask"].copy())
    plan = densify.plan_prune(mask)
    assert plan.n_keep_state == plan.n_dst == want["xyz"].shape[0] and plan.child_rows.numel() == 0
    got = apply_with_torch(m, plan)
    for g in GROUPS:
        for k in (g, f"{g}_exp_avg", f"{g}_exp_avg_sq"):
            assert torch.equal(got[k], want[k]), k
    # statistics are carried over by prune_points (gaussian_model.py:379-383)
    for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
        assert torch.equal(m[k][~mask], want[k])


@pytest.mark.parametrize("case", CASES)
def test_reset_opacity_matches_reference(case):
    out = gold_model(f"{case}_out")
    got = densify.reset_opacity_values(out["opacity"])
    assert torch.equal(got, torch.from_numpy(GOLD[f"{case}_reset_opacity"].copy()))


def test_build_rotation_is_orthonormal():
    torch.manual_seed(0)
    R = densify.build_rotation(torch.randn(50, 4))
    eye = torch.eye(3).expand(50, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), eye, atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(50), atol=1e-5)


def test_grad_arena_resized_and_pruned_cpu():
    from multiview_inpaint_b200.multiview import GradArena
    a = GradArena(10, 4, "cpu")
    a.grad_norm_accum.copy_(torch.arange(10.0))
    a.visible_count.copy_(torch.arange(10, dtype=torch.int32))
    a.max_radii.copy_(torch.arange(10, dtype=torch.int32) * 3)
    mask = torch.tensor([0, 1, 0, 0, 1, 0, 0, 0, 0, 1], dtype=torch.bool)
    b = a.pruned(mask)
    assert b.P == 7 and torch.equal(b.grad_norm_accum, torch.arange(10.0)[~mask])
    assert torch.equal(b.visible_count, torch.arange(10, dtype=torch.int32)[~mask])
    assert torch.equal(b.max_radii, (torch.arange(10, dtype=torch.int32) * 3)[~mask])
    c = a.resized(25)
    assert c.P == 25 and c.M == 4 and float(c.storage.abs().max()) == 0.0
    assert c.views["dL_dsh"].shape == (25, 4, 3)


def test_densification_stats_accumulate_like_train_py():
    """train.py:115-116 over several steps: SUM of per-step gradient norms, SUM of visibility counts, MAX of radii."""
    from multiview_inpaint_b200.multiview import GradArena
    torch.manual_seed(1)
    P = 50
    st = densify.DensificationStats(P, "cpu")
    arena = GradArena(P, 1, "cpu")
    want_g, want_c, want_r = torch.zeros(P), torch.zeros(P, dtype=torch.int32), torch.zeros(P, dtype=torch.int32)
    for step in range(4):
        arena.grad_norm_accum.copy_(torch.rand(P))
        arena.visible_count.copy_(torch.randint(0, 3, (P,), dtype=torch.int32))
        arena.max_radii.copy_(torch.randint(0, 40, (P,), dtype=torch.int32))
        st.add_step(arena)
        want_g += arena.grad_norm_accum
        want_c += arena.visible_count
        want_r = torch.maximum(want_r, arena.max_radii)
    assert torch.equal(st.grad_norm_accum, want_g) and torch.equal(st.visible_count, want_c)
    assert torch.equal(st.max_radii, want_r)
    mask = torch.rand(P) < 0.4
    p = st.pruned(mask)
    assert p.P == int((~mask).sum()) and torch.equal(p.visible_count, want_c[~mask])
    z = st.resized(70)
    assert z.P == 70 and int(z.visible_count.abs().max()) == 0 and float(z.grad_norm_accum.abs().max()) == 0.0
    # the accumulator feeds the plan exactly like the reference's two tensors
    m, _ = make_plan("a")
    st2 = densify.DensificationStats(m["xyz"].shape[0], "cpu")
    st2.grad_norm_accum.copy_(m["xyz_gradient_accum"].reshape(-1))
    st2.visible_count.copy_(m["denom"].reshape(-1).to(torch.int32))
    max_grad, min_op, extent, mss, percent_dense, nseed = GOLD["a_args"].tolist()
    torch.manual_seed(int(nseed))
    plan = densify.plan_densify_and_prune(m["xyz"], m["scaling"], m["rotation"], m["opacity"], st2.grad_norm_accum,
                                          st2.visible_count, max_grad, min_op, extent, int(mss), percent_dense=percent_dense)
    torch.manual_seed(int(nseed))
    _, ref = make_plan("a")
    assert torch.equal(plan.src_row, ref.src_row) and torch.equal(plan.child_xyz, ref.child_xyz)


# ---------------------------------------------------------------- world 2 (gloo): replicas densify identically
def _densify_rank(rank, world, port, out_dir):
    import torch.distributed as dist
    from multiview_inpaint_b200.multiview import GradArena
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = gold_model("a_in")
    P = m["xyz"].shape[0]
    max_grad, min_op, extent, mss, percent_dense, _ = GOLD["a_args"].tolist()
    # each rank saw its own views: the statistics are partial until the step's all-reduce (SUM / SUM / MAX)
    arena = GradArena(P, 4, "cpu")
    share = (torch.arange(P) % world == rank)
    arena.grad_norm_accum.copy_(m["xyz_gradient_accum"].reshape(-1) * share)
    arena.visible_count.copy_((m["denom"].reshape(-1) * share).to(torch.int32))
    arena.all_reduce()
    stats = densify.DensificationStats(P, "cpu")
    stats.add_step(arena)
    torch.manual_seed(1000 + 17 * rank)                 # replicas whose default generators have drifted apart
    unsynced = torch.rand(1).item()
    synced = densify.sync_rng("cpu")
    plan = densify.plan_densify_and_prune(m["xyz"], m["scaling"], m["rotation"], m["opacity"], stats.grad_norm_accum,
                                          stats.visible_count, max_grad, min_op, extent, int(mss), percent_dense=percent_dense)
    torch.save(dict(src_row=plan.src_row, child_xyz=plan.child_xyz, child_scaling=plan.child_scaling, n_keep=plan.n_keep_state,
                    counts=plan.counts, synced=synced, unsynced=unsynced), os.path.join(out_dir, f"plan{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_replicas_densify_identically(tmp_path):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_densify_rank, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "plan0.pt"), torch.load(tmp_path / "plan1.pt")
    assert a["synced"] and b["synced"] and a["unsynced"] != b["unsynced"]
    assert a["counts"] == b["counts"] and a["n_keep"] == b["n_keep"] and a["counts"]["split"] > 0
    for k in ("src_row", "child_xyz", "child_scaling"):
        assert torch.equal(a[k], b[k]), k
    # and the all-reduced statistics gave the reference's masks: same rows as the single-process golden plan
    _, ref = make_plan("a")
    assert torch.equal(a["src_row"], ref.src_row)


def test_sync_rng_is_a_no_op_without_a_process_group():
    assert densify.sync_rng("cpu") is False
