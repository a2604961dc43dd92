"""View-sharded step on CPU with the gloo backend, world_size 2.
The per-view fwd+bwd is the CPU oracle here (tests may use it; the product path is CUDA): what is
under test is the host logic -- view -> rank assignment, in-place accumulation into the flat
gradient arena, and the SUM / SUM / MAX reductions with the reference's densification semantics
(scene/gaussian_model.py:482-484, train.py:115)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multiview_inpaint_b200 import multiview as mv
from multiview_inpaint_b200 import scenes as S

N_VIEWS = 5
P, W, H, DEG = 400, 48, 32, 1


def _scene():
    return S.make_scene(P, W, H, DEG, 77, mu_s=S.default_mu_s(W, 8.0))


def _oracle_view(sc, cam, wt, arena):
    from oracle import oracle as O
    from tests.util import oracle_forward
    f = oracle_forward(O, sc, cam=cam)
    g = O.backward(f, wt.numpy())
    for name in ("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations"):
        arena.views[name] += torch.from_numpy(g[name]).view_as(arena.views[name])
    # the densification statistics with the reference's per-view semantics, accumulated by the CHECKER here (on the
    # device the product does this in gsr_accumulate_view_stats, tests/test_parity_gpu.py)
    radii, vis = torch.from_numpy(f.radii), torch.from_numpy(f.radii > 0)
    arena.grad_norm_accum += torch.norm(torch.from_numpy(g["dL_dmeans2D"])[:, :2], dim=-1) * vis   # gaussian_model.py:483
    arena.visible_count += vis.to(torch.int32)                                                    # gaussian_model.py:484
    torch.maximum(arena.max_radii, radii, out=arena.max_radii)                                    # train.py:115


def _run_rank(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sc = _scene()
    cams = S.orbit_cameras(N_VIEWS, W, H, max_deg=15.0)
    arena = mv.GradArena(P, sc["M"], "cpu")
    mine = mv.sharded_step(lambda v: _oracle_view(sc, cams[v], S.loss_weights(W, H, v), arena), N_VIEWS, arena)
    assert mine == list(range(rank, N_VIEWS, world))
    torch.save({"flat": arena.flat.clone(), "norm": arena.grad_norm_accum.clone(), "vis": arena.visible_count.clone(),
                "maxr": arena.max_radii.clone(), "mine": mine}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_views_round_robin():
    assert mv.shard_views(25, 0, 8) == [0, 8, 16, 24]
    assert mv.shard_views(25, 7, 8) == [7, 15, 23]
    assert sorted(sum((mv.shard_views(25, r, 8) for r in range(8)), [])) == list(range(25))
    assert mv.shard_views(3, 5, 8) == []


def test_arena_layout_is_flat_aligned_and_sized():
    for M, per in ((1, 14), (4, 23), (16, 59)):
        a = mv.GradArena(1000, M, "cpu")
        assert a.flat.numel() >= 1000 * per and a.flat.numel() <= 1000 * per + 16
        assert all(v.data_ptr() % 16 == 0 for v in a.views.values())
        assert a.views["dL_dsh"].shape == (1000, M, 3)
        a.views["dL_drotations"].fill_(1.0)
        assert a.flat.sum() == 4000
        a.zero_()
        assert a.flat.abs().sum() == 0


def test_world2_gloo_sharded_step_equals_single_rank_loop(tmp_path, oracle):
    world = 2
    port = _free_port()
    mp.spawn(_run_rank, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert r0["mine"] == [0, 2, 4] and r1["mine"] == [1, 3]
    # after the all-reduce both ranks hold the same sums
    for k in ("flat", "norm", "vis", "maxr"):
        assert torch.equal(r0[k], r1[k]), k
    # single-rank reference: loop over all views
    sc = _scene()
    cams = S.orbit_cameras(N_VIEWS, W, H, max_deg=15.0)
    ref = mv.GradArena(P, sc["M"], "cpu")
    for v in range(N_VIEWS):
        _oracle_view(sc, cams[v], S.loss_weights(W, H, v), ref)
    scale = ref.flat.abs().max()
    assert (r0["flat"] - ref.flat).abs().max() <= 1e-5 * scale      # fp32 sum order differs across ranks
    assert (r0["norm"] - ref.grad_norm_accum).abs().max() <= 1e-5 * ref.grad_norm_accum.max()
    assert torch.equal(r0["vis"], ref.visible_count) and torch.equal(r0["maxr"], ref.max_radii)
    assert int(ref.visible_count.max()) == N_VIEWS and ref.flat.abs().sum() > 0
    # per-view norm semantics: sum of norms, NOT norm of the summed gradient
    assert (ref.grad_norm_accum >= 0).all()


def test_async_views_bookkeeping():
    a = mv.AsyncViews(3)
    a.learn(0, 1000)
    # one capacity for every view (largest seen, rounded up to 2^20): buffers are interchangeable between views
    assert a.capacity(0) == 1 << 20 == a.capacity(1) and mv.AsyncViews(2).capacity(0) == 0
    a.slots[0, 0], a.slots[0, 1] = 900, 0
    a.slots[1, 0], a.slots[1, 1] = 5_000_000, 1 << 32          # overflow bit set by the kernels
    assert a.check([0, 1]) == [1]
    assert a.capacity(1) >= 5_000_000
    a.slots[2, 1] = 1                                           # prefiltered trap
    with pytest.raises(RuntimeError):
        a.check([2])


def test_view_stats_have_no_cpu_path():
    a = mv.GradArena(4, 1, "cpu")
    with pytest.raises(RuntimeError):
        a.add_view_stats(torch.zeros(4, 3), torch.ones(4, dtype=torch.int32))
