"""Host logic of the pipelined in-switch all-reduce (multiview.GradArena.all_reduce_range / all_reduce and the
Gaussian-range chunking of cuda_views_geom_backward_allreduce), checked on the CPU: the plans handed to
gsr_nvls_all_reduce_plan must tile the arena -- every float of every gradient slice and every statistic of every
Gaussian reduced exactly once over the chunks of a step, nothing outside the allocation, every range 16-byte aligned
(multimem.ld_reduce / multimem.st work on float4).  The CUDA entry points are replaced by recorders; the kernel itself
is compared with NCCL on 2 GPUs in tests/test_nvls_gpu.py."""
import types

import numpy as np
import pytest
import torch

from multiview_inpaint_b200 import _C
from multiview_inpaint_b200 import multiview as mv


def _fake_symmetric(arena, world=8, rank=3):
    """what symm_mem.rendezvous would give on a multi-GPU box: a multicast pointer and a handle with barrier()"""
    calls = {"barrier": 0}

    def barrier():
        calls["barrier"] += 1
    arena._mc = 0x7f0000000000
    arena.method = "nvls"
    arena._handle = types.SimpleNamespace(barrier=barrier, rank=rank, world_size=world)
    return calls


def _chunks(P, chunks, taper=False):
    """the Gaussian ranges cuda_views_geom_backward_allreduce walks (multiview.py)"""
    return mv.chunk_ranges(P, chunks, taper)


def test_chunk_ranges_tapered():
    for P, c in ((3_000_000, 5), (4097, 5), (100003, 6), (64, 5), (33, 4), (1, 5)):
        r = mv.chunk_ranges(P, c, taper=True)
        assert r[0][0] == 0 and r[-1][1] == P and all(a[1] == b[0] and a[0] < a[1] for a, b in zip(r, r[1:] + [(P, P + 1)]))
        assert all(g0 % 32 == 0 for g0, _ in r)
    r = mv.chunk_ranges(3_000_000, 5, taper=True)
    lens = [b - a for a, b in r]
    assert len(r) == 5 and abs(lens[0] - 375_000) <= 32 and abs(lens[1] - 750_000) <= 32 and abs(lens[-1] - 375_000) <= 64
    assert mv.chunk_ranges(3_000_000, 4) == [(k * 750_016, min(3_000_000, (k + 1) * 750_016)) for k in range(4)]


@pytest.mark.parametrize("taper", [False, True])
@pytest.mark.parametrize("P,M,chunks", [(4096, 16, 8), (4097, 16, 8), (100003, 4, 8), (5000, 1, 4), (33, 16, 2),
                                        (3_000_000, 16, 8), (1_000_001, 4, 5)])
def test_range_plans_tile_the_arena(monkeypatch, P, M, chunks, taper):
    arena = mv.GradArena(P, M, "cpu")
    calls = _fake_symmetric(arena)
    plans = []
    monkeypatch.setattr(_C, "nvls_all_reduce_plan",
                        lambda mc, dev, rank, world, dense=(), rows=None, add_s32=(0, 0), max_s32=(0, 0), blocks=0:
                        plans.append(dict(mc=mc, rank=rank, world=world, dense=list(dense), rows=rows, add=add_s32, mx=max_s32)))
    ranges = _chunks(P, chunks, taper)
    assert ranges[0][0] == 0 and ranges[-1][1] == P and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    for g0, g1 in ranges:
        arena.all_reduce_range(g0, g1, post_barrier=False)
    assert len(plans) == len(ranges) and calls["barrier"] == len(ranges)       # one barrier in front of every launch
    n_total = arena.storage.numel()
    cover = np.zeros(n_total, dtype=np.int16)            # how often each 4-byte word of the allocation is reduced
    rows_are_sparse = (3 * M) % 4 == 0 and 3 * M in (12, 48)
    for pl in plans:
        assert pl["mc"] == arena._mc and pl["rank"] == 3 and pl["world"] == 8
        assert len(pl["dense"]) <= 6
        for off, n in pl["dense"]:
            assert off % 16 == 0 and n % 4 == 0 and n > 0, (off, n)
            cover[off // 4: off // 4 + n] += 1
        if pl["rows"] is not None:
            off, n_rows, w, cnt_off = pl["rows"]
            assert rows_are_sparse and off % 16 == 0 and w % 4 == 0 and w == 3 * M and cnt_off % 4 == 0
            cover[off // 4: off // 4 + n_rows * w] += 1
            # the row counts the kernel consults are the visible_count entries of the same Gaussians
            g0 = (cnt_off - arena._off_cnt) // 4
            assert 0 <= g0 and g0 + n_rows <= P and off == 4 * (arena._offs["dL_dsh"] + g0 * w)
        else:
            assert not rows_are_sparse
        for off, n in (pl["add"], pl["mx"]):
            assert off % 4 == 0
            cover[off // 4: off // 4 + n] += 1
    assert cover.max() == 1, "a word is reduced twice (it would be summed R times too often)"
    # every word that carries data is covered; only slice padding (< 4 words per slice) may be left out or touched
    names = ("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations")
    data = np.zeros(n_total, dtype=bool)
    for name in names:
        o, w = arena._offs[name], arena._row_w[name]
        data[o: o + P * w] = True
    for base in (arena._n_flat, arena._n_flat + arena._Pp, arena._n_flat + 2 * arena._Pp):
        data[base: base + P] = True
    assert (cover[data] == 1).all(), "a gradient or statistic word is never reduced"
    assert (~data).sum() <= 4 * 8 and cover[~data].sum() <= (~data).sum()


@pytest.mark.parametrize("P,M", [(1000, 16), (1001, 4), (77, 1)])
def test_whole_arena_call_matches_layout(monkeypatch, P, M):
    arena = mv.GradArena(P, M, "cpu")
    calls = _fake_symmetric(arena, world=4, rank=1)
    seen = []
    monkeypatch.setattr(_C, "nvls_all_reduce", lambda *a: seen.append(a))
    monkeypatch.setattr(torch.distributed, "is_initialized", lambda: True)
    monkeypatch.setattr(torch.distributed, "get_world_size", lambda group=None: 4)
    arena.all_reduce()
    assert calls["barrier"] == 2 and len(seen) == 1              # barrier, kernel, barrier
    mc, dev, off_f32, n_f32, off_cnt, n_cnt, off_max, n_max, rank, world, blocks, sh_first, rows, row_f32 = seen[0]
    assert (mc, rank, world, off_f32) == (arena._mc, 1, 4, 0)
    assert n_f32 == arena._n_flat + arena._Pp and n_f32 % 4 == 0      # gradients + grad_norm_accum in one float range
    assert off_cnt == 4 * n_f32 and off_max == off_cnt + 4 * arena._Pp and n_cnt == n_max == P
    assert sh_first == arena._offs["dL_dsh"] and row_f32 == 3 * M
    assert rows == (P if (3 * M) % 4 == 0 else 0)                # M = 1: rows of 3 floats cannot be float4-skipped
    # the typed views the kernels write are windows of exactly these ranges
    assert arena.visible_count.data_ptr() - arena.storage.data_ptr() == off_cnt
    assert arena.max_radii.data_ptr() - arena.storage.data_ptr() == off_max
    assert arena.views["dL_dsh"].data_ptr() - arena.storage.data_ptr() == 4 * sh_first


def test_range_requires_alignment_and_nvls():
    arena = mv.GradArena(1000, 16, "cpu")
    with pytest.raises(AssertionError):
        arena.all_reduce_range(0, 100)                              # no multicast mapping: NCCL path only
    _fake_symmetric(arena)
    with pytest.raises(AssertionError):
        arena.all_reduce_range(2, 100)                              # g0 must be a multiple of 4
    with pytest.raises(AssertionError):
        arena.all_reduce_range(0, 1001)


def test_two_rank_partition_covers_every_element_once():
    """the split of csrc/nvls_allreduce.cu::p2p_allreduce2_kernel (rank r owns [min(n, per r), min(n, per r + per)) with
    per = ceil(n / 2) of every dense range, of the row groups and of the two integer ranges), restated: the two ranks'
    shares are disjoint and together cover everything, for odd and tiny sizes too"""
    def share(n, rank):
        per = (n + 1) // 2
        lo = min(n, per * rank)
        return lo, min(n, lo + per)
    for n in (0, 1, 2, 3, 4, 5, 95, 96, 97, 4001, 70001, 3_000_000):
        a, b = share(n, 0), share(n, 1)
        assert a[0] == 0 and a[1] == b[0] and b[1] == n, (n, a, b)
    for rows, row_f4 in ((4001, 3), (70001, 12), (5, 12), (0, 3), (33, 3)):
        G = 96 // row_f4
        groups = (rows + G - 1) // G
        (g0, g1), (h0, h1) = share(groups, 0), share(groups, 1)
        covered = np.zeros(rows, np.int16)
        for lo, hi in ((g0, g1), (h0, h1)):
            for g in range(lo, hi):
                covered[g * G:min(rows, (g + 1) * G)] += 1
        assert (covered == 1).all(), (rows, row_f4)
