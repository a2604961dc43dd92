"""World-2 (GSR_TEST_WORLD=N: world-N) check of the in-switch arena all-reduce (gsr_nvls_all_reduce over symmetric memory with an
NVSwitch multicast mapping) against NCCL.  Needs two GPUs: skipped on the single-GPU test box."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from multiview_inpaint_b200 import multiview as mv
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {}
    for M in (16, 4, 1):
        P = 70001
        a = mv.GradArena(P, M, dev, symmetric=True)
        b = mv.GradArena(P, M, dev, symmetric=False)
        g = torch.Generator(device=dev)
        g.manual_seed(5 + rank)
        for x in (a, b):
            g.manual_seed(5 + rank)
            x.flat.copy_(torch.randn(x.flat.shape, device=dev, generator=g))
            x.grad_norm_accum.copy_(torch.rand(P, device=dev, generator=g))
            x.visible_count.copy_(torch.randint(0, 3, (P,), device=dev, generator=g, dtype=torch.int32))
            x.max_radii.copy_(torch.randint(0, 900, (P,), device=dev, generator=g, dtype=torch.int32))
            x.views["dL_dsh"][x.visible_count == 0] = 0     # unseen Gaussians have zero rows (what the kernels write)
        a.method = "nvls"
        a.nvls_blocks = 148 if M == 4 else 0      # the leaner grid the pipelined step uses, and the library default
        a.all_reduce()
        b.all_reduce()
        torch.cuda.synchronize()
        # the same reduction in Gaussian-range chunks (gsr_nvls_all_reduce_plan), as the pipelined backward issues it
        c = mv.GradArena(P, M, dev, symmetric=True)
        g.manual_seed(5 + rank)
        c.flat.copy_(torch.randn(c.flat.shape, device=dev, generator=g))
        c.grad_norm_accum.copy_(torch.rand(P, device=dev, generator=g))
        c.visible_count.copy_(torch.randint(0, 3, (P,), device=dev, generator=g, dtype=torch.int32))
        c.max_radii.copy_(torch.randint(0, 900, (P,), device=dev, generator=g, dtype=torch.int32))
        c.views["dL_dsh"][c.visible_count == 0] = 0
        chunk_ok = True
        if c.uses_nvls:
            for g0 in range(0, P, 23456):
                c.all_reduce_range(g0, min(P, g0 + 23456))
            torch.cuda.synchronize()
            chunk_ok = bool((c.flat - b.flat).abs().max() <= 1e-5 * b.flat.abs().max()
                            and (c.grad_norm_accum - b.grad_norm_accum).abs().max() <= 1e-5
                            and torch.equal(c.visible_count, b.visible_count) and torch.equal(c.max_radii, b.max_radii))
        # two ranks: the peer-to-peer two-shot kernel, whole arena and in Gaussian ranges
        p2p_ok = True
        if world == 2 and a._peer_ptr():
            for ranges in ([(0, P)], [(g0, min(P, g0 + 23456)) for g0 in range(0, P, 23456)]):
                d = mv.GradArena(P, M, dev, symmetric=True)
                d.method = "p2p"
                g.manual_seed(5 + rank)
                d.flat.copy_(torch.randn(d.flat.shape, device=dev, generator=g))
                d.grad_norm_accum.copy_(torch.rand(P, device=dev, generator=g))
                d.visible_count.copy_(torch.randint(0, 3, (P,), device=dev, generator=g, dtype=torch.int32))
                d.max_radii.copy_(torch.randint(0, 900, (P,), device=dev, generator=g, dtype=torch.int32))
                d.views["dL_dsh"][d.visible_count == 0] = 0
                assert d.uses_nvls
                if len(ranges) == 1:
                    d.all_reduce()
                else:
                    for g0, g1 in ranges:
                        d.all_reduce_range(g0, g1)
                torch.cuda.synchronize()
                p2p_ok = p2p_ok and bool(torch.equal(d.flat, b.flat) and torch.equal(d.grad_norm_accum, b.grad_norm_accum)
                                         and torch.equal(d.visible_count, b.visible_count) and torch.equal(d.max_radii, b.max_radii))
        res[M] = dict(nvls=a.uses_nvls, chunks=chunk_ok, p2p=p2p_ok, err=float((a.flat - b.flat).abs().max()), scale=float(b.flat.abs().max()),
                      norm=float((a.grad_norm_accum - b.grad_norm_accum).abs().max()),
                      ints=bool(torch.equal(a.visible_count, b.visible_count) and torch.equal(a.max_radii, b.max_radii)))
    torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_nvls_arena_all_reduce_matches_nccl(tmp_path):
    world = int(os.environ.get("GSR_TEST_WORLD", "2"))     # 2 by default; the 8-GPU calls of tools/ run it with 8
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, 29541, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        for M, d in res.items():
            if not d["nvls"]:
                pytest.skip("no multicast mapping on this system")
            assert d["err"] <= 1e-5 * d["scale"] and d["norm"] <= 1e-5 and d["ints"] and d["chunks"] and d["p2p"], (M, d)
