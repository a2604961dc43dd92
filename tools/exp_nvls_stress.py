"""TWO (or more) ranks, torchrun: stress of GradArena.all_reduce() under rank skew.  Every iteration each rank fills its arena with a
pattern both ranks can reproduce, one rank is delayed by dummy work, the arena is reduced, and the result is compared ON THE
DEVICE with the sum formed locally (two operands: the in-switch add and torch's are the same IEEE add).  No host sync inside
the loop.  Prints the number of mismatching elements per segment over all iterations."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from multiview_inpaint_b200 import multiview as mv  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
method = sys.argv[2] if len(sys.argv) > 2 else "nvls"
for P, M in ((4001, 4), (70001, 16), (300000, 16)):
    arena = mv.GradArena(P, M, dev, symmetric=True)
    arena.method = method if arena._mc else "nccl"
    n = arena.storage.numel()
    bad = torch.zeros(4, dtype=torch.int64, device=dev)
    junk = torch.randn(2048, 2048, device=dev)
    gens = [torch.Generator(device=dev) for _ in range(world)]
    for it in range(iters):
        fl, cn, mr = [], [], []
        for r in range(world):
            gens[r].manual_seed(1000 * it + r)
            fl.append(torch.randn(arena._n_f32, device=dev, generator=gens[r]))
            cn.append(torch.randint(0, 2, (P,), device=dev, generator=gens[r], dtype=torch.int32))
            mr.append(torch.randint(0, 900, (P,), device=dev, generator=gens[r], dtype=torch.int32))
            o, w = arena._offs["dL_dsh"], 3 * M
            fl[r][o:o + P * w].view(P, w)[cn[r] == 0] = 0
        tot = sum(cn)
        if it % 3 == rank % 3:                     # skew: this rank arrives late at the exchange
            for _ in range(1 + it % 4):
                junk = junk @ junk * 1e-3
        arena.storage[:arena._n_f32].copy_(fl[rank])
        arena.visible_count.copy_(cn[rank])
        arena.max_radii.copy_(mr[rank])
        arena.all_reduce()
        exp = fl[0]
        for r in range(1, world):
            exp = exp + fl[r]
        got = arena.storage[:arena._n_f32]
        if world == 2:
            bad[0] += (got != exp).sum()
        else:
            bad[0] += ((got - exp).abs() > 1e-5 * (1 + exp.abs())).sum()
        bad[1] += (arena.visible_count != tot).sum()
        mx = mr[0]
        for r in range(1, world):
            mx = torch.maximum(mx, mr[r])
        bad[2] += (arena.max_radii != mx).sum()
        arena.storage[:arena._n_f32].mul_(0.5)     # the chain rule rewrites the arena in place right after the exchange
    torch.cuda.synchronize()
    allbad = [torch.zeros_like(bad) for _ in range(world)]
    dist.all_gather(allbad, bad)
    if rank == 0:
        print(f"P={P} M={M} method={arena.method} uses_nvls={arena.uses_nvls} iters={iters} mismatches per rank [f32, count, max]:",
              [b[:3].tolist() for b in allbad], flush=True)
dist.destroy_process_group()
