"""TWO ranks: tests/test_trainstep_world2_gpu.py step by step.  After each fused step rank 0 compares its parameters with a
single-rank run of all four views made in the same process, per parameter group."""
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
from multiview_inpaint_b200.trainstep import ViewLoss, fused_train_step  # noqa: E402
from tests.test_trainstep_world2_gpu import _setup  # noqa: E402

n_views = 4
for method in ("nvls", "nccl"):
    pa, settings, gts, lrs, M = _setup(dev)
    pb, *_ = _setup(dev)
    mine = mv.shard_views(n_views, rank, world)
    arena = mv.GradArena(pa.P, M, dev, symmetric=True)
    arena.method = method if arena._mc else "nccl"
    ref = mv.GradArena(pb.P, M, dev)
    losses = [ViewLoss(gts[v], 0.2, weight=1.0 / n_views) for v in mine]
    losses_all = [ViewLoss(gt, 0.2, weight=1.0 / n_views) for gt in gts[:n_views]]
    for step in range(3):
        fused_train_step(pa, [settings[v] for v in mine], losses, arena, lrs, all_reduce=True)
        fused_train_step(pb, settings[:n_views], losses_all, ref, lrs)
        torch.cuda.synchronize()
        if rank == 0:
            out = {}
            for name in ("_xyz", "_features", "_opacity", "_scaling", "_rotation"):
                d = (getattr(pa, name) - getattr(pb, name)).abs()
                out[name] = (round(float((d > 1e-6).float().mean()), 5), float(d.max()))
            gd = {k: (int(((arena.views[k] - ref.views[k]).abs() > 1e-9).sum()), float((arena.views[k] - ref.views[k]).abs().max())) for k in ref.views}
            print(method, "step", step, "params", out, flush=True)
            print("      arena after the step (chain-ruled grads)", gd, flush=True)
        dist.barrier()
dist.destroy_process_group()
