"""TWO ranks (torchrun --nproc-per-node 2): the gradient arena of tests/test_trainstep_world2_gpu.py's first step, reduced
by the in-switch kernel and by NCCL, against the sum both ranks can form locally (each rank renders BOTH ranks' views into
plain arenas).  Prints per segment the number of elements that differ and the largest difference."""
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
from multiview_inpaint_b200.trainstep import ViewLoss  # noqa: E402
from tests.test_trainstep_world2_gpu import _setup  # noqa: E402

pa, settings, gts, lrs, M = _setup(dev)
g = pa.activate()
n_views = 4
local = []
for r in range(world):
    a = mv.GradArena(pa.P, M, dev)
    vs = mv.shard_views(n_views, r, world)
    mv.cuda_views_fwd_bwd(g, [settings[v] for v in vs], [ViewLoss(gts[v], 0.2, weight=0.25) for v in vs], a)
    local.append(a)
torch.cuda.synchronize()
exp = local[0].storage.clone()
exp[:local[0]._n_f32] += local[1].storage[:local[1]._n_f32]
exp_vis = local[0].visible_count + local[1].visible_count
exp_rad = torch.maximum(local[0].max_radii, local[1].max_radii)

for method in ("nvls", "nccl", "nvls"):
    arena = mv.GradArena(pa.P, M, dev, symmetric=True)
    arena.method = method if arena._mc else "nccl"
    arena.storage.fill_(3.0)
    arena.visible_count.fill_(5)
    mine = mv.shard_views(n_views, rank, world)
    mv.cuda_views_fwd_bwd(g, [settings[v] for v in mine], [ViewLoss(gts[v], 0.2, weight=0.25) for v in mine], arena, all_reduce=True)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        out = {}
        for name, v in arena.views.items():
            o = arena._offs[name]
            e = exp[o:o + v.numel()].view_as(v)
            d = (v - e).abs()
            out[name] = (int((d > 0).sum()), float(d.max()))
        out["gnorm"] = (int((arena.grad_norm_accum != exp[arena._n_flat:arena._n_flat + arena.P]).sum()),)
        out["vis"] = (int((arena.visible_count != exp_vis).sum()),)
        out["rad"] = (int((arena.max_radii != exp_rad).sum()),)
        print(method, "uses_nvls", arena.uses_nvls, out, flush=True)
    dist.barrier()
dist.destroy_process_group()
