"""TWO ranks, torchrun: the world-2 fused-step test repeated, with the cross-rank comparison made after EVERY stage of every step
(reduced arena before Adam, parameters after Adam) so that a divergence names the stage it came from."""
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

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n_views = 4
events = []


def same_everywhere(t):
    mine = t.detach().contiguous().view(-1).view(torch.int32)
    other = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(other, mine)
    return int((other[0] != other[1]).sum())


for rep in range(reps):
    pa, settings, gts, lrs, M = _setup(dev)
    mine = mv.shard_views(n_views, rank, world)
    arena = mv.GradArena(pa.P, M, dev, symmetric=True)
    arena.flat.fill_(3.0)
    arena.visible_count.fill_(5)
    losses = [ViewLoss(gts[v], 0.2, weight=1.0 / n_views) for v in mine]
    if rep % 2 == rank:
        junk = torch.randn(1024, 1024, device=dev)
        for _ in range(rep % 5):
            junk = junk @ junk * 1e-3
    for step in range(2):
        fused_train_step(pa, [settings[v] for v in mine], losses, arena, lrs, all_reduce=True, apply=False)
        a = same_everywhere(arena.storage)
        pa.apply_gradients(arena, lrs)
        b = same_everywhere(pa.param)
        if a or b:
            events.append((rep, step, a, b))
torch.cuda.synchronize()
# the last repetition against one rank doing all four views (the test's reference)
pb, settings, gts, lrs, M = _setup(dev)
ref = mv.GradArena(pb.P, M, dev)
for _ in range(2):
    fused_train_step(pb, settings[:n_views], [ViewLoss(gt, 0.2, weight=1.0 / n_views) for gt in gts[:n_views]], ref, lrs)
torch.cuda.synchronize()
d = (pa.param - pb.param).abs()
if rank == 0:
    print("reps", reps, "divergences (rep, step, arena words, param words):", events, "| vs single rank: frac > 1e-6",
          round(float((d > 1e-6).float().mean()), 5), "max", float(d.max()), flush=True)
dist.destroy_process_group()
