"""Where the 8-GPU step goes, and which settings of the exchange are best (run under torchrun at N = 2 / 4 / 8):
one process per GPU builds the headline scene ONCE and times the product step (batched front end + blend, chunked
K8+K9 pipelined with the in-switch all-reduce) for a sweep of (chunks, CTAs of the reduction kernel), plus the
decomposition views-only / K8+K9-only / reduction-only.  Device times by CUDA events, max over ranks.  JSON to stdout
(rank 0).  usage: torchrun --nproc-per-node N tools/exp_scale8.py [steps]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C, multiview as mv, scenes as S  # noqa: E402
from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
name = "headline"
cfg = S.CONFIGS[name]
sc = S.make_config_scene(name)
P, W, H, M, D = sc["P"], sc["W"], sc["H"], sc["M"], sc["sh_degree"]
gauss = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
vpr = 4
n_views = vpr * world
cams = S.orbit_cameras(n_views, W, H, max_deg=5.0)
mine = mv.shard_views(n_views, rank, world)
bg = torch.zeros(3, device=dev)
cams_dev = {v: cams[v].to(dev) for v in mine}
wts = {v: S.loss_weights(W, H, cfg["seed"] + v).to(dev) for v in mine}


def settings(c):
    return GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg, scale_modifier=1.0,
                                         viewmatrix=c.world_view_transform, projmatrix=c.full_proj_transform, sh_degree=D,
                                         campos=c.camera_center, prefiltered=False)


arena = mv.GradArena(P, M, dev, symmetric=True)
av = mv.AsyncViews(n_views)
for v in mine:
    r = mv.cuda_view_fwd_bwd(gauss, settings(cams_dev[v]), lambda c, v=v: wts[v], arena, capacity=0)
    av.learn(v, r.num_rendered)
wss = [_C.Workspace(dev) for _ in mine]
throttle = mv.StepThrottle(2)
sets = [settings(cams_dev[v]) for v in mine]
fns = [lambda c, v=v: wts[v] for v in mine]
caps = [av.capacity(v) for v in mine]
slots = [av.slot(v) for v in mine]


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def full_step(chunks, all_reduce=True):
    def f():
        mv.cuda_views_fwd_bwd(gauss, sets, fns, arena, capacities=caps, async_results=slots, all_reduce=all_reduce, chunks=chunks,
                              workspaces=wss, batched=True)
        throttle.tick(dev)
    return f


out = {"world": world, "nvls": bool(arena._mc), "arena_mb": arena.storage.numel() * 4 / 1e6, "steps": steps, "variants": []}
ms_local = timed(full_step(1, all_reduce=False), steps)
out["no_exchange_ms"] = ms_local
methods = (["nvls"] if arena._mc else []) + ["nccl"]
for method in methods:
    arena.method = method
    for blocks in ((0, 148, 74, 37) if method == "nvls" else (0,)):
        arena.nvls_blocks = blocks
        # the reduction alone (all ranks idle otherwise)
        ar_ms = timed(lambda: arena.all_reduce(), 10)
        for chunks in ((1, 4, 8, 16) if method == "nvls" else (1,)):
            ms = timed(full_step(chunks), steps)
            out["variants"].append({"method": method, "nvls_blocks": blocks, "chunks": chunks, "allreduce_alone_ms": ar_ms, "step_ms": ms,
                                    "views_per_s": n_views / (ms / 1e3), "exposed_ms": ms - ms_local})
            if rank == 0:
                print(json.dumps(out["variants"][-1]), flush=True)
torch.cuda.synchronize()
assert not av.check(mine)
if rank == 0:
    best = min(out["variants"], key=lambda v: v["step_ms"])
    out["best"] = best
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
