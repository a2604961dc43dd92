"""One-GPU emulation of tests/test_trainstep_world2_gpu.py (four views over two ranks vs one rank): the two ranks' arenas are
produced one after the other on the same device and summed where the all-reduce would.  Prints the test's metric (fraction of
parameters that differ by more than 1e-6 after two fused steps) for: the same single-rank run twice (run-to-run noise), the
emulated 2-rank run, and the same under the library's switches -- to find which change of the round moved it."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multiview_inpaint_b200 import _C, multiview as mv  # noqa: E402
from multiview_inpaint_b200.trainstep import ViewLoss, fused_train_step  # noqa: E402
from tests.test_trainstep_world2_gpu import _setup  # noqa: E402

dev = torch.device("cuda", 0)


def single(flags, n_views=4):
    pa, settings, gts, lrs, M = _setup(dev)
    arena = mv.GradArena(pa.P, M, dev)
    losses = [ViewLoss(gt, 0.2, weight=1.0 / n_views) for gt in gts[:n_views]]
    for _ in range(2):
        fused_train_step(pa, settings[:n_views], losses, arena, lrs, flags=flags)
    torch.cuda.synchronize()
    return pa.param.cpu().clone()


def two_ranks(flags, n_views=4):
    pa, settings, gts, lrs, M = _setup(dev)
    arenas = [mv.GradArena(pa.P, M, dev) for _ in range(2)]
    for _ in range(2):
        g = pa.activate()
        for r in range(2):
            mine = mv.shard_views(n_views, r, 2)
            losses = [ViewLoss(gts[v], 0.2, weight=1.0 / n_views) for v in mine]
            mv.cuda_views_fwd_bwd(g, [settings[v] for v in mine], losses, arenas[r], flags=flags)
        arenas[0].flat += arenas[1].flat
        arenas[0].grad_norm_accum += arenas[1].grad_norm_accum
        arenas[0].visible_count += arenas[1].visible_count
        torch.maximum(arenas[0].max_radii, arenas[1].max_radii, out=arenas[0].max_radii)
        pa.apply_gradients(arenas[0], lrs)
    torch.cuda.synchronize()
    return pa.param.cpu().clone()


def frac(a, b):
    d = (a - b).abs()
    return round((d > 1e-6).float().mean().item(), 5), round(d.max().item(), 5)


for name, flags, knobs in (("default", None, {}), ("literal lists", 0, {}), ("shuffle K7", None, {3: 0}), ("tile_count atomics", None, {1: 1}),
                           ("precise", 2 | 32, {})):
    for k, v in knobs.items():
        _C.debug_set(k, v)
    a, b, c = single(flags), single(flags), two_ranks(flags)
    print(f"{name:20s} single vs single {frac(a, b)}   single vs two ranks {frac(a, c)}", flush=True)
    for k in knobs:
        _C.debug_set(k, {3: 1, 1: 2}[k])

# ---- stale arena contents: which elements does one non-accumulating multi-view backward leave untouched? ----
pa, settings, gts, lrs, M = _setup(dev)
g = pa.activate()
for nv in (2, 4):
    arena = mv.GradArena(pa.P, M, dev)
    arena.storage.fill_(3.0)
    arena.visible_count.fill_(5)
    arena.max_radii.fill_(7)
    losses = [ViewLoss(gts[v], 0.2, weight=0.25) for v in range(nv)]
    mv.cuda_views_fwd_bwd(g, settings[:nv], losses, arena)
    torch.cuda.synchronize()
    print("views", nv, {k: int((v == 3.0).sum()) for k, v in arena.views.items()}, "gnorm", int((arena.grad_norm_accum == 3.0).sum()),
          "vis==5", int((arena.visible_count == 5).sum()), "rad==7", int((arena.max_radii == 7).sum()), flush=True)
    ref = mv.GradArena(pa.P, M, dev)
    mv.cuda_views_fwd_bwd(g, settings[:nv], losses, ref)
    print("   max |stale-start - zero-start|", {k: float((arena.views[k] - ref.views[k]).abs().max()) for k in arena.views})
