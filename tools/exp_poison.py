"""Does the fused training step read memory nobody wrote?  The same two steps on a fresh allocator pool and on a pool poisoned
with 0xFF bytes (what tests/conftest.py does before every GPU test), under the library's switches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multiview_inpaint_b200 import _C, multiview as mv  # noqa: E402
from multiview_inpaint_b200.trainstep import ViewLoss, fused_train_step  # noqa: E402
from tests.test_trainstep_world2_gpu import _setup  # noqa: E402

dev = torch.device("cuda", 0)


def poison(byte):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    junk = torch.full((512 << 20,), byte, dtype=torch.uint8, device=dev)
    del junk


def single(flags, n_views=4, steps=2):
    pa, settings, gts, lrs, M = _setup(dev)
    arena = mv.GradArena(pa.P, M, dev)
    losses = [ViewLoss(gt, 0.2, weight=1.0 / n_views) for gt in gts[:n_views]]
    for _ in range(steps):
        fused_train_step(pa, settings[:n_views], losses, arena, lrs, flags=flags)
    torch.cuda.synchronize()
    return pa.param.cpu().clone()


def frac(a, b):
    d = (a - b).abs()
    return round((d > 1e-6).float().mean().item(), 5), round(d.max().item(), 5)


for name, flags, knobs in (("default", None, {}), ("literal lists", 0, {}), ("shuffle K7", None, {3: 0}), ("tile_count atomics", None, {1: 1}),
                           ("precise", 2 | 32, {}), ("literal + atomics + shuffle", 0, {1: 1, 3: 0})):
    for k, v in knobs.items():
        _C.debug_set(k, v)
    poison(0)
    a = single(flags)
    poison(0xFF)
    b = single(flags)
    poison(0x7F)
    c = single(flags)
    print(f"{name:28s} zero pool vs 0xFF pool {frac(a, b)}   vs 0x7F pool {frac(a, c)}", flush=True)
    for k in knobs:
        _C.debug_set(k, {3: 1, 1: 2}[k])
