"""EXPERIMENT (written at the end of round 1, NOT yet run on a GPU -- DESIGN.md section 10 item 4): does capturing the
asynchronous multi-view step in a CUDA graph close the gap between the device-resident number and the end-to-end one?

The steady-state step is graph-friendly by construction: GSR_FLAG_ASYNC (no host read inside), fixed capacities,
caller-owned Workspaces (no allocation), cameras and loss weights staged from pinned memory into fixed device buffers.
What is measured, per step of 4 headline views, device time by CUDA events, each with the per-step loss read-back
(`.item()`) that drains the queue -- the thing that hurts the eager loop:

    eager      cuda_views_fwd_bwd issued from Python every step (what bench.py's e2e leg does)
    graph      the same step captured once (torch.cuda.graph on a side stream, the ViewPipeline's streams fork from and
               join it), then one cudaGraphLaunch per step

Checks that the arena after a graph replay equals the arena after an eager step (bit pattern, up to the atomics' order:
1e-3 relative like every gradient comparison).  usage: python tools/exp_graph.py [steps]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C, multiview as mv, scenes as S  # noqa: E402
from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda")
sc = S.make_config_scene("headline")
P, W, H, M, D = sc["P"], sc["W"], sc["H"], sc["M"], sc["sh_degree"]
gauss = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
n_views = 4
cams_cpu = S.orbit_cameras(n_views, W, H, max_deg=5.0)
bg = torch.zeros(3, device=dev)
views = list(range(n_views))

# pinned host inputs and fixed device staging buffers (what bench.py's e2e leg uses)
wts_cpu = [S.loss_weights(W, H, 6 + v).pin_memory() for v in views]
cam_pinned = [torch.cat([c.world_view_transform.reshape(-1), c.full_proj_transform.reshape(-1),
                         c.camera_center.reshape(-1)]).pin_memory() for c in cams_cpu]
cam_stage = [torch.empty(35, device=dev) for _ in views]
wt_stage = [torch.empty(3, H, W, device=dev) for _ in views]
loss_dev = torch.zeros((), device=dev)


def settings(v):
    c = cams_cpu[v]
    return GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                                         scale_modifier=1.0, viewmatrix=cam_stage[v][:16].view(4, 4),
                                         projmatrix=cam_stage[v][16:32].view(4, 4), sh_degree=D, campos=cam_stage[v][32:35],
                                         prefiltered=False)


arena = mv.GradArena(P, M, dev)
av = mv.AsyncViews(n_views)
workspaces = [_C.Workspace(dev) for _ in views]
pipe = mv.ViewPipeline(dev, depth=2)

# learn the capacities with one exact-size step
for v in views:
    cam_stage[v].copy_(cam_pinned[v])
    wt_stage[v].copy_(wts_cpu[v])
for v in views:
    r = mv.cuda_view_fwd_bwd(gauss, settings(v), lambda c, v=v: wt_stage[v], arena, capacity=0)
    av.learn(v, r.num_rendered)
torch.cuda.synchronize()


def stage_inputs():
    for v in views:
        cam_stage[v].copy_(cam_pinned[v], non_blocking=True)
        wt_stage[v].copy_(wts_cpu[v], non_blocking=True)


def body():
    """everything of one step that runs on the device: H2D staging, 4 views, batched K8+K9, the loss reduction"""
    stage_inputs()
    loss_dev.zero_()
    parts = []

    def grad_fn(v):
        def f(color):
            parts.append((color * wt_stage[v]).sum())
            return wt_stage[v]
        return f
    mv.cuda_views_fwd_bwd(gauss, [settings(v) for v in views], [grad_fn(v) for v in views], arena,
                          capacities=[av.capacity(v) for v in views], async_results=[av.slot(v) for v in views],
                          pipeline=pipe, workspaces=workspaces)
    # the per-view partial sums were produced on the pipeline's streams, joined to this one by pipe.step()
    loss_dev.add_(torch.stack(parts).sum())


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def eager_step():
    body()
    return float(loss_dev.item())          # D2H of the step's result: drains the queue


out = {"views_per_step": n_views, "steps": steps}
l0 = _C.kernel_launches()
eager_loss = eager_step()
out["launches_per_step"] = _C.kernel_launches() - l0
ref_flat = arena.flat.clone()
out["eager_ms_per_step"] = timed(eager_step, steps)
assert not av.check(views), "capacity overflow in the eager loop"

# ---- capture ----
graph = torch.cuda.CUDAGraph()
side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream(dev))
try:
    with torch.cuda.stream(side):
        body()                              # warm-up on the capture stream (allocator pools, lazy module loads)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=side):
        body()
    captured = True
except Exception as ex:                     # report what blocks the capture instead of dying: that is the finding
    captured = False
    out["capture_error"] = repr(ex)[:500]

if captured:
    def graph_step():
        graph.replay()
        return float(loss_dev.item())
    graph_loss = graph_step()
    torch.cuda.synchronize()
    err = float((arena.flat - ref_flat).abs().max() / (ref_flat.abs().max() + 1e-30))
    out["graph_vs_eager_arena_rel_err"] = err
    out["graph_vs_eager_loss"] = [graph_loss, eager_loss]
    out["graph_ms_per_step"] = timed(graph_step, steps)
    out["speedup"] = out["eager_ms_per_step"] / out["graph_ms_per_step"]
    out["views_per_s"] = {"eager": n_views / out["eager_ms_per_step"] * 1e3, "graph": n_views / out["graph_ms_per_step"] * 1e3}
    assert not av.check(views), "capacity overflow in the graph loop"
print(json.dumps(out))
