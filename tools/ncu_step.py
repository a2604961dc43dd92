"""ONE steady-state multi-view step of the product path (batched front end, one blend launch per direction, batched
K8+K9) bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off` captures:
    ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_full \
        python tools/ncu_step.py [workload] [views]
The capacities are learned with an exact-size pass and two warm steps run before the bracket, so the captured launches
are exactly those of a timed bench step (asynchronous, caller-owned workspaces, no allocation)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C, multiview as mv, scenes as S  # noqa: E402
from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "headline"
n_views = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda")
sc = S.make_config_scene(wl)
cfg = S.CONFIGS[wl]
P, W, H, M, D = sc["P"], sc["W"], sc["H"], sc["M"], sc["sh_degree"]
gauss = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
cams = [c.to(dev) for c in S.orbit_cameras(n_views, W, H, max_deg=5.0)]
bg = torch.zeros(3, device=dev)
wts = [S.loss_weights(W, H, cfg["seed"] + v).to(dev) for v in range(n_views)]


def settings(c):
    return GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg, scale_modifier=1.0,
                                         viewmatrix=c.world_view_transform, projmatrix=c.full_proj_transform, sh_degree=D,
                                         campos=c.camera_center, prefiltered=False)


arena = mv.GradArena(P, M, dev)
av = mv.AsyncViews(n_views)
for v in range(n_views):
    r = mv.cuda_view_fwd_bwd(gauss, settings(cams[v]), lambda c, v=v: wts[v], arena, capacity=0)
    av.learn(v, r.num_rendered)
ws = [_C.Workspace(dev) for _ in range(n_views)]


def step():
    mv.cuda_views_fwd_bwd(gauss, [settings(c) for c in cams], [lambda c, v=v: wts[v] for v in range(n_views)], arena,
                          capacities=[av.capacity(v) for v in range(n_views)], async_results=[av.slot(v) for v in range(n_views)],
                          workspaces=ws, batched=True)


for _ in range(2):
    step()
torch.cuda.synchronize()
assert not av.check(range(n_views))
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("captured one", n_views, "view step of", wl, "N", [int(av.slots[v, 0]) for v in range(n_views)])
