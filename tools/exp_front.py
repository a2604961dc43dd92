"""One batched step of the headline config (4 views, asynchronous, caller-owned workspaces) repeated a few times: the
command ncu captures the front-end kernels from (`ncu --set full -k regex:... python tools/exp_front.py 3`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C, multiview as mv, scenes as S  # noqa: E402
from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
workload = sys.argv[2] if len(sys.argv) > 2 else "headline"
nv = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda")
if os.environ.get("GSR_RANK"):
    _C.debug_set(0, int(os.environ["GSR_RANK"]))
sc = S.make_config_scene(workload)
P, W, H, M, D = sc["P"], sc["W"], sc["H"], sc["M"], sc["sh_degree"]
gauss = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
cams = [c.to(dev) for c in S.orbit_cameras(nv, W, H, max_deg=5.0)]
bg = torch.zeros(3, device=dev)
rss = [GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                                     scale_modifier=1.0, viewmatrix=c.world_view_transform, projmatrix=c.full_proj_transform,
                                     sh_degree=D, campos=c.camera_center, prefiltered=False) for c in cams]
wts = [S.loss_weights(W, H, 6 + v).to(dev) for v in range(nv)]
arena = mv.GradArena(P, M, dev)
av = mv.AsyncViews(nv)
for v in range(nv):
    r = mv.cuda_view_fwd_bwd(gauss, rss[v], lambda c, v=v: wts[v], arena, capacity=0)
    av.learn(v, r.num_rendered)
wss = [_C.Workspace(dev) for _ in range(nv)]
for _ in range(steps):
    mv.cuda_views_fwd_bwd(gauss, rss, [lambda c, v=v: wts[v] for v in range(nv)], arena,
                          capacities=[av.capacity(v) for v in range(nv)], async_results=[av.slot(v) for v in range(nv)],
                          workspaces=wss, batched=True)
torch.cuda.synchronize()
assert not av.check(range(nv))
print("ok", steps, workload, nv)
