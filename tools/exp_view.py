"""Runs n fwd+bwd views of the headline scene through the multi-view path (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import multiview as mv, scenes as S
from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
wl = sys.argv[2] if len(sys.argv) > 2 else "headline"
dev = torch.device("cuda")
sc = S.make_config_scene(wl)
P, W, H, M, D = sc["P"], sc["W"], sc["H"], sc["M"], sc["sh_degree"]
gauss = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
cam = sc["camera"].to(dev)
rs = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=torch.zeros(3, device=dev),
                                   scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
                                   sh_degree=D, campos=cam.camera_center, prefiltered=False)
wt = S.loss_weights(W, H, 6).to(dev)
arena = mv.GradArena(P, M, dev)
for _ in range(n):
    r = mv.cuda_view_fwd_bwd(gauss, rs, lambda c: wt, arena)
torch.cuda.synchronize()
print("N", r.num_rendered)
