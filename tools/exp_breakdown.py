import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C, multiview as mv, scenes as S
from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings
dev = torch.device("cuda")
sc = S.make_config_scene("headline")
P,W,H,M,D = sc["P"],sc["W"],sc["H"],sc["M"],sc["sh_degree"]
gauss = {k: sc[k].to(dev) for k in ("means3D","shs","opacities","scales","rotations")}
cam = sc["camera"].to(dev); bg = torch.zeros(3, device=dev)
rs = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=D, campos=cam.camera_center, prefiltered=False)
wt = S.loss_weights(W,H,6).to(dev)
arena = mv.GradArena(P, M, dev)
def one(stats=True):
    r = mv.cuda_view_fwd_bwd(gauss, rs, lambda c: wt, arena)
    return r
def timeit(fn, n=10, label=""):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t=time.perf_counter()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{label:40s} wall {1000*(time.perf_counter()-t)/n:8.3f} ms  gpu-events {e0.elapsed_time(e1)/n:8.3f} ms")
r = one(); print("N", r.num_rendered, "V", int((r.radii>0).sum()))
timeit(one, label="view fwd+bwd accumulate (mv)")
# forward only
e = torch.empty(0, device=dev)
def fwd():
    return _C.rasterize_gaussians(rs.bg, gauss["means3D"], e, gauss["opacities"], gauss["scales"], gauss["rotations"], 1.0, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, H, W, gauss["shs"], D, rs.campos, False)
timeit(fwd, label="forward only")
out = fwd()
def bwd(flags=0, o=None):
    n, color, radii, geom, binning, img, depth = out
    return _C.rasterize_gaussians_backward(rs.bg, gauss["means3D"], radii, e, gauss["scales"], gauss["rotations"], 1.0, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, wt, gauss["shs"], D, rs.campos, geom, n, binning, img, flags=flags, out=o)
timeit(bwd, label="backward only (assign, torch.empty outs)")
timeit(lambda: bwd(8, arena.views), label="backward only (accumulate into arena)")
timeit(lambda: arena.zero_(), label="arena.zero_")
n, color, radii, geom, binning, img, depth = out
g = bwd()
timeit(lambda: arena.add_view_stats(g[0], radii), label="add_view_stats")
_C.profile_enable(True)
timeit(one, label="view fwd+bwd with profiler on")
ms, cnt = _C.profile_collect(); _C.profile_enable(False)
print({k: round(ms[k]/max(cnt[k],1),3) for k in ms})
# autograd path
from diff_gaussian_rasterization import GaussianRasterizer
leaves = {k: v.clone().requires_grad_(True) for k,v in gauss.items()}
def autograd_step():
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth = GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    (color*wt).sum().backward()
    for v in leaves.values(): v.grad = None
timeit(autograd_step, label="autograd single-view step (train.py style)")
