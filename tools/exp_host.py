"""Host-side cost per view vs GPU time (is the multi-view loop host-bound?)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C, multiview as mv, scenes as S
from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings
dev = torch.device("cuda")
sc = S.make_config_scene("headline")
P, W, H, M, D = sc["P"], sc["W"], sc["H"], sc["M"], sc["sh_degree"]
gauss = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
cam = sc["camera"].to(dev); bg = torch.zeros(3, device=dev)
rs = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=D, campos=cam.camera_center, prefiltered=False)
wt = S.loss_weights(W, H, 6).to(dev)
arena = mv.GradArena(P, M, dev)
av = mv.AsyncViews(1)
r = mv.cuda_view_fwd_bwd(gauss, rs, lambda c: wt, arena, capacity=0)
av.learn(0, r.num_rendered)
print("nproc", os.cpu_count(), "N", r.num_rendered)
e = torch.empty(0, device=dev)
def view():
    mv.cuda_view_fwd_bwd(gauss, rs, lambda c: wt, arena, capacity=av.capacity(0), async_result=av.slot(0))
def view_parts(acc):
    t0 = time.perf_counter()
    out = _C.rasterize_gaussians(rs.bg, gauss["means3D"], e, gauss["opacities"], gauss["scales"], gauss["rotations"], 1.0, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, H, W, gauss["shs"], D, rs.campos, False, capacity=av.capacity(0), async_result=av.slot(0))
    t1 = time.perf_counter()
    n, color, radii, geom, binning, img, depth = out
    g = _C.rasterize_gaussians_backward(rs.bg, gauss["means3D"], radii, e, gauss["scales"], gauss["rotations"], 1.0, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, wt, gauss["shs"], D, rs.campos, geom, n, binning, img, flags=_C.FLAG_ACCUMULATE, out=arena.views)
    t2 = time.perf_counter()
    arena.add_view_stats(g[0], radii)
    t3 = time.perf_counter()
    acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2
def run(fn, n=20, label=""):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    th = time.perf_counter() - t
    e1.record(); torch.cuda.synchronize()
    print(f"reserved {torch.cuda.memory_reserved()/2**30:.2f} GiB", end="  ")
    print(f"{label:36s} host-issue {1000*th/n:7.3f} ms/view   gpu {e0.elapsed_time(e1)/n:7.3f} ms/view")
run(view, label="cuda_view_fwd_bwd async")
acc = [0, 0, 0]
run(lambda: view_parts(acc), label="parts async")
print("host ms per call (23 calls): fwd %.3f bwd %.3f stats %.3f" % tuple(1000 * a / 23 for a in acc))
# same, after a deliberate sync each view (GPU never waits on the host queue; host waits)
def view_sync():
    view(); torch.cuda.synchronize()
run(view_sync, label="async + sync per view")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20): view()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumtime").print_stats(18)
