"""Two views in flight on two streams: does the latency-bound front end of view v+1 hide behind the blend of view v?"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C, multiview as mv, scenes as S
from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings
dev = torch.device("cuda")
sc = S.make_config_scene("headline")
P, W, H, M, D = sc["P"], sc["W"], sc["H"], sc["M"], sc["sh_degree"]
gauss = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
cams = [c.to(dev) for c in S.orbit_cameras(4, W, H, max_deg=5.0)]
bg = torch.zeros(3, device=dev)
def settings(cam):
    return GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=D, campos=cam.camera_center, prefiltered=False)
rss = [settings(c) for c in cams]
wt = S.loss_weights(W, H, 6).to(dev)
arena = mv.GradArena(P, M, dev)
av = mv.AsyncViews(4)
for v in range(4):
    r = mv.cuda_view_fwd_bwd(gauss, rss[v], lambda c: wt, arena, capacity=0)
    av.learn(v, r.num_rendered)
e = torch.empty(0, device=dev)
def fwd(v):
    rs = rss[v]
    return _C.rasterize_gaussians(rs.bg, gauss["means3D"], e, gauss["opacities"], gauss["scales"], gauss["rotations"], 1.0, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, H, W, gauss["shs"], D, rs.campos, False, capacity=av.capacity(v), async_result=av.slot(v))
def bwd(v, out):
    rs = rss[v]
    n, color, radii, geom, binning, img, depth = out
    g = _C.rasterize_gaussians_backward(rs.bg, gauss["means3D"], radii, e, gauss["scales"], gauss["rotations"], 1.0, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, wt, gauss["shs"], D, rs.campos, geom, n, binning, img, flags=_C.FLAG_ACCUMULATE, out=arena.views)
    arena.add_view_stats(g[0], radii)
def step_serial():
    for v in range(4):
        bwd(v, fwd(v))
def make_step_streams(streams):
    main = torch.cuda.current_stream()
    def step():
        start = torch.cuda.Event(); start.record(main)
        prev_bwd = None
        for v in range(4):
            st = streams[v % len(streams)]
            st.wait_event(start)
            with torch.cuda.stream(st):
                out = fwd(v)
                if prev_bwd is not None:
                    st.wait_event(prev_bwd)
                bwd(v, out)
                prev_bwd = torch.cuda.Event(); prev_bwd.record(st)
                for t in out[1:]:
                    t.record_stream(st)
        for st in streams:
            ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
    return step
def run(fn, n=10, label=""):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{label:44s} {e0.elapsed_time(e1)/n/4:7.3f} ms/view", flush=True)
run(step_serial, label="serial, one stream")
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
run(make_step_streams([torch.cuda.Stream(), torch.cuda.Stream()]), label="2 streams alternating")
run(make_step_streams([torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()]), label="3 streams alternating")
run(make_step_streams([torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()]), label="4 streams alternating")
def make_step_fb(prio_f, prio_b):
    F, B = torch.cuda.Stream(priority=prio_f), torch.cuda.Stream(priority=prio_b)
    main = torch.cuda.current_stream()
    def step():
        start = torch.cuda.Event(); start.record(main)
        F.wait_event(start); B.wait_event(start)
        outs, evs = {}, {}
        def issue_fwd(v):
            with torch.cuda.stream(F):
                outs[v] = fwd(v)
                evs[v] = torch.cuda.Event(); evs[v].record(F)
        issue_fwd(0)
        for v in range(4):
            if v + 1 < 4: issue_fwd(v + 1)
            with torch.cuda.stream(B):
                B.wait_event(evs[v])
                bwd(v, outs[v])
                for t in outs[v][1:]: t.record_stream(B)
            del outs[v]
        for st in (F, B):
            ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
    return step
run(make_step_fb(0, 0), label="fwd stream / bwd stream, equal priority")
run(make_step_fb(-1, 0), label="fwd stream HIGH priority / bwd stream")
run(make_step_fb(0, -1), label="fwd stream / bwd stream HIGH priority")
run(step_serial, label="serial again")
assert not av.check(range(4))
