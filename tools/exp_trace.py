"""Kernel timeline of the resident multi-view step (what bench.py times) from torch.profiler (CUPTI):
which kernels overlap across the pipeline's streams and where the GPU idles.
  python tools/exp_trace.py [streams] [views]  ->  gpurun_out/trace_kernels.csv  + a per-step summary on stdout"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C, multiview as mv, scenes as S
from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings
from torch.profiler import profile, ProfilerActivity

n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 2
vpr = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda")
cfg = S.CONFIGS["headline"]; sc = S.make_config_scene("headline")
P, W, H, M, D = sc["P"], sc["W"], sc["H"], sc["M"], sc["sh_degree"]
gauss = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
cams = [c.to(dev) for c in S.orbit_cameras(vpr, W, H, max_deg=5.0)]
bg = torch.zeros(3, device=dev)
wts = [S.loss_weights(W, H, cfg["seed"] + v).to(dev) for v in range(vpr)]
def settings(c):
    return GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg, scale_modifier=1.0,
                                         viewmatrix=c.world_view_transform, projmatrix=c.full_proj_transform, sh_degree=D,
                                         campos=c.camera_center, prefiltered=False)
arena = mv.GradArena(P, M, dev)
av = mv.AsyncViews(vpr)
for v in range(vpr):
    r = mv.cuda_view_fwd_bwd(gauss, settings(cams[v]), lambda c, v=v: wts[v], arena, capacity=0)
    av.learn(v, r.num_rendered)
pipe = mv.ViewPipeline(dev, n_streams) if n_streams > 1 else None
wss = [_C.Workspace(dev) for _ in range(vpr)]
thr = mv.StepThrottle(2)
def step():
    mv.cuda_views_fwd_bwd(gauss, [settings(c) for c in cams], [lambda c, v=v: wts[v] for v in range(vpr)], arena,
                          capacities=[av.capacity(v) for v in range(vpr)], async_results=[av.slot(v) for v in range(vpr)],
                          pipeline=pipe, all_reduce=True, workspaces=wss)
    thr.tick(dev)
for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record(); torch.cuda.synchronize()
print(f"untraced: {e0.elapsed_time(e1) / 10:.3f} ms/step ({vpr} views, {n_streams} streams)")
os.makedirs("gpurun_out", exist_ok=True)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
path = "/tmp/trace.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ks.sort(key=lambda e: e["ts"])
t0 = ks[0]["ts"]
with open(f"gpurun_out/trace_kernels_s{n_streams}.csv", "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for e in ks:
        f.write(f"{e['ts'] - t0:.1f},{e['dur']:.1f},{e['args'].get('stream')},\"{e['name'][:70]}\"\n")
# host-side launch activity: when did the CPU issue each cudaLaunchKernel
ls = [e for e in ev if e.get("cat") == "cuda_runtime" and "Launch" in e.get("name", "")]
ls.sort(key=lambda e: e["ts"])
with open(f"gpurun_out/trace_launches_s{n_streams}.csv", "w") as f:
    f.write("start_us,dur_us,name\n")
    for e in ls:
        f.write(f"{e['ts'] - t0:.1f},{e['dur']:.1f},{e['name']}\n")
# GPU busy / idle over the traced window (union of kernel intervals)
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ks)
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s, t in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s; cur_s, cur_e = s, t
    else:
        cur_e = max(cur_e, t)
busy += cur_e - cur_s
span = iv[-1][1] - iv[0][0]
print(f"traced 3 steps: span {span / 1000:.3f} ms, >=1 kernel resident {busy / 1000:.3f} ms ({100 * busy / span:.1f} %), "
      f"sum of kernel durations {sum(e['dur'] for e in ks) / 1000:.3f} ms, {len(ks)} GPU activities, {len(ls)} launches")
