"""Measures the training-step stages either side of the rasterizer at BASELINE.json's headline
shape -- 3 M Gaussians, SH degree 3 (M = 16), 1600x1008 image -- against the torch structure the reference runs
(gs-simp/utils/loss_utils.py through F.conv2d + autograd; getters of scene/gaussian_model.py:95-115 + autograd;
torch.optim.Adam over six groups).  CUDA events on the launching stream, 3 warm-ups; an L2 flush (512 MB write)
between iterations for the image-sized kernels (their working set is below the 126 MB L2; the per-Gaussian arrays are
far above it).  Prints one JSON object; `python tools/exp_trainstep.py [P] [iters]`.

usage (GPU box):  python tools/exp_trainstep.py > gpurun_out/trainstep.json
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C  # noqa: E402
from multiview_inpaint_b200.multiview import GradArena  # noqa: E402
from multiview_inpaint_b200.trainstep import GaussianParamArena  # noqa: E402

DEV = "cuda"
P = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
M, H, W = 16, 1008, 1600
PEAK = 6552.3
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

_flush = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, flush=False, iters=ITERS, warm=3):
    """mean milliseconds per call, CUDA events on the current stream around each call"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush:
            _flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def row(name, ms, alg_bytes, ms_torch=None):
    gbs = alg_bytes / ms / 1e6
    d = {"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": int(alg_bytes), "achieved_gbs": round(gbs, 1),
         "peak_gbs": PEAK, "frac": round(gbs / PEAK, 3)}
    if ms_torch is not None:
        d["torch_structure_ms"] = round(ms_torch, 4)
        d["speedup_vs_torch_structure"] = round(ms_torch / ms, 2)
    return d


out = {"P": P, "M": M, "H": H, "W": W, "iters": ITERS, "rows": []}
torch.manual_seed(0)

# ---------------- loss ----------------
gt = torch.rand(3, H, W, device=DEV)
img = (gt + 0.05 * torch.randn_like(gt)).clamp(0, 1).contiguous()
temp = torch.empty(_C.loss_temp_bytes(3, H, W), dtype=torch.uint8, device=DEV)
out3 = torch.empty(3, device=DEV)
dL = torch.empty_like(img)
npx = 3 * H * W
ms_f = timed(lambda: _C.loss_l1_ssim_forward(img, gt, 0.2, out3=out3, temp=temp), flush=True)
ms_b = timed(lambda: _C.loss_l1_ssim_backward(img, gt, 0.2, temp, out=dL), flush=True)

w1 = torch.tensor([__import__("math").exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)], device=DEV)
w1 = w1 / w1.sum()
w2 = torch.outer(w1, w1).expand(3, 1, 11, 11).contiguous()


def torch_loss(x):
    blur = lambda t: F.conv2d(t[None], w2, padding=5, groups=3)[0]
    mx, my = blur(x), blur(gt)
    vx, vy, cxy = blur(x * x) - mx * mx, blur(gt * gt) - my * my, blur(x * gt) - mx * my
    s = ((2 * mx * my + 1e-4) * (2 * cxy + 9e-4)) / ((mx * mx + my * my + 1e-4) * (vx + vy + 9e-4))
    return 0.8 * (x - gt).abs().mean() + 0.2 * (1 - s.mean())


def torch_loss_fb():
    x = img.detach().requires_grad_(True)
    torch_loss(x).backward()
    return x.grad


ms_t = timed(torch_loss_fb, flush=True)
out["rows"].append(row("loss forward (ssim_l1_forward + finalize)", ms_f, 20 * npx))
out["rows"].append(row("loss backward (ssim_l1_backward)", ms_b, 24 * npx))
out["rows"].append(row("loss forward+backward", ms_f + ms_b, 44 * npx, ms_t))
del temp, dL

# ---------------- activations ----------------
pa = GaussianParamArena(P, M, DEV)
pa.param.normal_()
arena = GradArena(P, M, DEV)
arena.flat.normal_()
ms_af = timed(lambda: pa.activate())
g = arena.views
ms_ab = timed(lambda: _C.activate_backward(pa._scaling, pa._rotation, pa._opacity, g["dL_dscales"], g["dL_drotations"], g["dL_dopacity"]))

leaves = [pa._scaling.clone().requires_grad_(True), pa._rotation.clone().requires_grad_(True), pa._opacity.clone().requires_grad_(True)]
f_dc = pa._features[:, :1].clone().requires_grad_(True)
f_rest = pa._features[:, 1:].clone().requires_grad_(True)
gs, gq, go, gsh = torch.randn(P, 3, device=DEV), torch.randn(P, 4, device=DEV), torch.randn(P, 1, device=DEV), torch.randn(P, M, 3, device=DEV)


def torch_getters_fb():
    for t in leaves + [f_dc, f_rest]:
        t.grad = None
    s, q, o = torch.exp(leaves[0]), F.normalize(leaves[1]), torch.sigmoid(leaves[2])
    sh = torch.cat((f_dc, f_rest), dim=1)
    torch.autograd.backward([s, q, o, sh], [gs, gq, go, gsh])


ms_tg = timed(torch_getters_fb)
out["rows"].append(row("activations forward (1 kernel)", ms_af, 64 * P))
out["rows"].append(row("activations backward in place (1 kernel)", ms_ab, 96 * P))
out["rows"].append(row("getters fwd+bwd incl. SH cat/split (ours: no cat, SH is one tensor)", ms_af + ms_ab, 160 * P, ms_tg))
del leaves, f_dc, f_rest, gs, gq, go, gsh

# ---------------- Adam ----------------
lrs = dict(xyz=0.00016, f_dc=0.0025, f_rest=0.0025 / 20, opacity=0.05, scaling=0.005, rotation=0.001)
arena.flat.normal_().mul_(1e-3)


def ours_adam():
    pa.step_count += 1
    segs = []
    for pname, gname, kw in (("_xyz", "dL_dmeans3D", dict(lr=lrs["xyz"])),
                             ("_features", "dL_dsh", dict(lr=lrs["f_dc"], lr_rest=lrs["f_rest"], row_len=3 * M, row_split=3)),
                             ("_opacity", "dL_dopacity", dict(lr=lrs["opacity"])),
                             ("_scaling", "dL_dscales", dict(lr=lrs["scaling"])),
                             ("_rotation", "dL_drotations", dict(lr=lrs["rotation"]))):
        m, v = pa.moments(pname)
        segs.append(dict(param=getattr(pa, pname), grad=g[gname], exp_avg=m, exp_avg_sq=v, **kw))
    _C.adam_step(segs, pa.step_count)


ms_adam = timed(ours_adam)
n_par = P * (3 + 3 * M + 1 + 3 + 4)
params = {"xyz": torch.randn(P, 3, device=DEV), "f_dc": torch.randn(P, 1, 3, device=DEV), "f_rest": torch.randn(P, M - 1, 3, device=DEV),
          "opacity": torch.randn(P, 1, device=DEV), "scaling": torch.randn(P, 3, device=DEV), "rotation": torch.randn(P, 4, device=DEV)}
params = {k: torch.nn.Parameter(v) for k, v in params.items()}
opt = torch.optim.Adam([{"params": [params[k]], "lr": lrs[k], "name": k} for k in params], lr=0.0, eps=1e-15)
for k, p in params.items():
    p.grad = torch.randn_like(p) * 1e-3
ms_tadam = timed(lambda: opt.step())
out["rows"].append(row("Adam, one launch over the arena (6 groups)", ms_adam, 28 * n_par, ms_tadam))
out["n_params"] = n_par
out["kernel_launches_total"] = int(_C._lib.gsr_kernel_launches())
print(json.dumps(out, indent=1))
