"""Times the densification re-pack at the headline model size (3 M Gaussians, SH degree 3) on one GPU:
  fused            GaussianParamArena.densify_and_prune: index plan (P-sized torch ops) + ONE gsr_gather_rows launch
  gather           the gather launch alone (CUDA events), against its algorithmic bytes
  torch_structure  the reference's four rounds (gaussian_model.py:467-480): torch.cat of the clones, torch.cat of the
                   split children, mask-prune of the parents, mask-prune by opacity / size, each over six parameter
                   tensors and their two Adam moments
usage: python tools/exp_densify.py [P] > gpurun_out/densify.json"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from multiview_inpaint_b200 import _C, densify  # noqa: E402
from multiview_inpaint_b200.multiview import GradArena  # noqa: E402
from multiview_inpaint_b200.trainstep import GaussianParamArena  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
M, dev = 16, "cuda"
ARGS = dict(max_grad=0.0002, min_opacity=0.005, extent=5.0, max_screen_size=20)


def make():
    g = torch.Generator(device=dev).manual_seed(1)
    pa = GaussianParamArena(P, M, dev)
    pa.param.normal_(generator=g)
    pa._scaling.mul_(0.8).sub_(3.0)
    pa._opacity.mul_(2.5).sub_(1.0)
    pa.exp_avg.normal_(generator=g)
    pa.exp_avg_sq.uniform_(generator=g)
    st = GradArena(P, M, dev)
    st.visible_count.copy_(torch.randint(0, 6, (P,), device=dev, generator=g, dtype=torch.int32))
    st.grad_norm_accum.copy_(torch.rand(P, device=dev, generator=g) * 0.0006 * st.visible_count)
    return pa, st


def ev():
    return torch.cuda.Event(enable_timing=True)


out = {"P": P, "M": M}
# ---- fused ----
for rep in range(2):                          # first pass warms the allocator
    pa, st = make()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    t0 = time.perf_counter()
    e0.record()
    plan = pa.densify_and_prune(st, **ARGS)
    e1.record()
    torch.cuda.synchronize()
    out["fused_ms"] = e0.elapsed_time(e1)
    out["fused_wall_ms"] = (time.perf_counter() - t0) * 1e3
out["counts"] = plan.counts
out["P_new"] = pa.P
# ---- the gather alone ----
pa, st = make()
plan = densify.plan_densify_and_prune(pa._xyz, pa._scaling, pa._rotation, pa._opacity, st.grad_norm_accum, st.visible_count,
                                      ARGS["max_grad"], ARGS["min_opacity"], ARGS["extent"], ARGS["max_screen_size"])
new = GaussianParamArena(plan.n_dst, M, dev)
segs = []
for name in ("_xyz", "_features", "_opacity", "_scaling", "_rotation"):
    segs.append(dict(src=getattr(pa, name), dst=getattr(new, name)))
    for a, b in zip(pa.moments(name), new.moments(name)):
        segs.append(dict(src=a, dst=b, zero_new=True))
for _ in range(3):
    _C.gather_rows(plan.src_row, P, segs, n_keep_state=plan.n_keep_state)
e0, e1 = ev(), ev()
e0.record()
for _ in range(10):
    _C.gather_rows(plan.src_row, P, segs, n_keep_state=plan.n_keep_state)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
row_b = 59 * 4
# read: index + parameter row for every row, moment rows only for rows that keep state; write: 3 rows each
alg = plan.n_dst * (4 + row_b + 3 * row_b) + plan.n_keep_state * 2 * row_b
out["gather"] = {"ms": ms, "alg_bytes": alg, "gbps": alg / ms / 1e6, "rows": plan.n_dst, "keep_state": plan.n_keep_state}
del new, segs


# ---- reference structure: four rounds over 6 tensors x {param, exp_avg, exp_avg_sq} ----
def torch_structure(pa, st):
    names = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation")
    T = {n: getattr(pa, n).contiguous().clone() for n in names}
    m1 = {n: torch.randn_like(T[n]) for n in names}
    m2 = {n: torch.rand_like(T[n]) for n in names}
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    grads = (st.grad_norm_accum / st.visible_count.float()).reshape(-1, 1)
    grads[grads.isnan()] = 0.0
    thr = 0.01 * ARGS["extent"]

    def cat(ext):
        for n in names:
            T[n] = torch.cat((T[n], ext[n]), dim=0)
            m1[n] = torch.cat((m1[n], torch.zeros_like(ext[n])), dim=0)
            m2[n] = torch.cat((m2[n], torch.zeros_like(ext[n])), dim=0)

    def prune(valid):
        for n in names:
            T[n], m1[n], m2[n] = T[n][valid], m1[n][valid], m2[n][valid]

    sel = (torch.norm(grads, dim=-1) >= ARGS["max_grad"]) & (torch.exp(T["_scaling"]).max(dim=1).values <= thr)
    cat({n: T[n][sel] for n in names})
    n1 = T["_xyz"].shape[0]
    padded = torch.zeros(n1, device=dev)
    padded[:grads.shape[0]] = grads.squeeze()
    sel = (padded >= ARGS["max_grad"]) & (torch.exp(T["_scaling"]).max(dim=1).values > thr)
    sc = torch.exp(T["_scaling"][sel]).repeat(2, 1)
    samples = torch.normal(mean=torch.zeros_like(sc), std=sc)
    rots = densify.build_rotation(T["_rotation"][sel]).repeat(2, 1, 1)
    ext = {n: T[n][sel].repeat(2, *([1] * (T[n].dim() - 1))) for n in names}
    ext["_xyz"] = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + ext["_xyz"]
    ext["_scaling"] = torch.log(sc / 1.6)
    cat(ext)
    prune(~torch.cat((sel, torch.zeros(2 * int(sel.sum()), device=dev, dtype=torch.bool))))
    pm = (torch.sigmoid(T["_opacity"]) < ARGS["min_opacity"]).squeeze()
    pm = pm | (torch.exp(T["_scaling"]).max(dim=1).values > 0.1 * ARGS["extent"])
    prune(~pm)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), T["_xyz"].shape[0]


pa, st = make()
for _ in range(2):
    ms_t, p_t = torch_structure(pa, st)
out["torch_structure_ms"] = ms_t
out["torch_structure_P_new"] = p_t
out["speedup"] = ms_t / out["fused_ms"]
print(json.dumps(out))
