"""Build-container-only fuzz (needs /root/reference; not collected by pytest): 300 random models, thresholds and seeds --
including empty results, no-ops, max_grad = 0, min_opacity > 1 -- through the REFERENCE's own GaussianModel.densify_and_prune
(gs-simp/scene/gaussian_model.py:467-480, loaded unmodified, device redirected to the CPU) and through
densify.plan_densify_and_prune; parameters and both Adam moments must agree bit for bit.
Last run (round 1): cases 300, mismatches 0.   usage: python tests/golden/fuzz_densify_vs_reference.py"""
import importlib.util, os, sys, types, numpy as np, torch
sys.path.insert(0, "/root/repo")
REF = "/root/reference/gs-simp"; sys.path.insert(0, REF)
sys.modules["plyfile"] = types.SimpleNamespace(PlyData=None, PlyElement=None)
sys.modules["simple_knn"] = types.ModuleType("simple_knn"); sys.modules["simple_knn._C"] = types.SimpleNamespace(distCUDA2=None)
_z = torch.zeros
torch.zeros = lambda *a, **k: _z(*a, **{kk: v for kk, v in k.items() if kk != "device"})
torch.cuda.empty_cache = lambda: None
spec = importlib.util.spec_from_file_location("ref_gm", os.path.join(REF, "scene", "gaussian_model.py"))
gm = importlib.util.module_from_spec(spec); spec.loader.exec_module(gm)
from multiview_inpaint_b200 import densify
torch.set_num_threads(1)
GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
ATTR = dict(xyz="_xyz", f_dc="_features_dc", f_rest="_features_rest", opacity="_opacity", scaling="_scaling", rotation="_rotation")
rng = np.random.default_rng(0)
def snapshot(g):
    out = {}
    for n in GROUPS:
        p = getattr(g, ATTR[n]); st = g.optimizer.state[p]
        out[n] = p.detach().clone(); out[n+"_m"] = st["exp_avg"].clone(); out[n+"_v"] = st["exp_avg_sq"].clone()
    return out
bad = 0; total = 0; stats = []
for it in range(300):
    P = int(rng.integers(1, 400)); deg = int(rng.integers(0, 4)); M = (deg+1)**2
    torch.manual_seed(int(rng.integers(1 << 30)))
    g = gm.GaussianModel(deg)
    mk = lambda t: torch.nn.Parameter(t.contiguous().requires_grad_(True))
    g._xyz = mk(torch.randn(P, 3) * 2); g._features_dc = mk(torch.randn(P, 1, 3)); g._features_rest = mk(torch.randn(P, M-1, 3) * .2)
    g._opacity = mk(torch.randn(P, 1) * 2.5 - 1); g._scaling = mk(torch.randn(P, 3) * rng.uniform(.2, 1.5) - rng.uniform(1, 4)); g._rotation = mk(torch.randn(P, 4))
    g.max_radii2D = torch.zeros(P); g.spatial_lr_scale = 1.0
    args = types.SimpleNamespace(percent_dense=float(rng.choice([0.01, 0.001, 0.1])), position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
                                 position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001)
    g.training_setup(args)
    for grp in g.optimizer.param_groups:
        p = grp["params"][0]; p.grad = torch.randn_like(p) * 1e-3
    g.optimizer.step()
    denom = torch.randint(0, 4, (P, 1)).float(); g.denom = denom
    g.xyz_gradient_accum = torch.rand(P, 1) * 0.0006 * denom; g.max_radii2D = torch.rand(P) * 40
    max_grad = float(rng.choice([0.0002, 0.0, 0.0004, 1.0, 0.0001])); min_op = float(rng.choice([0.005, 0.0, 0.05, 0.5, 1.1]))
    extent = float(rng.choice([5.0, 0.5, 50.0, 1e-3])); mss = rng.choice([None, 20, 0, 5])
    mss = None if mss is None else int(mss)
    before = snapshot(g); accum, den = g.xyz_gradient_accum.clone(), g.denom.clone()
    seed = int(rng.integers(1 << 30))
    torch.manual_seed(seed); g.densify_and_prune(max_grad, min_op, extent, mss); after = snapshot(g)
    torch.manual_seed(seed)
    plan = densify.plan_densify_and_prune(before["xyz"], before["scaling"], before["rotation"], before["opacity"], accum, den,
                                          max_grad, min_op, extent, mss, percent_dense=args.percent_dense)
    idx = plan.src_row.long(); ok = plan.n_dst == after["xyz"].shape[0]
    if ok:
        for n in GROUPS:
            got = before[n][idx].clone()
            if n == "xyz": got[plan.child_rows] = plan.child_xyz
            if n == "scaling": got[plan.child_rows] = plan.child_scaling
            ok &= torch.equal(got, after[n])
            for s in ("_m", "_v"):
                v = before[n+s][idx].clone(); v[plan.n_keep_state:] = 0
                ok &= torch.equal(v, after[n+s])
    total += 1; bad += (not ok); stats.append((P, plan.n_dst, plan.counts["cloned"], plan.counts["split"], plan.counts["pruned"]))
    if not ok: print("MISMATCH", it, P, deg, max_grad, min_op, extent, mss, plan.counts, after["xyz"].shape[0])
s = np.array(stats)
print("cases", total, "mismatches", bad, "| P->P' examples", stats[:5], "| empty results", int((s[:,1]==0).sum()), "no-op", int(((s[:,2]==0)&(s[:,3]==0)&(s[:,4]==0)).sum()))
