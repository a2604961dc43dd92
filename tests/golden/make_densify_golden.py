"""Generates tests/golden/densify.npz by RUNNING the reference's own GaussianModel densification code on seeded
inputs (build container only: /root/reference must exist; the .npz is committed and travels to the GPU box).

Executed reference code (gs-simp/scene/gaussian_model.py, loaded from its file, unmodified):
    training_setup            :149-167   six Adam groups, eps 1e-15
    densify_and_prune         :467-480   -> densify_and_clone :451-465, densify_and_split :427-449,
                                            densification_postfix :408-425, cat_tensors_to_optimizer :385-406,
                                            prune_points :365-383, _prune_optimizer :346-363
    reset_opacity             :263-266   -> replace_tensor_to_optimizer :331-344
The module hard-codes device="cuda"; this container has no GPU, so `torch.zeros` is wrapped to drop the device
keyword (values are unaffected).  `plyfile` and `simple_knn` (imported at :18,:20, not used by these methods) are
stubbed.  torch.normal (:438) draws from the global CPU generator: each case records the seed set right before
densify_and_prune so a test can reproduce the same samples with the same torch call.

usage: python tests/golden/make_densify_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/gs-simp"
sys.path.insert(0, REF)
sys.modules["plyfile"] = types.SimpleNamespace(PlyData=None, PlyElement=None)
sys.modules["simple_knn"] = types.ModuleType("simple_knn")
sys.modules["simple_knn._C"] = types.SimpleNamespace(distCUDA2=None)

_zeros = torch.zeros


def _zeros_cpu(*a, **k):
    k.pop("device", None)
    return _zeros(*a, **k)


torch.zeros = _zeros_cpu
torch.cuda.empty_cache = lambda: None

spec = importlib.util.spec_from_file_location("ref_gaussian_model", os.path.join(REF, "scene", "gaussian_model.py"))
gm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gm)

torch.set_num_threads(1)
GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
ATTR = dict(xyz="_xyz", f_dc="_features_dc", f_rest="_features_rest", opacity="_opacity", scaling="_scaling",
            rotation="_rotation")
out = {}


def make_model(P, sh_degree, seed):
    torch.manual_seed(seed)
    M = (sh_degree + 1) ** 2
    g = gm.GaussianModel(sh_degree)
    mk = lambda t: torch.nn.Parameter(t.contiguous().requires_grad_(True))
    g._xyz = mk(torch.randn(P, 3) * 2.0)
    g._features_dc = mk(torch.randn(P, 1, 3))
    g._features_rest = mk(torch.randn(P, M - 1, 3) * 0.2)
    g._opacity = mk(torch.randn(P, 1) * 2.5 - 1.0)
    sc = torch.randn(P, 3) * 0.8 - 3.0
    sc[::97] += 3.0                      # a few world-space giants (big_points_ws)
    g._scaling = mk(sc)
    g._rotation = mk(torch.randn(P, 4))
    g.max_radii2D = torch.zeros(P)
    args = types.SimpleNamespace(percent_dense=0.01, position_lr_init=0.00016, position_lr_final=0.0000016,
                                 position_lr_delay_mult=0.01, position_lr_max_steps=30000, feature_lr=0.0025,
                                 opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001)
    g.spatial_lr_scale = 1.0
    g.training_setup(args)
    # two optimizer steps on seeded gradients: non-trivial exp_avg / exp_avg_sq
    for _ in range(2):
        for grp in g.optimizer.param_groups:
            p = grp["params"][0]
            p.grad = torch.randn_like(p) * 1e-3
        g.optimizer.step()
    # densification statistics as train.py:115-116 leaves them
    denom = torch.randint(0, 6, (P, 1)).float()         # zeros -> NaN grads -> 0 (:468-469)
    g.denom = denom
    g.xyz_gradient_accum = torch.rand(P, 1) * 0.0006 * denom
    g.max_radii2D = torch.rand(P) * 40.0
    return g


def dump(prefix, g):
    for name in GROUPS:
        p = getattr(g, ATTR[name])
        out[f"{prefix}_{name}"] = p.detach().numpy().copy()
        st = g.optimizer.state.get(p, None)
        assert st is not None
        out[f"{prefix}_{name}_exp_avg"] = st["exp_avg"].numpy().copy()
        out[f"{prefix}_{name}_exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
    out[f"{prefix}_xyz_gradient_accum"] = g.xyz_gradient_accum.numpy().copy()
    out[f"{prefix}_denom"] = g.denom.numpy().copy()
    out[f"{prefix}_max_radii2D"] = g.max_radii2D.numpy().copy()


# name: (P, sh_degree, model seed, normal seed, max_grad, min_opacity, extent, max_screen_size)
CASES = {
    "a": (700, 1, 11, 1001, 0.0002, 0.005, 5.0, 20),
    "b": (333, 0, 12, 1002, 0.0002, 0.005, 5.0, None),
    "c": (300, 3, 13, 1003, 0.0003, 0.05, 8.0, 20),
    "d": (64, 1, 14, 1004, 1.0, 0.0, 5.0, None),        # nothing cloned, split or pruned
}
for name, (P, deg, seed, nseed, max_grad, min_op, extent, mss) in CASES.items():
    g = make_model(P, deg, seed)
    dump(f"{name}_in", g)
    out[f"{name}_args"] = np.array([max_grad, min_op, extent, -1.0 if mss is None else mss, g.percent_dense, nseed],
                                   dtype=np.float64)
    torch.manual_seed(nseed)
    g.densify_and_prune(max_grad, min_op, extent, mss)
    dump(f"{name}_out", g)
    print(name, "P", P, "->", g.get_xyz.shape[0])
    # reset_opacity on the densified model (train.py:122-123)
    g.reset_opacity()
    out[f"{name}_reset_opacity"] = g._opacity.detach().numpy().copy()
    st = g.optimizer.state[g._opacity]
    assert float(st["exp_avg"].abs().max()) == 0.0 and float(st["exp_avg_sq"].abs().max()) == 0.0

# prune_points alone on an arbitrary mask (gaussian_model.py:365-383)
g = make_model(200, 1, 21)
dump("p_in", g)
torch.manual_seed(5)
mask = torch.rand(200) < 0.3
out["p_mask"] = mask.numpy().copy()
g.prune_points(mask)
dump("p_out", g)

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "densify.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes")
