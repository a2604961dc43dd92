"""Generates tests/golden/init_from_pcd.npz by RUNNING the reference's own GaussianModel.create_from_pcd
(gs-simp/scene/gaussian_model.py:124-147) on a seeded point cloud.  Build container only (no GPU): `.cuda()` is made
a no-op and `torch.zeros/ones(device="cuda")` redirected to the CPU; `distCUDA2` (third-party simple-knn, called at
:134) is replaced by this repository's brute-force oracle of it (oracle.knn3_mean_dist2 -- the CUDA kernel is
bit-identical to that oracle, tests/test_knn_gpu.py); `plyfile` is stubbed.  The dist2 vector is stored too, so the
test can hand the very same neighbour distances to the code under test.

usage: python tests/golden/make_init_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/gs-simp"
sys.path.insert(0, REF)
sys.modules["plyfile"] = types.SimpleNamespace(PlyData=None, PlyElement=None)
captured = {}


def dist_cpu(points):
    d = torch.from_numpy(O.knn3_mean_dist2(points.numpy()))
    captured["dist2"] = d.clone()
    return d


sys.modules["simple_knn"] = types.ModuleType("simple_knn")
sys.modules["simple_knn._C"] = types.SimpleNamespace(distCUDA2=dist_cpu)
torch.Tensor.cuda = lambda self, *a, **k: self
for name in ("zeros", "ones"):
    orig = getattr(torch, name)
    setattr(torch, name, (lambda o: lambda *a, **k: o(*a, **{kk: v for kk, v in k.items() if kk != "device"}))(orig))

spec = importlib.util.spec_from_file_location("ref_gaussian_model", os.path.join(REF, "scene", "gaussian_model.py"))
gm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gm)
from utils.graphics_utils import BasicPointCloud  # noqa: E402

O.build()
out = {}
for name, (n, deg, seed) in {"a": (500, 1, 5), "b": (64, 3, 6), "c": (33, 0, 7)}.items():
    rng = np.random.default_rng(seed)
    pts = rng.normal(size=(n, 3)) * np.array([2.0, 1.0, 0.5])
    pts[1] = pts[0]                                   # a coincident pair: one of the three neighbour distances is 0
    cols = rng.uniform(size=(n, 3))
    g = gm.GaussianModel(deg)
    g.create_from_pcd(BasicPointCloud(points=pts, colors=cols, normals=np.zeros_like(pts)), 3.5)
    out[f"{name}_points"], out[f"{name}_colors"], out[f"{name}_deg"] = pts, cols, np.array(deg)
    out[f"{name}_dist2"] = captured["dist2"].numpy()
    for k in ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity"):
        out[f"{name}{k}"] = getattr(g, k).detach().numpy().copy()
    assert g.spatial_lr_scale == 3.5 and g.max_radii2D.shape == (n,)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "init_from_pcd.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst))
