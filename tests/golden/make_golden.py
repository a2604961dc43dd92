"""Generates tests/golden/reference_fragments.npz by IMPORTING the reference's own in-tree Python
(run in the build container, where /root/reference exists; the .npz is committed and travels).

What the reference tree can pin for the rasterizer path:
  * SH -> RGB           utils/sh_utils.py:57-112 `eval_sh` + the `+0.5, clamp_min 0` of
                        gaussian_renderer/__init__.py:73-78 (the --convert_SHs_python path)
  * 3D covariance       utils/general_utils.py:66-112 `build_scaling_rotation`, `strip_symmetric`
                        as composed in scene/gaussian_model.py:27-31 (--compute_cov3D_python path)
  * camera matrices     utils/graphics_utils.py:38-70 composed as scene/cameras.py:60-63
The rasterizer itself (diff_gaussian_rasterization) is NOT in the reference tree, so nothing else
can be generated: "parity unpinned" beyond these fragments.

usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/gs-simp"
sys.path.insert(0, REF)

# general_utils hard-codes device="cuda" (utils/general_utils.py:67,85,104); this container has no
# GPU, so route those allocations to the CPU.  The arithmetic is untouched.
_zeros = torch.zeros
torch.zeros = lambda *a, **k: _zeros(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})

from utils.sh_utils import eval_sh  # noqa: E402
from utils.graphics_utils import getProjectionMatrix, getWorld2View2, focal2fov  # noqa: E402
from utils.general_utils import build_scaling_rotation, strip_symmetric  # noqa: E402

g = torch.Generator().manual_seed(20240117)
P = 96
out = {}

# ---- SH -> RGB (render(): lines 73-78) ----
xyz = torch.randn(P, 3, generator=g) * 3.0
campos = torch.tensor([0.3, -0.2, 0.1])
features = torch.cat([0.5 * torch.randn(P, 1, 3, generator=g), 0.3 * torch.randn(P, 15, 3, generator=g)], 1)
out["sh_xyz"], out["sh_campos"], out["sh_features"] = xyz.numpy(), campos.numpy(), features.numpy()
for deg in range(4):
    shs_view = features.transpose(1, 2).view(-1, 3, 16)
    dir_pp = xyz - campos.repeat(features.shape[0], 1)
    dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
    sh2rgb = eval_sh(deg, shs_view, dir_pp_normalized)
    out[f"sh_rgb_deg{deg}"] = torch.clamp_min(sh2rgb + 0.5, 0.0).numpy()

# ---- cov3D (gaussian_model.py:27-31) ----
scaling = torch.exp(torch.randn(P, 3, generator=g) * 0.7 - 2.0)
rotation = torch.randn(P, 4, generator=g)  # build_rotation normalises internally (general_utils.py:80-83)
out["cov_scaling"], out["cov_rotation"] = scaling.numpy(), rotation.numpy()
for mod in (1.0, 0.7):
    L = build_scaling_rotation(mod * scaling, rotation)
    out[f"cov3D_mod{mod}"] = strip_symmetric(L @ L.transpose(1, 2)).numpy()

# ---- camera matrices (cameras.py:54-63) ----
cams = []
for k in range(4):
    A = torch.randn(3, 3, generator=g).numpy().astype(np.float64)
    Q, _ = np.linalg.qr(A)
    if np.linalg.det(Q) < 0:
        Q[:, 0] *= -1
    T = torch.randn(3, generator=g).numpy().astype(np.float64)
    W_, H_ = [(256, 256), (1296, 928), (1024, 576), (384, 512)][k]
    focal = 1.1 * W_
    fovx, fovy = focal2fov(focal, W_), focal2fov(focal, H_)
    wvt = torch.tensor(getWorld2View2(Q, T, np.array([0.0, 0.0, 0.0]), 1.0)).transpose(0, 1)
    proj = getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fovx, fovY=fovy).transpose(0, 1)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    center = wvt.inverse()[3, :3]
    out[f"cam{k}_R"], out[f"cam{k}_T"] = Q, T
    out[f"cam{k}_WH"] = np.array([W_, H_])
    out[f"cam{k}_fov"] = np.array([fovx, fovy])
    out[f"cam{k}_world_view_transform"] = wvt.numpy()
    out[f"cam{k}_full_proj_transform"] = full.numpy()
    out[f"cam{k}_camera_center"] = center.numpy()

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fragments.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, {k: v.shape for k, v in out.items() if k.startswith("sh_rgb")})
