"""Shared helpers: run the same seeded scene through the CPU oracle and through the CUDA C ABI."""
from __future__ import annotations

import numpy as np
import torch

from multiview_inpaint_b200 import scenes as S


def oracle_forward(O, sc, bg=None, cam=None, use_cov3D=False, use_colors=False, scale_modifier=1.0):
    cam = sc["camera"] if cam is None else cam
    bg = sc["bg"] if bg is None else bg
    kw = dict(means3D=sc["means3D"].numpy(), opacities=sc["opacities"].numpy(),
              viewmatrix=cam.world_view_transform.numpy(), projmatrix=cam.full_proj_transform.numpy(),
              campos=cam.camera_center.contiguous().numpy(), W=cam.image_width, H=cam.image_height,
              tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, sh_degree=sc["sh_degree"],
              scale_modifier=scale_modifier)
    if use_colors:
        kw["colors_precomp"] = sc["colors_precomp"].numpy()
    else:
        kw["shs"] = sc["shs"].numpy()
    if use_cov3D:
        kw["cov3D_precomp"] = sc["cov3D_precomp"].numpy()
    else:
        kw["scales"], kw["rotations"] = sc["scales"].numpy(), sc["rotations"].numpy()
    return O.forward(np.asarray(bg, dtype=np.float32), **kw)


def cuda_forward(sc, bg=None, cam=None, flags=0, use_cov3D=False, use_colors=False, scale_modifier=1.0,
                 prefiltered=False, device="cuda", **extra):
    """-> (num_rendered, color, radii, geom, binning, img, depth), tensors dict on device."""
    from multiview_inpaint_b200 import _C
    cam = (sc["camera"] if cam is None else cam).to(device)
    bg = (sc["bg"] if bg is None else torch.as_tensor(bg, dtype=torch.float32)).to(device)
    d = {k: v.to(device) for k, v in sc.items() if isinstance(v, torch.Tensor)}
    e = torch.empty(0, device=device)
    out = _C.rasterize_gaussians(
        bg, d["means3D"], d["colors_precomp"] if use_colors else e, d["opacities"],
        e if use_cov3D else d["scales"], e if use_cov3D else d["rotations"], scale_modifier,
        d["cov3D_precomp"] if use_cov3D else e, cam.world_view_transform, cam.full_proj_transform,
        cam.tanfovx, cam.tanfovy, cam.image_height, cam.image_width, e if use_colors else d["shs"],
        sc["sh_degree"], cam.camera_center, prefiltered, flags=flags, **extra)
    return out, d, cam, bg


def cuda_backward(out, d, cam, bg, sc, dL_dcolor, flags=0, use_cov3D=False, use_colors=False,
                  scale_modifier=1.0):
    from multiview_inpaint_b200 import _C
    n, color, radii, geom, binning, img, depth = out
    e = torch.empty(0, device=color.device)
    g = _C.rasterize_gaussians_backward(
        bg, d["means3D"], radii, d["colors_precomp"] if use_colors else e,
        e if use_cov3D else d["scales"], e if use_cov3D else d["rotations"], scale_modifier,
        d["cov3D_precomp"] if use_cov3D else e, cam.world_view_transform, cam.full_proj_transform,
        cam.tanfovx, cam.tanfovy, dL_dcolor.to(color.device), e if use_colors else d["shs"],
        sc["sh_degree"], cam.camera_center, geom, n, binning, img, flags=flags, return_conic=True)
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh",
             "dL_dscales", "dL_drotations", "dL_dconic"]
    return dict(zip(names, g))


def rel_err(a, b):
    """max |a-b| / max |b|  (gradient tolerance of BASELINE.json is relative: atomics reorder sums)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def small_scene(P=3000, W=96, H=80, deg=3, seed=11, radius_px=6.0, **kw):
    return S.make_scene(P, W, H, deg, seed, mu_s=S.default_mu_s(W, radius_px), **kw)
