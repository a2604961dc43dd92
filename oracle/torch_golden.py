"""Differentiable torch-CPU golden of the same math (BASELINE.json: "a torch-CPU golden of the same
preprocess/sort/blend math"), used ONLY by tests to validate the hand-derived backward of the C
oracle and of the CUDA kernels through torch.autograd.   TEST INFRASTRUCTURE -- parity unpinned.

It restates SURVEY.md Appendix A.2/A.5 with torch ops (float64 by default) and reproduces the
public rasterizer's *gradient conventions* where they differ from the exact derivative:
  * 0.99 alpha clamp and the 1/255, 1e-4 thresholds are pass-through / non-differentiable (A.6);
  * the +-1.3 tan(fov) frustum clamp zeroes d/dt.x (x_grad_mul) and has no path to t.z (A.7);
  * dL/dmeans2D is in NDC-scaled units: d pix / d means2D = (0.5 W, 0.5 H) (A.6);
  * no gradient flows through depth (section 0.3).
The integer stages (binning/sort/ranges) are taken from the C oracle -- they are not differentiable.
In-tree corroboration: SH polynomial utils/sh_utils.py:57-112; R(q) utils/general_utils.py:80-101;
Sigma = L L^T scene/gaussian_model.py:27-31.
"""
from __future__ import annotations

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def sh_to_rgb(deg, sh, dirs):
    """sh (P,M,3), dirs (P,3) unit -> (P,3) before the +0.5."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = C0 * sh[:, 0]
    if deg > 0:
        res = res - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2 * zz - xx - yy) * sh[:, 6]
                   + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = (res + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10]
                       + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
                       + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14]
                       + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return res


def rotation(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1)
    return R.view(-1, 3, 3)


def preprocess(means3D, means2D_dummy, opacities, scales, rotations, shs, sh_degree, viewmatrix,
               projmatrix, campos, W, H, tanfovx, tanfovy, scale_modifier=1.0, colors_precomp=None,
               cov3D_precomp=None):
    """Differentiable part of A.2: returns pix (P,2), conic (P,3), opacity (P,), rgb (P,3), depth (P,)."""
    dt = means3D.dtype
    V = viewmatrix.to(dt)          # column-major W2C: p_view = p @ V[:3,:3] + V[3,:3]
    PM = projmatrix.to(dt)
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=dt)
    hom = torch.cat([means3D, ones], 1)
    pv = hom @ V                    # (P,4); cameras.py:60 storage => right-multiply
    ph = hom @ PM
    pw = 1.0 / (ph[:, 3:4] + 1e-7)
    pproj = ph[:, :3] * pw
    if cov3D_precomp is None:
        R = rotation(rotations)
        L = R * (scale_modifier * scales).unsqueeze(1)      # R @ diag(s)
        Sigma = L @ L.transpose(1, 2)
    else:
        c = cov3D_precomp
        Sigma = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], 1).view(-1, 3, 3)
    fx, fy = W / (2 * tanfovx), H / (2 * tanfovy)
    tz = pv[:, 2]
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    txtz, tytz = pv[:, 0] / tz, pv[:, 1] / tz
    cx = (txtz < -limx) | (txtz > limx)
    cy = (tytz < -limy) | (tytz > limy)
    tx = torch.where(cx, (txtz.clamp(-limx, limx) * tz).detach(), pv[:, 0])
    ty = torch.where(cy, (tytz.clamp(-limy, limy) * tz).detach(), pv[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz), zero, fy / tz, -(fy * ty) / (tz * tz)], 1).view(-1, 2, 3)
    Rw = V[:3, :3].t()              # R_w2c[r][c] = V[c][r]
    T = J @ Rw
    cov = T @ Sigma @ T.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    conic = torch.stack([c / det, -b / det, a / det], 1)
    scale_ndc = torch.tensor([0.5 * W, 0.5 * H], dtype=dt)
    pix = torch.stack([((pproj[:, 0] + 1.0) * W - 1.0) * 0.5, ((pproj[:, 1] + 1.0) * H - 1.0) * 0.5], 1)
    pix = pix + means2D_dummy[:, :2] * scale_ndc
    if colors_precomp is None:
        d = means3D - campos.to(dt)
        d = d / d.norm(dim=1, keepdim=True)
        rgb = sh_to_rgb(sh_degree, shs, d) + 0.5
        rgb = torch.clamp_min(rgb, 0.0)          # gaussian_renderer/__init__.py:78
    else:
        rgb = colors_precomp
    return pix, conic, opacities.reshape(-1), rgb, tz


def blend(pix, conic, opac, rgb, bg, ranges, point_list, W, H):
    """A.5 with autograd; ranges (G,2) / point_list (N,) come from the C oracle."""
    dt = pix.dtype
    gx, gy = (W + 15) // 16, (H + 15) // 16
    out = torch.zeros(3, H, W, dtype=dt)
    out = out + bg.to(dt).view(3, 1, 1)
    tiles = []
    for tile in range(gx * gy):
        r0, r1 = int(ranges[tile, 0]), int(ranges[tile, 1])
        if r1 <= r0:
            continue
        bx, by = tile % gx, tile // gx
        ys = torch.arange(by * 16, min(by * 16 + 16, H))
        xs = torch.arange(bx * 16, min(bx * 16 + 16, W))
        py, px = torch.meshgrid(ys, xs, indexing="ij")
        px, py = px.reshape(-1, 1).to(dt), py.reshape(-1, 1).to(dt)
        ids = torch.as_tensor(point_list[r0:r1].astype("int64"))
        dx = pix[ids, 0].unsqueeze(0) - px
        dy = pix[ids, 1].unsqueeze(0) - py
        cn = conic[ids]
        power = -0.5 * (cn[:, 0] * dx * dx + cn[:, 2] * dy * dy) - cn[:, 1] * dx * dy
        a_raw = opac[ids].unsqueeze(0) * torch.exp(power)
        alpha = a_raw + (torch.clamp_max(a_raw, 0.99) - a_raw).detach()   # pass-through clamp
        keep = (power <= 0) & (alpha.detach() >= 1.0 / 255.0)
        alpha = torch.where(keep, alpha, torch.zeros_like(alpha))
        test_T = torch.cumprod(1 - alpha, dim=1)
        stop = (keep & (test_T.detach() < 1e-4)).to(torch.int64).cumsum(1) > 0
        alpha = torch.where(stop, torch.zeros_like(alpha), alpha)
        T_incl = torch.cumprod(1 - alpha, dim=1)
        T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], 1)
        wgt = alpha * T_excl
        col = wgt @ rgb[ids] + T_incl[:, -1:] * bg.to(dt).view(1, 3)
        tiles.append((ys, xs, col))
    # assemble without in-place ops on a leaf
    img = torch.zeros(3, H, W, dtype=dt) + bg.to(dt).view(3, 1, 1)
    pieces = []
    for ys, xs, col in tiles:
        pieces.append((ys, xs, col.t().reshape(3, len(ys), len(xs))))
    rows = []
    idx = {(int(ys[0]) // 16, int(xs[0]) // 16): blk for ys, xs, blk in pieces}
    for by in range(gy):
        row = []
        for bx in range(gx):
            hh, ww = min(16, H - by * 16), min(16, W - bx * 16)
            row.append(idx.get((by, bx), img[:, by * 16:by * 16 + hh, bx * 16:bx * 16 + ww]))
        rows.append(torch.cat(row, 2))
    return torch.cat(rows, 1)
