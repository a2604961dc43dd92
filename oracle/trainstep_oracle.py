"""CPU oracle (numpy) of the training-step stages either side of the rasterizer.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; the product package never does.

Unlike the rasterizer's oracle, THIS oracle is PINNED: the algorithms it restates are in the reference tree (or in
torch, which the reference calls directly), so tests/golden/make_trainstep_golden.py imports the reference's own
`utils/loss_utils.py` and torch's own `exp / sigmoid / normalize / optim.Adam`, runs them on seeded inputs with
autograd, and commits inputs + outputs as tests/golden/trainstep.npz.  tests/test_trainstep_cpu.py checks every
function below against those vectors.

    loss      gs-simp/utils/loss_utils.py:17-18 (l1_loss), :23-31 (window), :33-62 (ssim / _ssim), composed as
              gs-simp/train.py:91-92
    activate  gs-simp/scene/gaussian_model.py:33-41 (setup_functions), :95-115 (getters)
    adam      gs-simp/scene/gaussian_model.py:154-165 (six groups, Adam(lr=0.0, eps=1e-15)); the update rule is
              torch/optim/adam.py `_single_tensor_adam` (torch 2.x), which train.py:127 runs through optimizer.step()

Arithmetic is float64 on float32 inputs (the reference computes in float32): comparisons against the reference
and against the CUDA kernels use the tolerances written in the tests, not bit equality.
"""
from __future__ import annotations

import math

import numpy as np

WINDOW_SIZE = 11
SIGMA = 1.5
C1 = 0.01 ** 2
C2 = 0.03 ** 2


def gaussian_window(window_size: int = WINDOW_SIZE, sigma: float = SIGMA) -> np.ndarray:
    """loss_utils.py:23-25: float32 tensor of python-double exps, divided by its float32 sum."""
    g = np.array([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                 dtype=np.float32)
    return (g / g.sum(dtype=np.float32)).astype(np.float32)


def _filter2d(a: np.ndarray, w1: np.ndarray) -> np.ndarray:
    """Depthwise conv2d with the separable window w1 x w1, zero padding window_size // 2 (loss_utils.py:46-55;
    the 2-D window is the outer product of the 1-D one, :28-29).  a: (C,H,W) float64."""
    k = len(w1)
    h = k // 2
    C, H, W = a.shape
    pad = np.zeros((C, H + 2 * h, W + 2 * h), dtype=np.float64)
    pad[:, h:h + H, h:h + W] = a
    tmp = np.zeros((C, H + 2 * h, W), dtype=np.float64)
    for j in range(k):
        tmp += w1[j] * pad[:, :, j:j + W]
    out = np.zeros((C, H, W), dtype=np.float64)
    for j in range(k):
        out += w1[j] * tmp[:, j:j + H, :]
    return out


def ssim_map(img1: np.ndarray, img2: np.ndarray):
    """loss_utils.py:45-60.  Returns (ssim_map, intermediates)."""
    w = gaussian_window().astype(np.float64)
    x = img1.astype(np.float64)
    y = img2.astype(np.float64)
    mu1, mu2 = _filter2d(x, w), _filter2d(y, w)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq = _filter2d(x * x, w) - mu1_sq
    sigma2_sq = _filter2d(y * y, w) - mu2_sq
    sigma12 = _filter2d(x * y, w) - mu1_mu2
    A = 2 * mu1_mu2 + C1
    B = 2 * sigma12 + C2
    Cc = mu1_sq + mu2_sq + C1
    D = sigma1_sq + sigma2_sq + C2
    s = (A * B) / (Cc * D)
    return s, dict(mu1=mu1, mu2=mu2, A=A, B=B, C=Cc, D=D, w=w)


def loss_forward(image: np.ndarray, gt: np.ndarray, lambda_dssim: float):
    """train.py:91-92 -> (Ll1, ssim, loss) as python floats."""
    l1 = float(np.abs(image.astype(np.float64) - gt.astype(np.float64)).mean())   # loss_utils.py:18
    s, _ = ssim_map(image, gt)
    ss = float(s.mean())                                                            # loss_utils.py:62
    return l1, ss, (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - ss)


def loss_backward(image: np.ndarray, gt: np.ndarray, lambda_dssim: float, dL_dloss: float = 1.0) -> np.ndarray:
    """d loss / d image, what autograd produces for train.py:93 `loss.backward()` at the image.
    The filter is symmetric and zero padded, hence self-adjoint: the gradient of sum(ssim_map) is the same filter
    applied to the partial derivatives with respect to the three filtered moments that involve img1."""
    x = image.astype(np.float64)
    y = gt.astype(np.float64)
    n = x.size
    s, t = ssim_map(image, gt)
    mu1, mu2, A, B, Cc, D, w = t["mu1"], t["mu2"], t["A"], t["B"], t["C"], t["D"], t["w"]
    ds_dmu1 = 2 * mu2 * (B - A) / (Cc * D) + 2 * mu1 * s * (1.0 / D - 1.0 / Cc)
    ds_dxx = -s / D
    ds_dxy = 2 * A / (Cc * D)
    dsum = _filter2d(ds_dmu1, w) + 2 * x * _filter2d(ds_dxx, w) + y * _filter2d(ds_dxy, w)
    g = (1.0 - lambda_dssim) * np.sign(x - y) / n - lambda_dssim * dsum / n
    return dL_dloss * g


# ---------------------------------------------------------------------------------------------- activations
def activate_forward(raw_scales, raw_rotations, raw_opacities):
    """gaussian_model.py:95-115: exp, F.normalize (v / max(||v||, 1e-12)), sigmoid."""
    rs, rq, ro = (np.asarray(a, dtype=np.float64) for a in (raw_scales, raw_rotations, raw_opacities))
    scales = np.exp(rs)
    n = np.maximum(np.sqrt((rq * rq).sum(-1, keepdims=True)), 1e-12)
    return scales, rq / n, 1.0 / (1.0 + np.exp(-ro))


def activate_backward(raw_scales, raw_rotations, raw_opacities, g_scales, g_rotations, g_opacities):
    """Chain rule of the three activations (torch derivatives.yaml: exp -> grad * result; sigmoid -> grad * (1 - y) * y;
    normalize = div by clamp_min(norm, eps): the clamp passes gradient where norm >= eps)."""
    rs, rq, ro, gs, gq, go = (np.asarray(a, dtype=np.float64) for a in
                              (raw_scales, raw_rotations, raw_opacities, g_scales, g_rotations, g_opacities))
    d_scales = gs * np.exp(rs)
    norm = np.sqrt((rq * rq).sum(-1, keepdims=True))
    safe = np.maximum(norm, 1e-300)
    yq = rq / safe
    d_rot = np.where(norm >= 1e-12, (gq - yq * (yq * gq).sum(-1, keepdims=True)) / safe, gq / 1e-12)
    y = 1.0 / (1.0 + np.exp(-ro))
    return d_scales, d_rot, go * (1.0 - y) * y


# ---------------------------------------------------------------------------------------------- Adam
def adam_step(param, grad, exp_avg, exp_avg_sq, step: int, lr, beta1: float = 0.9, beta2: float = 0.999,
              eps: float = 1e-15):
    """One step of torch.optim.Adam (`_single_tensor_adam`, no weight decay / amsgrad / maximize) in float32, the
    reference's dtype.  `lr` is a scalar or an array broadcastable to param (per-element learning rates model the
    f_dc / f_rest columns of one SH tensor).  Returns new (param, exp_avg, exp_avg_sq); `step` is 1-based."""
    f = np.float32
    p, g, m, v = (np.asarray(a, dtype=f) for a in (param, grad, exp_avg, exp_avg_sq))
    m = (m + (g - m) * f(1.0 - beta1)).astype(f)                          # exp_avg.lerp_(grad, 1 - beta1)
    v = (v * f(beta2) + f(1.0 - beta2) * g * g).astype(f)                 # mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = np.asarray(lr, dtype=np.float64) / bc1
    denom = (np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)).astype(f)
    p = (p - step_size.astype(f) * (m / denom)).astype(f)                 # addcdiv_(exp_avg, denom, value=-step_size)
    return p, m, v


# ---------------------------------------------------------------------------------------------- image output
def save_image_u8(image, affine=None) -> np.ndarray:
    """The tensor half of torchvision.utils.save_image for one image, as the reference calls it after render()
    (gs-simp/render.py:36-39, render_depth.py:39, gen_seq.py:45-55): make_grid repeats a 1-channel image to 3
    channels; then `grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(uint8)` -- fp32 multiply, fp32 add,
    clamp, truncation.  torchvision is an unpinned third-party dependency of the reference (not importable here);
    this restates its published algorithm.  `affine` = (lo, inv_range): normalize_0_to_1 of scene/helpers.py:159-162
    applied first, in fp32.  (C,H,W) float -> (H,W,3) uint8.  NaN -> 0."""
    x = np.asarray(image, dtype=np.float32)
    if affine is not None:
        x = ((x - np.float32(affine[0])).astype(np.float32) * np.float32(affine[1])).astype(np.float32)
    if x.shape[0] == 1:
        x = np.repeat(x, 3, axis=0)
    v = ((x * np.float32(255.0)).astype(np.float32) + np.float32(0.5)).astype(np.float32)
    v = np.where(np.isnan(v), np.float32(0), v)
    return np.clip(v, 0, 255).astype(np.uint8).transpose(1, 2, 0).copy()
