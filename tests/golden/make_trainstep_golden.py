"""Generates tests/golden/trainstep.npz by RUNNING the reference's own code on seeded inputs (build container
only: /root/reference must exist; the .npz is committed and travels to the GPU box).

    loss      imports /root/reference/gs-simp/utils/loss_utils.py (l1_loss, ssim) and composes them exactly as
              gs-simp/train.py:91-93 does, then loss.backward() for d loss / d image          -> loss_*
    activate  torch.exp / torch.nn.functional.normalize / torch.sigmoid -- the callables
              gs-simp/scene/gaussian_model.py:33-41 installs -- under autograd                -> act_*
    adam      torch.optim.Adam over the six parameter groups of gaussian_model.py:154-165 (lr=0.0, eps=1e-15),
              five optimizer.step() calls on seeded gradients                                   -> adam_*

usage: python tests/golden/make_trainstep_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/gs-simp"
sys.path.insert(0, REF)
from utils.loss_utils import l1_loss, ssim  # noqa: E402

torch.manual_seed(20240229)
torch.set_num_threads(1)
out = {}

# ---- loss: three shapes (tile-aligned, ragged both ways with partial tiles, smaller than the window) ----
LAMBDA = 0.2   # arguments/__init__.py:88
cases = {"a": (3, 48, 64), "b": (3, 37, 53), "c": (3, 7, 9), "d": (1, 20, 70)}
for name, (C, H, W) in cases.items():
    gt = torch.rand(C, H, W)
    # a rendered image: the target plus structured + random error, some pixels exactly equal (sign(0) = 0)
    img = (gt + 0.15 * torch.randn(C, H, W) + 0.1 * torch.sin(torch.arange(W) / 3.0)).clamp(0, 1)
    img[:, ::5, ::7] = gt[:, ::5, ::7]
    img.requires_grad_(True)
    Ll1 = l1_loss(img, gt)
    s = ssim(img, gt)
    loss = (1.0 - LAMBDA) * Ll1 + LAMBDA * (1.0 - s)
    loss.backward()
    out[f"loss_{name}_img"] = img.detach().numpy().copy()
    out[f"loss_{name}_gt"] = gt.numpy().copy()
    out[f"loss_{name}_out"] = np.array([Ll1.item(), s.item(), loss.item()], dtype=np.float64)
    out[f"loss_{name}_grad"] = img.grad.numpy().copy()
out["loss_lambda"] = np.array(LAMBDA)

# ---- activations under autograd ----
P = 257
raw_s = (torch.randn(P, 3) * 0.8 - 3.0).requires_grad_(True)
raw_q = torch.randn(P, 4)
raw_q[3] = 0.0                       # a zero quaternion: normalize clamps the norm at eps
raw_q[5] *= 1e-3
raw_q.requires_grad_(True)
raw_o = (torch.randn(P, 1) * 2.5).requires_grad_(True)
sc, q, o = torch.exp(raw_s), torch.nn.functional.normalize(raw_q), torch.sigmoid(raw_o)
gs, gq, go = torch.randn(P, 3), torch.randn(P, 4), torch.randn(P, 1)
(sc * gs).sum().backward()
(q * gq).sum().backward()
(o * go).sum().backward()
for k, v in dict(raw_s=raw_s, raw_q=raw_q, raw_o=raw_o, s=sc, q=q, o=o, gs=gs, gq=gq, go=go,
                 d_raw_s=raw_s.grad, d_raw_q=raw_q.grad, d_raw_o=raw_o.grad).items():
    out[f"act_{k}"] = v.detach().numpy().copy()

# ---- Adam: the six groups, learning rates of arguments/__init__.py:79-86 with spatial_lr_scale 1 ----
M = 4
P = 61
shapes = {"xyz": (P, 3), "f_dc": (P, 1, 3), "f_rest": (P, M - 1, 3), "opacity": (P, 1), "scaling": (P, 3),
          "rotation": (P, 4)}
lrs = {"xyz": 0.00016, "f_dc": 0.0025, "f_rest": 0.0025 / 20.0, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001}
params = {k: torch.nn.Parameter(torch.randn(*shp)) for k, shp in shapes.items()}
opt = torch.optim.Adam([{"params": [params[k]], "lr": lrs[k], "name": k} for k in shapes], lr=0.0, eps=1e-15)
STEPS = 5
for k in shapes:
    out[f"adam_p0_{k}"] = params[k].detach().numpy().copy()
    out[f"adam_lr_{k}"] = np.array(lrs[k])
for t in range(STEPS):
    for k in shapes:
        g = torch.randn(*shapes[k]) * (10.0 ** float(torch.randint(-6, 1, (1,))))
        if t == 2 and k == "f_rest":
            g.zero_()                                # an all-zero gradient step (invisible Gaussians)
        params[k].grad = g
        out[f"adam_g{t}_{k}"] = g.numpy().copy()
    opt.step()
    for k in shapes:
        st = opt.state[params[k]]
        out[f"adam_p{t + 1}_{k}"] = params[k].detach().numpy().copy()
        out[f"adam_m{t + 1}_{k}"] = st["exp_avg"].numpy().copy()
        out[f"adam_v{t + 1}_{k}"] = st["exp_avg_sq"].numpy().copy()
out["adam_steps"] = np.array(STEPS)
out["adam_M"] = np.array(M)

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "trainstep.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, f"{os.path.getsize(dst) / 1024:.1f} KiB,", len(out), "arrays; torch", torch.__version__)
