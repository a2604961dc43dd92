"""Build-container-only fuzz (needs /root/reference; not collected by pytest): the numpy oracle of the training-step
stages (oracle/trainstep_oracle.py -- the checker of the CUDA loss / Adam kernels) against the REFERENCE's own
utils/loss_utils.py (l1_loss, ssim, autograd) and torch.optim.Adam on random shapes and values, beyond the committed
golden vectors.  Tolerances as in tests/test_trainstep_cpu.py.   usage: python tests/golden/fuzz_trainstep_oracle_vs_reference.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/gs-simp")
from utils.loss_utils import l1_loss, ssim  # noqa: E402
from oracle import trainstep_oracle as T  # noqa: E402

torch.set_num_threads(2)
rng = np.random.default_rng(1)
worst = dict(loss=0.0, grad=0.0, adam_p=0.0)
n_loss = n_adam = 0
for it in range(120):
    C = int(rng.choice([1, 3])); H = int(rng.integers(1, 70)); W = int(rng.integers(1, 90)); lam = float(rng.choice([0.2, 0.0, 1.0, 0.5]))
    torch.manual_seed(int(rng.integers(1 << 30)))
    gt = torch.rand(C, H, W)
    img = (gt + float(rng.uniform(0, 0.3)) * torch.randn(C, H, W)).clamp(0, 1)
    if it % 3 == 0:
        img[:, ::2] = gt[:, ::2]                          # exact equalities: sign(0) = 0 in the L1 gradient
    img.requires_grad_(True)
    l1, s = l1_loss(img, gt), ssim(img, gt)
    loss = (1.0 - lam) * l1 + lam * (1.0 - s)
    loss.backward()
    got = T.loss_forward(img.detach().numpy(), gt.numpy(), lam)
    e = max(abs(got[0] - l1.item()), abs(got[1] - s.item()), abs(got[2] - loss.item()))
    g = T.loss_backward(img.detach().numpy(), gt.numpy(), lam)
    # relative to the gradient's size, with a floor: for img == gt the true gradient is 0 and both sides hold ~1e-10 of noise
    ge = float(np.abs(g - img.grad.numpy()).max() / max(np.abs(img.grad.numpy()).max(), 1.0 / (C * H * W)))   # 1/(CHW): the size of an L1 gradient entry
    worst["loss"], worst["grad"] = max(worst["loss"], e), max(worst["grad"], ge)
    assert e < 1e-5 and ge < 1e-4, (it, C, H, W, lam, e, ge)
    n_loss += 1
for it in range(40):
    n = int(rng.integers(1, 3000)); lr = float(10 ** rng.uniform(-5, -1)); steps = int(rng.integers(1, 8))
    torch.manual_seed(int(rng.integers(1 << 30)))
    p = torch.nn.Parameter(torch.randn(n))
    opt = torch.optim.Adam([p], lr=lr, eps=1e-15)
    p_np, m, v = p.detach().numpy().copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for t in range(1, steps + 1):
        gr = torch.randn(n) * float(10 ** rng.uniform(-4, 0))
        p.grad = gr.clone()
        opt.step()
        p_np, m, v = T.adam_step(p_np, gr.numpy(), m, v, t, lr)
        e = float(np.abs(p_np - p.detach().numpy()).max())
        worst["adam_p"] = max(worst["adam_p"], e / max(1.0, float(np.abs(p_np).max())))
        assert e <= 3e-7 * max(1.0, float(np.abs(p_np).max())), (it, t, e)
    n_adam += 1
print("loss cases", n_loss, "adam cases", n_adam, "worst", worst)
