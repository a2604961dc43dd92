"""Launches every training-step kernel (csrc/loss.cu, csrc/optim.cu) twice at a bounded size, for one `ncu --set full`
capture:  ncu --set full -k regex:'ssim_l1|loss_finalize|adam_kernel|activate_' ... python tools/exp_trainstep_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C  # noqa: E402
from multiview_inpaint_b200.multiview import GradArena  # noqa: E402
from multiview_inpaint_b200.trainstep import GaussianParamArena  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M, H, W, DEV = 16, 1008, 1600, "cuda"
torch.manual_seed(0)
gt = torch.rand(3, H, W, device=DEV)
img = (gt + 0.05 * torch.randn_like(gt)).clamp(0, 1).contiguous()
pa = GaussianParamArena(P, M, DEV)
pa.param.normal_()
arena = GradArena(P, M, DEV)
arena.flat.normal_().mul_(1e-3)
lrs = dict(xyz=0.00016, f_dc=0.0025, f_rest=0.0025 / 20, opacity=0.05, scaling=0.005, rotation=0.001)
for _ in range(2):
    out3, temp = _C.loss_l1_ssim_forward(img, gt, 0.2)
    _C.loss_l1_ssim_backward(img, gt, 0.2, temp)
    pa.activate()
    pa.apply_gradients(arena, lrs)      # activate_backward + adam_kernel
torch.cuda.synchronize()
print("done", out3.tolist())
