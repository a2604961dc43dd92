"""Generates tests/golden/lr_schedule.npz by RUNNING the reference's own get_expon_lr_func
(gs-simp/utils/general_utils.py:31-64, the position learning-rate schedule GaussianModel.update_learning_rate applies
every iteration, scene/gaussian_model.py:164-175, train.py:68) for the two optimisation presets of
gs-simp/arguments/__init__.py (:79-82 and :100-103) and a delayed variant.  Build container only.

usage: python tests/golden/make_lr_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/gs-simp")
from utils.general_utils import get_expon_lr_func  # noqa: E402

CASES = {
    # name: (lr_init, lr_final, lr_delay_steps, lr_delay_mult, max_steps)   -- lr_init already times spatial_lr_scale
    "default_30k": (0.00016 * 4.7, 0.0000016 * 4.7, 0, 0.01, 30_000),
    "inpaint_300": (0.001 * 2.0, 0.00002 * 2.0, 0, 0.02, 300),
    "delayed": (0.01, 0.0001, 500, 0.01, 10_000),
    "disabled": (0.0, 0.0, 0, 1.0, 1000),
}
steps = np.array([-1, 0, 1, 2, 10, 99, 100, 299, 300, 301, 499, 500, 501, 1000, 2999, 15000, 29999, 30000, 30001, 100000])
out = {"steps": steps}
for name, args in CASES.items():
    f = get_expon_lr_func(lr_init=args[0], lr_final=args[1], lr_delay_steps=args[2], lr_delay_mult=args[3], max_steps=args[4])
    out[f"{name}_args"] = np.array(args, dtype=np.float64)
    out[f"{name}_lr"] = np.array([float(f(int(s))) for s in steps], dtype=np.float64)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lr_schedule.npz")
np.savez(dst, **out)
print("wrote", dst)
