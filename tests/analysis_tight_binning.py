"""CPU analysis (not a test): how many (tile, Gaussian) instances of the reference's 3-sigma square rect can never
contribute to any pixel of their tile?  Two tile-level tests, both conservative and both already applied per SUB-tile by
the blend kernels (blend_common.cuh): the {alpha >= 1/255} ellipse's bounding box against the tile's pixel-centre
rectangle, and the exact ellipse-vs-rectangle minimum.  Instances that fail them are sorted and staged for nothing.

usage: python tests/analysis_tight_binning.py [workload]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import scenes as S  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.util import oracle_forward  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "headline"
SLACK = 0.05
O.build()
sc = S.make_config_scene(workload)
W, H = sc["W"], sc["H"]
f = oracle_forward(O, sc)
N, G = f.num_rendered, f.ranges.shape[0]
gx = (W + 15) // 16
lens = (f.ranges[:, 1] - f.ranges[:, 0]).astype(np.int64)
tile_of = np.repeat(np.arange(G, dtype=np.int64), lens)
pos_in_tile = np.arange(N, dtype=np.int64) - np.repeat(f.ranges[:, 0].astype(np.int64), lens)
gid = f.point_list.astype(np.int64)
co = f.conic_opacity.astype(np.float64)
a, b, c, op = co[:, 0], co[:, 1], co[:, 2], co[:, 3]
det = a * c - b * b
with np.errstate(divide="ignore", invalid="ignore"):
    cov_xx, cov_yy = c / det, a / det
    tau = np.log(255.0 * op) + SLACK
    hx = np.sqrt(np.maximum(2.0 * tau * cov_xx, 0.0)) * 1.001 + 0.01
    hy = np.sqrt(np.maximum(2.0 * tau * cov_yy, 0.0)) * 1.001 + 0.01
dead = ~(tau > 0)
hx[dead], hy[dead] = -1.0, -1.0
mx, my = f.means2D[:, 0].astype(np.float64), f.means2D[:, 1].astype(np.float64)
tx0 = (tile_of % gx) * 16.0
ty0 = (tile_of // gx) * 16.0
x, y, ex, ey = mx[gid], my[gid], hx[gid], hy[gid]
bbox_keep = (x + ex >= tx0) & (x - ex <= tx0 + 15.0) & (y + ey >= ty0) & (y - ey <= ty0 + 15.0) & (ex >= 0)
gx0, gx1 = x - (tx0 + 15.0), x - tx0
gy0, gy1 = y - (ty0 + 15.0), y - ty0
A_, B_, C_ = a[gid], b[gid], c[gid]
def q(dx_, dy_):
    return A_ * dx_ * dx_ + 2.0 * B_ * dx_ * dy_ + C_ * dy_ * dy_
inside_c = (gx0 <= 0) & (gx1 >= 0) & (gy0 <= 0) & (gy1 >= 0)
cand = []
for xe in (gx0, gx1):
    cand.append(q(xe, np.clip(-B_ * xe / C_, gy0, gy1)))
for ye in (gy0, gy1):
    cand.append(q(np.clip(-B_ * ye / A_, gx0, gx1), ye))
qmin = np.where(inside_c, 0.0, np.minimum.reduce(cand))
exact_keep = bbox_keep & (qmin <= 2.0 * tau[gid])
nc_pad = np.zeros((((H + 15) // 16) * 16, gx * 16), dtype=np.int64)
nc_pad[:H, :W] = f.n_contrib
tile_last = nc_pad.reshape(-1, 16, gx, 16).max(axis=(1, 3)).reshape(-1)
walked = pos_in_tile < tile_last[tile_of]
# per-Gaussian: tiles of the reference rect vs tiles of the bbox rect (what K1 could store instead)
out = {"workload": workload, "N": int(N),
       "kept_by_bbox": float(bbox_keep.mean()), "kept_by_exact": float(exact_keep.mean()),
       "walked": float(walked.mean()),
       "walked_kept_by_bbox": float((walked & bbox_keep).sum() / max(walked.sum(), 1)),
       "walked_kept_by_exact": float((walked & exact_keep).sum() / max(walked.sum(), 1))}
# the rect K1 could store: tiles overlapped by the bbox, clipped to the reference rect (a sub-rectangle, no per-tile test)
txi, tyi = tile_of % gx, tile_of // gx
bx0 = np.floor((x - ex) / 16.0); bx1 = np.floor((x + ex) / 16.0)   # pixel centre p belongs to tile floor(p / 16)
by0 = np.floor((y - ey) / 16.0); by1 = np.floor((y + ey) / 16.0)
# a tile is kept when some pixel centre of it lies inside the bbox: centre columns tx0 .. tx0 + 15
rect_keep = (np.ceil(x - ex) <= tx0 + 15.0) & (np.floor(x + ex) >= tx0) & (np.ceil(y - ey) <= ty0 + 15.0) & (np.floor(y + ey) >= ty0) & (ex >= 0)
out["kept_by_integer_bbox_rect"] = float(rect_keep.mean())
print(json.dumps(out, indent=1))
