"""CPU analysis (not a test; lives under tests/ because it drives the oracle, which only test infrastructure may):
how much of the blend kernels' work is useful?  For the headline view it counts, over all (tile, Gaussian) instances,

  hits        (8x4 sub-tile, Gaussian) pairs that pass the bounding-box cull of blend_common.cuh::stage_entry
              (these are the iterations of the K6 / K7 inner loops: 28 warp instructions each in K7 before the vote)
  live hits   hits in which at least one of the 32 pixels really contributes (alpha >= 1/255, power <= 0, and the
              pixel has not terminated: list position < n_contrib) -- K7 runs its full 107-instruction path for these
  lanes       contributing pixels per live hit (lane efficiency of the pixel-parallel warp)

which decides between the options of DESIGN.md section 10 item 3 (tighter culling vs a transposed, lane = Gaussian
backward).  Geometry state comes from the oracle's forward (bit-identical to K1, tests/test_parity_gpu.py); the
cull extents are recomputed as preprocess.cu computes them.

usage: python tests/analysis_blend_hits.py [workload] [sample_instances]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import scenes as S  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.util import oracle_forward  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "headline"
n_sample = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
SLACK = 0.05  # GSR_POWER_SLACK

O.build()
sc = S.make_config_scene(workload)
W, H = sc["W"], sc["H"]
f = oracle_forward(O, sc)
N, G = f.num_rendered, f.ranges.shape[0]
gx = (W + 15) // 16
print(f"{workload}: P={sc['P']} V={(f.radii > 0).sum()} N={N} G={G}", file=sys.stderr)

# per-instance: tile, list position inside the tile, Gaussian id
lens = (f.ranges[:, 1] - f.ranges[:, 0]).astype(np.int64)
tile_of = np.repeat(np.arange(G, dtype=np.int64), lens)
pos_in_tile = np.arange(N, dtype=np.int64) - np.repeat(f.ranges[:, 0].astype(np.int64), lens)
gid = f.point_list.astype(np.int64)

# cull extents as K1 stores them: half extents of the {alpha >= 1/255} ellipse's bounding box, from the covariance
co = f.conic_opacity.astype(np.float64)
a, b, c, op = co[:, 0], co[:, 1], co[:, 2], co[:, 3]
det = a * c - b * b
with np.errstate(divide="ignore", invalid="ignore"):
    cov_xx, cov_yy = c / det, a / det
    tau = np.log(255.0 * op) + SLACK                     # power >= -tau  <=>  alpha >= 1/255 (with slack)
    hx = np.sqrt(np.maximum(2.0 * tau * cov_xx, 0.0)) * 1.001 + 0.01
    hy = np.sqrt(np.maximum(2.0 * tau * cov_yy, 0.0)) * 1.001 + 0.01
dead = ~(tau > 0)                                        # 255 * opacity < 1: nothing can contribute
hx[dead], hy[dead] = -1.0, -1.0

mx, my = f.means2D[:, 0].astype(np.float64), f.means2D[:, 1].astype(np.float64)
tx0 = (tile_of % gx) * 16.0
ty0 = (tile_of // gx) * 16.0
x, y, ex, ey = mx[gid], my[gid], hx[gid], hy[gid]
# K7 walks a sub-tile's list only up to warp_last = max n_contrib of its 32 pixels (entries behind every pixel's
# last contributor are masked off before the loop); K6 stops a sub-tile once all its pixels are done -- same bound.
nc_pad = np.zeros((((H + 15) // 16) * 16, gx * 16), dtype=np.int64)
nc_pad[:H, :W] = f.n_contrib
warp_last = nc_pad.reshape(-1, 4, 4, gx, 2, 8).max(axis=(3 - 1, 5)).transpose(0, 2, 1, 3)   # [tile_y, tile_x, row, col]
warp_last = warp_last.reshape(-1, 8)                                                          # [tile, v = row * 2 + col]
hits = np.zeros(N, dtype=np.int8)
bbox_hits = np.zeros(N, dtype=np.int8)
hit_mask = np.zeros((N, 8), dtype=bool)
for v in range(8):                                       # warp v = row * 2 + col: 8 wide, 4 high
    sx0, sy0 = tx0 + (v & 1) * 8.0, ty0 + (v >> 1) * 4.0
    m = (x + ex >= sx0) & (x - ex <= sx0 + 7.0) & (y + ey >= sy0) & (y - ey <= sy0 + 3.0) & (ex >= 0)
    bbox_hits += m
    m &= pos_in_tile < warp_last[tile_of, v]
    hit_mask[:, v] = m
    hits += m
total_hits = int(hits.sum())

# pixel-exact part on a random sample of instances
rng = np.random.default_rng(0)
sel = rng.choice(N, size=min(n_sample, N), replace=False)
n_contrib = f.n_contrib.astype(np.int64)
live, lanes = 0, 0
exact_culled, exact_culled_live = 0, 0
hits_sel = int(hits[sel].sum())
lane_hist = np.zeros(33, dtype=np.int64)
ly, lx = np.mgrid[0:4, 0:8]
for v in range(8):
    s = sel[hit_mask[sel, v]]
    if s.size == 0:
        continue
    sx0 = (tx0[s] + (v & 1) * 8.0)[:, None, None] + lx[None]
    sy0 = (ty0[s] + (v >> 1) * 4.0)[:, None, None] + ly[None]
    g = gid[s]
    dx = mx[g][:, None, None] - sx0
    dy = my[g][:, None, None] - sy0
    power = -0.5 * (a[g][:, None, None] * dx * dx + c[g][:, None, None] * dy * dy) - b[g][:, None, None] * dx * dy
    alpha = np.minimum(0.99, op[g][:, None, None] * np.exp(np.minimum(power, 0.0)))
    inside = (sx0 < W) & (sy0 < H)
    px, py = np.minimum(sx0, W - 1).astype(np.int64), np.minimum(sy0, H - 1).astype(np.int64)
    alive = pos_in_tile[s][:, None, None] < n_contrib[py, px]
    contrib = inside & alive & (power <= 0.0) & (alpha >= 1.0 / 255.0)
    k = contrib.reshape(len(s), -1).sum(1)
    # would an EXACT ellipse-vs-rect test at staging time have culled this hit?  (min over the sub-tile's pixel centres'
    # bounding rect of the quadratic form > 2 tau: no pixel can reach alpha >= 1/255)  -- geometric misses of the bbox cull
    gx0, gx1 = mx[g] - (tx0[s] + (v & 1) * 8.0 + 7.0), mx[g] - (tx0[s] + (v & 1) * 8.0)       # dx range [gx0, gx1]
    gy0, gy1 = my[g] - (ty0[s] + (v >> 1) * 4.0 + 3.0), my[g] - (ty0[s] + (v >> 1) * 4.0)
    A_, B_, C_ = a[g], b[g], c[g]
    def q(dx_, dy_):
        return A_ * dx_ * dx_ + 2.0 * B_ * dx_ * dy_ + C_ * dy_ * dy_
    inside_c = (gx0 <= 0) & (gx1 >= 0) & (gy0 <= 0) & (gy1 >= 0)
    cand = []
    for xe in (gx0, gx1):
        cand.append(q(xe, np.clip(-B_ * xe / C_, gy0, gy1)))
    for ye in (gy0, gy1):
        cand.append(q(np.clip(-B_ * ye / A_, gx0, gx1), ye))
    qmin = np.where(inside_c, 0.0, np.minimum.reduce(cand))
    exact_keep = qmin <= 2.0 * tau[g]
    exact_culled += int((~exact_keep).sum())
    exact_culled_live += int(((~exact_keep) & (k > 0)).sum())      # must be 0: the test is conservative
    live += int((k > 0).sum())
    lanes += int(k.sum())
    lane_hist += np.bincount(k, minlength=33)

out = {
    "workload": workload, "P": int(sc["P"]), "V": int((f.radii > 0).sum()), "N": int(N),
    "bbox_overlaps_per_instance": float(bbox_hits.sum()) / N,          # before the saturation bound
    "hits_per_instance": total_hits / N,
    "instances_walked": float((pos_in_tile < warp_last[tile_of].max(1)).mean()),    # list position below the tile's last contributor
    "instances_with_no_hit": float((hits == 0).mean()),
    "sample_instances": int(len(sel)),
    "live_fraction_of_hits": live / max(hits_sel, 1),
    "hits_an_exact_ellipse_rect_test_would_cull": exact_culled / max(hits_sel, 1),
    "live_hits_wrongly_culled_by_it": exact_culled_live,
    "live_hits_per_instance": live / len(sel),
    "contributing_lanes_per_live_hit": lanes / max(live, 1),
    "contributing_pairs_per_instance": lanes / len(sel),
    "lanes_histogram_of_hits": {str(i): int(n) for i, n in enumerate(lane_hist) if n},
}
# instruction model of blend_backward_kernel<false> (SASS counts): 28 per hit up to the vote,
# 107 for a live hit; the measured kernel executes 72.5 warp instructions per instance (5.95e8 / 8.21e6)
h, lv = out["hits_per_instance"], out["live_hits_per_instance"]
out["model_warp_inst_per_instance"] = {"early_exit_hits": 28 * (h - lv), "live_hits": 107 * lv, "sum": 28 * (h - lv) + 107 * lv}
print(json.dumps(out, indent=1))
