#!/usr/bin/env python
"""bench.py -- fwd+bwd views/s of the differentiable Gaussian rasterizer (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload headline]

A "step" is one pass of the hot path over one batch of synthetic views: every rank runs
forward + backward for `--views-per-rank` views of the named scene (gradients ADDED into one flat
arena), then the arena is all-reduced (N > 1).  Weak scaling: per-GPU work is fixed.

Prints ONE JSON line (rank 0).  `value` = views/s with inputs resident in HBM; `e2e` = the same
through the public API with that step's camera + loss-weight image copied from pinned host memory
and the scalar loss read back, inside the timed region.  `roofline` is for the dominant kernel,
timed live with CUDA events on the launching stream (gsr_profile_* hooks of the C ABI);
`cpu_baseline` / `--impl reference` time the CPU oracle (the reference's rasterizer source is not
in the tree and cannot be built here -- DESIGN.md), on a bounded tile-strided sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fwd+bwd views/s"
FWD_STAGES = ("preprocess", "depth_sort", "scan", "duplicate", "tile_sort", "tile_ranges", "blend_forward")
BWD_STAGES = ("accum_clear", "blend_backward", "geom_backward")
UNIT = "views/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline")
    ap.add_argument("--mode", default="train", choices=["train", "infer"],
                    help="train: fwd+bwd views/s (the BASELINE metric, default); infer: forward-only colour+depth views/s "
                         "(BASELINE config 4, render.py / render_depth.py), no collective")
    ap.add_argument("--views-per-rank", type=int, default=4)
    ap.add_argument("--total-views", type=int, default=0, help="STRONG scaling: a fixed batch of this many views per step, sharded over "
                    "the ranks (BASELINE configs 3 / 4: the 25-frame orbit, the 200-view render set); 0 = weak scaling, views-per-rank each")
    ap.add_argument("--no-balance", action="store_true", help="strong scaling: keep the round-robin view -> rank map instead of "
                    "balancing the ranks' instance counts (longest-processing-time assignment on the learnt num_rendered)")
    ap.add_argument("--orbit-deg", type=float, default=5.0, help="views lie on a +-deg orbit about the scene centre")
    ap.add_argument("--flags", type=int, default=32, help="GSR_FLAG_* bits (32 = tight binning, the Python layers' default; 0 = the reference's literal lists; 1 = 64-bit key binning)")
    ap.add_argument("--streams", type=int, default=1, help="view groups in flight per GPU (ViewPipeline depth; 1 = one stream): with the batched "
                    "front end the rank's views are split into this many groups, one stream each")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "nccl", "nvls"],
                    help="arena collective: NCCL, the in-switch multimem kernel, or whichever is faster here")
    ap.add_argument("--ar-chunks", type=int, default=8, help="Gaussian-range chunks of the pipelined per-Gaussian backward + in-switch all-reduce (1 = sequential)")
    ap.add_argument("--per-view-backward", action="store_true", help="K8+K9 per view (accumulate) instead of one batched launch per step")
    ap.add_argument("--no-batched", action="store_true", help="front end + blend view by view on the ViewPipeline's streams (round-1 structure) "
                    "instead of ONE launch per stage for all views of the rank (gsr_forward_views / gsr_backward_blend_views)")
    ap.add_argument("--e2e-blocking", action="store_true", help="e2e leg: read each step's loss with a blocking copy before queueing the next "
                                                               "step (round-1 behaviour) instead of one step later from a pinned ring")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-structure", action="store_true", help="skip the GSR_FLAG_REFERENCE ablation leg")
    ap.add_argument("--no-dropin", action="store_true", help="skip the drop-in leg (GaussianRasterizer + loss.backward(), one view at a time)")
    ap.add_argument("--no-train-step", action="store_true", help="skip the fused-optimisation-step leg (SURVEY 8f rows 1, 4)")
    ap.add_argument("--files", type=int, default=0, help="infer mode: also time views/s all the way to PNG files with N writer threads "
                    "(AsyncImageWriter) against the reference's blocking save_image structure")
    ap.add_argument("--cpu-tile-step", type=int, default=0, help="0 = auto (about 10-30 s of CPU work)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def survey_bytes(P, V, N, G, W, H, M):
    """SURVEY.md section 8(d) per-view figures, per stage: the traffic of the REFERENCE's structure (one 64-bit key sort
    of six passes, a scan, AoS reads).  Kept beside the design model below; not used for any printed GB/s."""
    pre_b = {1: 190, 4: 260, 16: 550}.get(M, 119 + 27 * M)
    return {
        "preprocess": P * (119 + 12 * M),
        "scan": 8 * P,
        "duplicate": 20 * P + 12 * N,
        "tile_sort": N * (8 + 6 * 24),
        "tile_ranges": 8 * N + 8 * G,
        "blend_forward": 44 * N + 24 * W * H,
        "blend_backward": 44 * N + 20 * W * H + 36 * N,
        "geom_backward": P * pre_b,
    }


def algorithmic_bytes(P, V, N, G, W, H, M, nv=1, clear_in_k1=False):
    """Bytes per view THIS design must move, per stage, for `nv` views sharing one batched launch
    (the per-Gaussian inputs are then read once for all of them):
      preprocess     reads 44 B of mean / scale / rotation / opacity per Gaussian and the 12 M-byte SH row of the V visible
                     ones (once per batch); writes the 48-byte blend record + depth + clamp bits per visible Gaussian and
                     radius, tiles_touched, depth key and tile rect (20 B) per Gaussian
      depth_sort     histogram read (4 P) + compacting first pass (4 P in, 8 V out) + three more passes of 16 V
      duplicate      order + rect in, offsets out (16 V), 8 N of (tile, id) pairs out; the digit counters stay in shared memory
      tile_sort      16 N per 8-bit pass of the tile id
      tile_ranges    ranges out (the 32-ary searches in the sorted keys touch a few lines per tile: not a streaming stage)
      blend_forward  list + records (44 N) and the per-pixel outputs (24 WH): an upper bound, the walk stops at saturation
      accum_clear    the packed 48-byte accumulator per Gaussian; with the batched front end K1 clears it on the way
                     (clear_in_k1: the 48 P bytes are charged to preprocess and this stage has no launch)
      blend_backward 80 N + 20 WH (as SURVEY: records in, nine reductions out per instance)
      geom_backward  batched K8+K9: parameters + SH rows once per batch, per view the record + accumulator of the visible
                     Gaussians (96 V) and radius + clamp bits (5 P); the gradient row (4 (11 + 3 M) B per Gaussian) and the
                     three statistics written once per batch"""
    tile_passes = (max(1, math.ceil(math.log2(max(G, 2)))) + 7) // 8
    grad_row = 4 * (3 + 3 * M + 1 + 3 + 4)
    out = {
        "preprocess": (44 * P + 12 * M * V) / nv + 53 * V + 20 * P + (48 * P if clear_in_k1 else 0),
        "depth_sort": 8 * P + 56 * V,
        "duplicate": 16 * V + 8 * N,
        "tile_sort": 16 * N * tile_passes,
        "tile_ranges": 12 * G,
        "blend_forward": 44 * N + 24 * W * H,
        "accum_clear": 48 * P,
        "blend_backward": 80 * N + 20 * W * H,
        "geom_backward": (44 * P + 12 * M * V + (grad_row + 12) * P) / nv + 96 * V + 5 * P,
    }
    if clear_in_k1:
        del out["accum_clear"]
    del out["tile_ranges"]   # latency of G x log32(N) probes, not bytes: no bandwidth figure
    return out


def view_roofline(alg, N, W, H, ms_view, peak, survey=None):
    """Whole-view bandwidth position.  The blend kernels enter with their per-pixel I/O only: their list gathers (44 N and
    80 N in the stage models: every binned instance) are an upper bound -- the walk stops at saturation -- and are served
    by the L2 (profiles/traffic.json), so charging them to HBM put whole views of early-saturating shapes above the peak."""
    b = sum(v for k, v in alg.items() if not k.startswith("blend"))
    if "blend_forward" in alg:
        b += 24 * W * H
    if "blend_backward" in alg:
        b += 20 * W * H
    gather = (44 * N if "blend_forward" in alg else 0) + (80 * N if "blend_backward" in alg else 0)
    g = b / (ms_view / 1000.0) / 1e9
    out = {"alg_bytes": int(b), "blend_list_gather_upper_bound_bytes": int(gather), "ms": ms_view, "gbps": g, "frac": g / peak}
    if survey:
        # SURVEY.md 8(d)'s B_view = B_fwd + B_bwd: what the REFERENCE's structure would move for this view (scan, six 64-bit
        # sort passes, 44 N + 80 N of blend gathers), with N = the instances THIS run binned (tight binning: fewer than the
        # reference's).  An equivalence figure -- bytes this design does not move -- kept because the survey's headline
        # "% of HBM roofline" is defined on it; it can exceed 1.
        sb = float(sum(v for v in survey.values() if v))
        out["survey_8d"] = {"bytes": int(sb), "equivalent_gbps": sb / (ms_view / 1000.0) / 1e9,
                            "equivalent_frac_of_hbm_peak": sb / (ms_view / 1000.0) / 1e9 / peak}
    return out


def issue_counters():
    """ncu counters of the issue-bound blend kernels (profiles/roofline_counters.json, written from the committed
    `ncu --set full` captures by tools/make_profiles.py): issue-slot utilisation is their roof, not HBM."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_counters.json")))
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksThrottleReason") or k.startswith("nvmlClocksEventReason")}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if isinstance(bit, int) and bit and (r & bit) == bit and "None" not in name and "All" not in name:
                        self.reasons.add(name.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", ""))
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# -------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle on the host cores
# -------------------------------------------------------------------------------------------------
def cpu_view_seconds(sc, cam, wt, tile_step, keep=None):
    """One fwd+bwd view on the CPU: full preprocess + binning, blend fwd/bwd on every
    `tile_step`-th tile; returns (estimated seconds for the full view, measured seconds, parts)."""
    import numpy as np
    from oracle import oracle as O
    kw = dict(means3D=sc["means3D"].numpy(), opacities=sc["opacities"].numpy(),
              viewmatrix=cam.world_view_transform.numpy(), projmatrix=cam.full_proj_transform.numpy(),
              campos=cam.camera_center.contiguous().numpy(), W=cam.image_width, H=cam.image_height,
              tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, sh_degree=sc["sh_degree"], shs=sc["shs"].numpy(),
              scales=sc["scales"].numpy(), rotations=sc["rotations"].numpy())
    t0 = time.perf_counter()
    f = O.bin_and_sort(O.preprocess(**kw))
    t1 = time.perf_counter()
    O.blend(f, np.zeros(3, np.float32), 0, tile_step)
    g_full = O.backward(f, wt, 0, tile_step)          # blend backward on the sample + full per-Gaussian backward
    t2 = time.perf_counter()
    # the per-Gaussian backward inside O.backward is not strided: time it alone to avoid scaling it
    t3 = time.perf_counter()
    g = O.backward(f, wt, 0, 1 << 30)        # tile_step >= G: only tile 0 blended, full K8+K9
    t4 = time.perf_counter()
    geom_bwd = t4 - t3
    blend_sample = max((t2 - t1) - geom_bwd, 0.0)
    est = (t1 - t0) + blend_sample * tile_step + geom_bwd
    if keep is not None:
        keep["f"], keep["g"] = f, g_full
    return est, (t2 - t0), dict(preprocess_bin_s=t1 - t0, blend_sample_s=blend_sample, geom_bwd_s=geom_bwd,
                                N=int(f.num_rendered), V=int((f.radii > 0).sum()))


def run_reference_install(args):
    """SURVEY 0.1 / 8d baseline B0: if the driver (or a maintainer) has put an install of the reference's rasterizer
    -- the third-party `diff_gaussian_rasterization` wheel the reference imports at gaussian_renderer/__init__.py:14 --
    under baseline/_ref/, time THAT through its own public API on this config and return its JSON line.  Returns None
    when there is no such install (the case in this repository: the package is neither in /root/reference nor in the
    offline wheelhouse), and the caller falls back to the CPU oracle port."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    pkg = os.path.join(ref_dir, "diff_gaussian_rasterization")
    if not os.path.isdir(pkg) or not any(f.startswith("_C") and f.endswith(".so") for f in os.listdir(pkg)):
        return None
    try:
        import importlib.util
        import torch
        if not torch.cuda.is_available():
            return None
        spec = importlib.util.spec_from_file_location("_ref_dgr", os.path.join(pkg, "__init__.py"), submodule_search_locations=[pkg])
        ref = importlib.util.module_from_spec(spec)
        sys.modules["_ref_dgr"] = ref
        spec.loader.exec_module(ref)
        from multiview_inpaint_b200 import scenes as S
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        cfg = S.CONFIGS[args.workload]
        sc = S.make_config_scene(args.workload)
        g = {k: sc[k].to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        cams = [c.to(dev) for c in S.orbit_cameras(max(args.views_per_rank, 2), cfg["W"], cfg["H"], max_deg=args.orbit_deg)]
        wt = S.loss_weights(cfg["W"], cfg["H"], cfg["seed"]).to(dev)
        bg = torch.zeros(3, device=dev)

        def view(c):
            rs = ref.GaussianRasterizationSettings(image_height=cfg["H"], image_width=cfg["W"], tanfovx=c.tanfovx, tanfovy=c.tanfovy,
                                                   bg=bg, scale_modifier=1.0, viewmatrix=c.world_view_transform,
                                                   projmatrix=c.full_proj_transform, sh_degree=cfg["sh_degree"],
                                                   campos=c.camera_center, prefiltered=False)
            m2d = torch.zeros_like(g["means3D"], requires_grad=True)
            out = ref.GaussianRasterizer(rs)(means3D=g["means3D"], means2D=m2d, opacities=g["opacities"], shs=g["shs"],
                                             scales=g["scales"], rotations=g["rotations"], colors_precomp=None, cov3D_precomp=None)
            (out[0] * wt).sum().backward()
            for t in g.values():
                t.grad = None
        for i in range(max(args.warmup, 3)):
            view(cams[i % len(cams)])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            view(cams[i % len(cams)])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        value = args.steps / (ms / 1000.0)
        return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, **cfg, "note": "the reference's own diff_gaussian_rasterization install under baseline/_ref, one view per step on one GPU"},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "GPU run of the reference extension (not a CPU baseline)"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    except Exception as ex:   # a broken install must not lose the reference arm: fall back to the oracle port
        sys.stderr.write(f"baseline/_ref present but unusable ({ex!r}); timing the CPU oracle port instead\n")
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from multiview_inpaint_b200 import scenes as S
    ref_line = run_reference_install(args)     # the reference's own CUDA rasterizer, if an install exists
    if ref_line is not None:
        print(json.dumps(ref_line), flush=True)
        return
    from oracle import oracle as O
    O.build()
    # all host cores, whatever the launcher exported: torchrun sets OMP_NUM_THREADS=1 for its workers
    O.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    cfg = S.CONFIGS[args.workload]
    sc = S.make_config_scene(args.workload)
    cams = S.orbit_cameras(max(args.views_per_rank, 2), cfg["W"], cfg["H"], max_deg=args.orbit_deg)
    wt = S.loss_weights(cfg["W"], cfg["H"], cfg["seed"]).numpy()
    G = ((cfg["W"] + 15) // 16) * ((cfg["H"] + 15) // 16)
    tile_step = args.cpu_tile_step or 1
    cores = O.num_threads()
    for i in range(args.warmup):
        cpu_view_seconds(sc, cams[i % len(cams)], wt, tile_step)
    ests, t0 = [], time.perf_counter()
    parts = {}
    for i in range(args.steps):
        est, _, parts = cpu_view_seconds(sc, cams[i % len(cams)], wt, tile_step)
        ests.append(est)
    wall = time.perf_counter() - t0
    est_view = sum(ests) / len(ests)
    value = 1.0 / est_view
    sample = (f"per step: 1 view of '{args.workload}' (P={cfg['P']}, {cfg['W']}x{cfg['H']}, SH deg {cfg['sh_degree']}): full "
              f"preprocess+binning+per-Gaussian backward, blend fwd+bwd on every {tile_step}-th of {G} tiles, "
              f"blend time scaled x{tile_step}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **cfg, "note": "CPU oracle port of the rasterizer (reference source not in tree)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, **parts},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
def bench_train_step(args, torch, _C, mv, S, sc, gauss, settings_list, dev, timed, av, mine, pipe, workspaces):
    """iterations/s of one optimisation step over this rank's views, fused vs torch structure (see run_ours)."""
    import torch.nn.functional as F
    from multiview_inpaint_b200.rasterizer import GaussianRasterizer
    from multiview_inpaint_b200.trainstep import GaussianParamArena, ViewLoss, fused_train_step
    P, M, H, W = sc["P"], sc["M"], sc["H"], sc["W"]
    lrs = dict(xyz=0.00016, f_dc=0.0025, f_rest=0.0025 / 20.0, opacity=0.05, scaling=0.005, rotation=0.001)   # arguments/__init__.py:79-86
    raw = dict(xyz=gauss["means3D"], f_dc=gauss["shs"][:, :1].contiguous(), f_rest=gauss["shs"][:, 1:].contiguous(),
               opacity=torch.logit(gauss["opacities"].clamp(1e-4, 1 - 1e-4)).reshape(P, 1), scaling=torch.log(gauss["scales"]),
               rotation=gauss["rotations"].clone())
    g = torch.Generator().manual_seed(77)
    gts = [torch.rand(3, H, W, generator=g).to(dev) for _ in mine]     # synthetic targets (the metric is time, not quality)
    steps = max(2, min(args.steps, 5))

    # fused
    pa = GaussianParamArena.from_tensors(raw["xyz"], raw["f_dc"], raw["f_rest"], raw["opacity"], raw["scaling"], raw["rotation"])
    arena = mv.GradArena(P, M, dev)
    losses = [ViewLoss(gt, 0.2, weight=1.0 / len(mine)) for gt in gts]
    caps, slots = [av.capacity(v) for v in mine], [av.slot(v) for v in mine]

    def step_fused():
        # apply=True with asynchronous views: the capacities carry a 25 % margin over the learnt high-water mark and the
        # overflow words are checked after the loop (`capacity_overflow` below)
        fused_train_step(pa, settings_list, losses, arena, lrs, flags=args.flags, pipeline=pipe, capacities=caps,
                         async_results=slots, workspaces=workspaces, apply=True, batched=not args.no_batched)
    l0 = _C.kernel_launches()
    for _ in range(2):
        step_fused()
    ms_f = timed(step_fused, steps)
    launches = (_C.kernel_launches() - l0) // (steps + 2)
    overflow = bool(av.check(mine))     # the parameters move: a view may outgrow its binning capacity (25 % margin)
    loss_fused = float(sum(l.out3[2].item() for l in losses) / len(losses))
    del pa, arena, losses

    # torch structure around the same rasterizer (drop-in autograd Function, one view at a time)
    leaves = {k: torch.nn.Parameter(v.clone()) for k, v in raw.items()}
    opt = torch.optim.Adam([{"params": [leaves[k]], "lr": lrs[k], "name": k} for k in leaves], lr=0.0, eps=1e-15)
    w1 = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)], device=dev)
    w1 = w1 / w1.sum()
    w2 = torch.outer(w1, w1).expand(3, 1, 11, 11).contiguous()
    blur = lambda t: F.conv2d(t[None], w2, padding=5, groups=3)[0]

    def step_torch():
        opt.zero_grad(set_to_none=True)
        for rs, gt in zip(settings_list, gts):
            shs = torch.cat((leaves["f_dc"], leaves["f_rest"]), dim=1)
            means2D = torch.zeros_like(leaves["xyz"], requires_grad=True)
            x, _radii, _depth = GaussianRasterizer(rs)(means3D=leaves["xyz"], means2D=means2D, opacities=torch.sigmoid(leaves["opacity"]),
                                                       shs=shs, scales=torch.exp(leaves["scaling"]), rotations=F.normalize(leaves["rotation"]))
            mx, my = blur(x), blur(gt)
            vx, vy, cxy = blur(x * x) - mx * mx, blur(gt * gt) - my * my, blur(x * gt) - mx * my
            ssim = (((2 * mx * my + 1e-4) * (2 * cxy + 9e-4)) / ((mx * mx + my * my + 1e-4) * (vx + vy + 9e-4))).mean()
            loss = (0.8 * (x - gt).abs().mean() + 0.2 * (1.0 - ssim)) / len(gts)
            loss.backward()
        opt.step()
    for _ in range(2):
        step_torch()
    ms_t = timed(step_torch, steps)
    nv = len(mine)
    return {"fused": {"iters_per_s": steps / (ms_f / 1000.0), "ms_per_step": ms_f / steps, "views_per_s": nv * steps / (ms_f / 1000.0),
                      "gpu_launches_per_step": int(launches), "mean_loss_last_step": loss_fused, "capacity_overflow": overflow},
            "torch_structure": {"iters_per_s": steps / (ms_t / 1000.0), "ms_per_step": ms_t / steps,
                                "views_per_s": nv * steps / (ms_t / 1000.0),
                                "note": "torch getters + cat, F.conv2d SSIM, autograd, torch.optim.Adam (six groups) around THIS repo's "
                                        "rasterizer (drop-in autograd Function), one view at a time, one optimizer.step() per iteration"},
            "views_per_step": nv, "steps": steps, "speedup": ms_t / ms_f}


def parity_against_oracle(torch, _C, gauss, rs, wt, kept, P, W, H, tile_step, flags):
    """Full-size parity (round-1 verdict item 1): the view the cpu_baseline leg has just pushed through the C oracle
    against the same view through the C ABI (exact-size path).  Integers bit-exact: radii, num_rendered, the sorted
    point list and the tile ranges; colour / depth against BASELINE's 1e-5 bar on the tiles the oracle blended;
    gradients as max |a - b| / max |b|."""
    import numpy as np
    f, g_ref = kept["f"], kept["g"]
    e = torch.empty(0, device=wt.device)
    tight = bool(flags & _C.FLAG_TIGHT_BINNING) and not (flags & (_C.FLAG_BINNING_KEY64 | _C.FLAG_REFERENCE))
    lit_flags = flags & ~_C.FLAG_TIGHT_BINNING     # the oracle holds the reference's literal lists

    def fwd(fl):
        return _C.rasterize_gaussians(
            rs.bg, gauss["means3D"], e, gauss["opacities"], gauss["scales"], gauss["rotations"], rs.scale_modifier, e, rs.viewmatrix,
            rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, gauss["shs"], rs.sh_degree, rs.campos, rs.prefiltered,
            flags=fl, capacity=0)
    n, color, radii, geom, binning, img, depth = fwd(lit_flags)
    st = _C.unpack_state(P, W, H, n, geom, binning, img, lit_flags)
    out = {"view": "first view of rank 0", "tile_step": tile_step, "num_rendered": [int(n), int(f.num_rendered)]}
    out["radii_equal"] = bool(np.array_equal(radii.cpu().numpy(), f.radii))
    out["point_list_equal"] = bool(n == f.num_rendered and np.array_equal(st["point_list"].cpu().numpy().view(np.uint32), f.point_list))
    out["ranges_equal"] = bool(np.array_equal(st["ranges"].cpu().numpy().view(np.uint32), f.ranges))
    gx = (W + 15) // 16
    tiles = np.arange(0, gx * ((H + 15) // 16), tile_step)
    mask = np.zeros((H, W), bool)
    for t in tiles:
        mask[(t // gx) * 16:(t // gx) * 16 + 16, (t % gx) * 16:(t % gx) * 16 + 16] = True
    c_err = np.abs(color.cpu().numpy() - f.color).max(0)[mask]
    d_err = np.abs(depth.cpu().numpy()[0] - f.depth[0])[mask]
    out["pixels_compared"] = int(mask.sum())
    out["colour_max_abs_err"] = float(c_err.max())
    out["colour_pixels_over_1e-5"] = int((c_err > 1e-5).sum())
    out["depth_pixels_over_1e-5"] = int((d_err > 1e-5).sum())
    if tight:
        # the timed path bins only the tiles the {alpha >= 1/255} box reaches: a sub-list of the list just compared, and
        # the same image bit for bit; the gradients below are those of THIS path
        del st
        lit_color, lit_depth, lit_radii = color, depth, radii
        n, color, radii, geom, binning, img, depth = fwd(flags)
        out["tight_binning"] = {"num_rendered": int(n), "fraction_of_literal_list": float(n) / max(int(f.num_rendered), 1),
                                "colour_bit_identical_to_literal": bool(torch.equal(color, lit_color)),
                                "depth_bit_identical_to_literal": bool(torch.equal(depth, lit_depth)),
                                "radii_equal": bool(torch.equal(radii, lit_radii))}
    if tile_step == 1:
        grads = _C.rasterize_gaussians_backward(rs.bg, gauss["means3D"], radii, e, gauss["scales"], gauss["rotations"], rs.scale_modifier,
                                                e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, wt, gauss["shs"], rs.sh_degree,
                                                rs.campos, geom, n, binning, img, flags=flags)
        names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
        rel = {}
        for name, t in zip(names, grads):
            if name in ("dL_dcolors", "dL_dcov3D"):
                continue
            a, b = t.cpu().numpy().astype(np.float64), g_ref[name].astype(np.float64)
            rel[name] = float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
        out["grad_rel_err"] = rel
        out["grad_ok_1e-3"] = bool(max(rel.values()) < 1e-3)
    out["ok"] = bool(out["radii_equal"] and out["point_list_equal"] and out["ranges_equal"] and out["colour_max_abs_err"] < 5e-3
                     and out["colour_pixels_over_1e-5"] <= max(2, int(1e-4 * out["pixels_compared"]))
                     and out.get("grad_ok_1e-3", True)
                     and all(v for k, v in out.get("tight_binning", {}).items() if k not in ("num_rendered", "fraction_of_literal_list")))
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from multiview_inpaint_b200 import _C, multiview as mv, scenes as S
    from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    own_pg = False
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        if not dist.is_initialized():   # tools/exp_configs_multi.py runs several configurations in one process group
            dist.init_process_group("nccl", device_id=dev)
            own_pg = True

    cfg = dict(S.CONFIGS[args.workload])
    sc = S.make_config_scene(args.workload)
    P, W, H, M, D = sc["P"], sc["W"], sc["H"], sc["M"], sc["sh_degree"]
    G = ((W + 15) // 16) * ((H + 15) // 16)
    gauss = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    vpr = args.views_per_rank
    strong = args.total_views > 0
    n_views = args.total_views if strong else vpr * world
    cams_cpu = S.orbit_cameras(n_views, W, H, max_deg=args.orbit_deg)
    mine = mv.shard_views(n_views, rank, world)
    bg = torch.zeros(3, device=dev)
    balance = None
    if strong and world > 1 and not args.no_balance:
        # Views of an orbit differ in cost (num_rendered): learn every view's N with the round-robin map, then give each
        # rank a set of views of about equal total N (longest-processing-time first).  SURVEY 8e.
        e = torch.empty(0, device=dev)
        n_of = torch.zeros(n_views, dtype=torch.int64, device=dev)
        for v in mine:
            c = cams_cpu[v].to(dev)
            n_of[v] = _C.rasterize_gaussians(bg, gauss["means3D"], e, gauss["opacities"], gauss["scales"], gauss["rotations"], 1.0, e,
                                             c.world_view_transform, c.full_proj_transform, c.tanfovx, c.tanfovy, H, W, gauss["shs"],
                                             D, c.camera_center, False, flags=args.flags, capacity=0)[0]
        dist.all_reduce(n_of)
        n_list = [int(x) for x in n_of.tolist()]
        loads, assign = [0] * world, [[] for _ in range(world)]
        for v in sorted(range(n_views), key=lambda v: -n_list[v]):
            r = min(range(world), key=lambda r: (loads[r], len(assign[r])))
            assign[r].append(v)
            loads[r] += n_list[v]
        rr_loads = [sum(n_list[v] for v in mv.shard_views(n_views, r, world)) for r in range(world)]
        mine = sorted(assign[rank])
        balance = {"views_per_rank": [len(a) for a in assign], "instances_max_over_mean": max(loads) / (sum(loads) / world),
                   "round_robin_max_over_mean": max(rr_loads) / (sum(rr_loads) / world)}
        torch.cuda.empty_cache()

    def settings(cam):
        return GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                             bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
                                             projmatrix=cam.full_proj_transform, sh_degree=D, campos=cam.camera_center,
                                             prefiltered=False)

    # device-resident inputs (for `value`) and pinned host copies (for `e2e`)
    cams_dev = {v: cams_cpu[v].to(dev) for v in mine}
    wts_cpu = {v: S.loss_weights(W, H, cfg["seed"] + v).pin_memory() for v in mine}
    wts_dev = {v: wts_cpu[v].to(dev) for v in mine}
    cam_pinned = {}
    for v in mine:
        c = cams_cpu[v]
        buf = torch.cat([c.world_view_transform.reshape(-1), c.full_proj_transform.reshape(-1), c.camera_center.reshape(-1)]).pin_memory()
        cam_pinned[v] = buf
    arena = mv.GradArena(P, M, dev, symmetric=(args.allreduce != "nccl"))
    comm = {"method": "none"}
    if world > 1:
        if args.allreduce == "auto":
            comm = arena.calibrate()
        else:
            arena.method = args.allreduce if arena._mc else "nccl"
            comm = {"method": arena.method}
    stats = {}
    av = mv.AsyncViews(n_views)

    def step_learn():
        """first step: exact-size path, learns each view's num_rendered for the capacity hints"""
        arena.zero_()
        for v in mine:
            wt = wts_dev[v]
            r = mv.cuda_view_fwd_bwd(gauss, settings(cams_dev[v]), lambda c, wt=wt: wt, arena, flags=args.flags, capacity=0)
            stats["N"], stats["radii"] = r.num_rendered, r.radii
            av.learn(v, r.num_rendered)
        arena.all_reduce()

    pipe = mv.ViewPipeline(dev, depth=args.streams) if args.streams > 1 else None
    workspaces = [_C.Workspace(dev) for _ in mine]   # per-view scratch + outputs, reused every step (no allocator traffic)
    throttle = mv.StepThrottle(2)   # host at most two steps ahead of the GPU (bounded scratch; see StepThrottle)
    import contextlib

    def step_resident(pipe=pipe):
        """steady state: nothing blocks the host (GSR_FLAG_ASYNC); overflow is checked after the timed region.
        Per view K1..K7 (two views in flight), then ONE batched K8+K9 that writes the arena, then the all-reduce."""
        if args.per_view_backward:
            arena.zero_()
            with (pipe.step() if pipe else contextlib.nullcontext()):
                for v in mine:
                    wt = wts_dev[v]
                    mv.cuda_view_fwd_bwd(gauss, settings(cams_dev[v]), lambda c, wt=wt: wt, arena, flags=args.flags,
                                         capacity=av.capacity(v), async_result=av.slot(v), pipeline=pipe)
        else:
            mv.cuda_views_fwd_bwd(gauss, [settings(cams_dev[v]) for v in mine], [lambda c, wt=wts_dev[v]: wt for v in mine],
                                  arena, flags=args.flags, capacities=[av.capacity(v) for v in mine],
                                  async_results=[av.slot(v) for v in mine], pipeline=pipe, all_reduce=True, chunks=args.ar_chunks,
                                  workspaces=workspaces, batched=not args.no_batched)
        if args.per_view_backward:
            arena.all_reduce()
        throttle.tick(dev)

    n_slots = max(args.streams, 1)
    cam_stage = {v: torch.empty(35, device=dev) for v in mine}   # per view: the batched K8+K9 reads every view's camera at the end
    wt_stage = {v: torch.empty(3, H, W, device=dev) for v in mine}
    loss_parts = [torch.zeros((), device=dev) for _ in range(n_slots)]
    # Host -> device staging on its own stream (copy engine): a view's camera is needed when its K1 starts, its loss
    # weights ("GT image", 19 MB) only after its forward -- issued at the top of the step, the copies run under the
    # first views' kernels instead of in front of them.  All cameras first (140 B each), then the images.
    copy_stream = torch.cuda.Stream(dev)
    cam_done = {v: torch.cuda.Event() for v in mine}
    wt_done = {v: torch.cuda.Event() for v in mine}

    def step_e2e():
        for lp in loss_parts:
            lp.zero_()
        copy_stream.wait_stream(torch.cuda.current_stream(dev))   # last step's readers of the staging buffers are done
        with torch.cuda.stream(copy_stream):
            for v in mine:
                cam_stage[v].copy_(cam_pinned[v], non_blocking=True)
                cam_done[v].record(copy_stream)
            for v in mine:
                wt_stage[v].copy_(wts_cpu[v], non_blocking=True)
                wt_done[v].record(copy_stream)
        stages, grads = [], []
        if True:
            for v in mine:
                def stage(v=v):   # runs on the view's stream
                    torch.cuda.current_stream(dev).wait_event(cam_done[v])
                    c = cams_cpu[v]
                    return GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy,
                                                         bg=bg, scale_modifier=1.0, viewmatrix=cam_stage[v][:16].view(4, 4),
                                                         projmatrix=cam_stage[v][16:32].view(4, 4), sh_degree=D,
                                                         campos=cam_stage[v][32:35], prefiltered=False)

                def loss_grad(col, v=v):
                    k = pipe.slot if pipe else 0
                    torch.cuda.current_stream(dev).wait_event(wt_done[v])
                    loss_parts[k].add_((col * wt_stage[v]).sum())
                    return wt_stage[v]
                stages.append(stage)
                grads.append(loss_grad)
        if args.per_view_backward:
            arena.zero_()
            with (pipe.step() if pipe else contextlib.nullcontext()):
                for k, v in enumerate(mine):
                    mv.cuda_view_fwd_bwd(gauss, stages[k], grads[k], arena, flags=args.flags,
                                         capacity=av.capacity(v), async_result=av.slot(v), pipeline=pipe)
        else:
            mv.cuda_views_fwd_bwd(gauss, stages, grads, arena, flags=args.flags, capacities=[av.capacity(v) for v in mine],
                                  async_results=[av.slot(v) for v in mine], pipeline=pipe, all_reduce=True, chunks=args.ar_chunks,
                                  workspaces=workspaces, batched=not args.no_batched)
        if args.per_view_backward:
            arena.all_reduce()
        # D2H: the step's result.  The loss lands in a pinned ring slot (asynchronous copy + event) and the host reads
        # it ONE step later, after the next step has been queued -- what a training loop that logs its loss does; a
        # blocking .item() here drains the queue and the GPU then idles while Python issues the next step's first
        # launches (the 8 % between `value` and `e2e` of round 1).  Every step's loss is read inside the timed region
        # (e2e_drain() before the closing event); the per-view (N, status) words are checked with the same lag.
        slot = e2e_state["step"] % len(loss_ring_ev)
        e2e_state["step"] += 1
        loss_ring[slot:slot + 1].copy_(torch.stack(loss_parts).sum().reshape(1), non_blocking=True)
        loss_ring_ev[slot].record()
        e2e_state["pending"].append(slot)
        out = None
        if len(e2e_state["pending"]) > 1 or args.e2e_blocking:
            out = e2e_read_one()
        return out

    loss_ring = torch.zeros(4, dtype=torch.float32).pin_memory()
    loss_ring_ev = [torch.cuda.Event() for _ in range(4)]
    e2e_state = {"step": 0, "pending": [], "losses": []}

    def e2e_read_one():
        s_ = e2e_state["pending"].pop(0)
        loss_ring_ev[s_].synchronize()
        val = float(loss_ring[s_])
        e2e_state["losses"].append(val)
        assert not av.check(mine), "capacity overflow inside the timed region"
        return val

    def e2e_drain():
        while e2e_state["pending"]:
            e2e_read_one()
    h2d = len(mine) * (35 * 4 + 3 * H * W * 4)
    d2h = 4 + len(mine) * 4   # loss scalar + num_rendered per view

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()   # host-side reads still owed by the loop (the last step's loss): inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up ----
    step_learn()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    assert not av.check(mine)
    if world > 1 and args.allreduce == "auto" and not args.per_view_backward:
        # Pick the exchange on the STEP it is part of, not on the collective alone: the in-switch reduction overlaps the chunked
        # per-Gaussian backward, an NCCL all-reduce follows it.  A few steps of every candidate (collective, Gaussian-range chunks,
        # CTAs of the reduction kernel), max over ranks, keep the fastest -- part of the warm-up, outside every timed region.
        cands = [("nccl", 1, 0, False)]
        if arena._mc:
            # tools/exp_scale8.py: the sweep these come from; the tapered ones (half-length first and last Gaussian range,
            # multiview.chunk_ranges) start the links earlier and shorten the reduction tail nobody overlaps
            cands += [("nvls", 4, 148, False), ("nvls", 4, 74, False), ("nvls", 8, 74, False), ("nvls", 4, 0, False),
                      ("nvls", 5, 74, True), ("nvls", 6, 74, True)]
            if world == 2 and arena._peer_ptr():
                # two ranks: peer loads / stores over NVLink move the same bytes at the link rate (csrc/nvls_allreduce.cu)
                cands += [("p2p", 4, 148, False), ("p2p", 4, 74, False), ("p2p", 6, 148, True), ("p2p", 1, 0, False)]
        trials = []
        for m_, c_, b_, tp_ in cands:
            arena.method, arena.nvls_blocks, args.ar_chunks, arena.taper = m_, b_, c_, tp_
            for _ in range(2):
                step_resident()
            t_ = timed(step_resident, 5) / 5
            trials.append({"method": m_, "chunks": c_, "nvls_blocks": b_, "taper": tp_, "ms_per_step": round(t_, 4)})
        best = min(trials, key=lambda t: t["ms_per_step"])
        arena.method, arena.nvls_blocks, args.ar_chunks, arena.taper = best["method"], best["nvls_blocks"], best["chunks"], best["taper"]
        comm = dict(comm, method=best["method"], chunks=best["chunks"], nvls_blocks=best["nvls_blocks"], taper=best["taper"], step_trials=trials)
        assert not av.check(mine)
    V = int((stats["radii"] > 0).sum().item())
    N = int(stats["N"])

    # ---- timed: resident (the headline `value`) ----
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _C.kernel_launches()
    a0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    ms_total = timed(step_resident, args.steps)
    launches = _C.kernel_launches() - l0
    device_allocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - a0   # cudaMalloc calls inside the timed region (want 0)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert not av.check(mine), "capacity overflow inside the timed region"

    # ---- the same loop once more with the stage profiler on: CUDA events on the launching stream
    #      around every kernel stage (events between kernels cost a few %, so it is a separate pass) ----
    step_resident(None)            # warm the caching allocator's pool of the main stream (the timed loop used the pipeline's)
    torch.cuda.synchronize()
    _C.profile_enable(True)
    _C.profile_collect()
    ms_prof = timed(lambda: step_resident(None), args.steps)   # one stream: stage brackets must not overlap
    _C.profile_enable(False)
    stage_ms, stage_cnt = _C.profile_collect()

    # ---- timed: end to end ----
    for _ in range(2):
        step_e2e()
    e2e_drain()
    n_before = len(e2e_state["losses"])
    ms_e2e = timed(step_e2e, args.steps, finish=e2e_drain)
    assert len(e2e_state["losses"]) - n_before == args.steps, "every timed e2e step's loss must have been read by the host"

    # ---- ablation baseline on the same device and data (N = 1 only): kernels with the STRUCTURE of the public
    #      rasterizer the reference depends on (GSR_FLAG_REFERENCE: host round trip for N, one 64-bit cub sort,
    #      thread-per-pixel blend, 9 atomics per pair), one view at a time on one stream, arena zeroed per step.
    #      A stand-in: the reference's own extension is not in its tree and cannot be built offline. ----
    # ---- the DROP-IN path itself: GaussianRasterizer(...)(...) + loss.backward() one view at a time on one stream,
    #      exactly as gs-simp/train.py:86-93 drives it through gaussian_renderer/__init__.py:85-93 (autograd Function,
    #      allocator-owned buffers, capacity speculation inside the binding) -- what an UNCHANGED train.py sees. ----
    dropin = None
    if world == 1 and not args.no_dropin:
        from multiview_inpaint_b200.rasterizer import GaussianRasterizer
        leaves = {k: gauss[k].detach().clone().requires_grad_(True) for k in gauss}
        sets = [settings(cams_dev[v]) for v in mine]

        def step_dropin():
            for rs, v in zip(sets, mine):
                means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
                color, radii, depth = GaussianRasterizer(raster_settings=rs)(
                    means3D=leaves["means3D"], means2D=means2D, shs=leaves["shs"], colors_precomp=None,
                    opacities=leaves["opacities"], scales=leaves["scales"], rotations=leaves["rotations"], cov3D_precomp=None)
                (color * wts_dev[v]).sum().backward()
            for t in leaves.values():
                t.grad = None
        for _ in range(3):
            step_dropin()
        dsteps = max(3, min(args.steps, 20))
        ms_d = timed(step_dropin, dsteps)
        dropin = {"value": len(mine) * dsteps / (ms_d / 1000.0), "unit": UNIT, "ms_per_view": ms_d / (len(mine) * dsteps), "steps": dsteps,
                  "note": "GaussianRasterizer + loss.backward(), one view per call, default stream, as gs-simp/train.py:86-93 drives the "
                          "drop-in; includes torch autograd, the (color * w).sum() loss and the per-call host read of num_rendered"}
        del leaves

    ref_struct = None
    if world == 1 and not args.no_reference_structure:
        def step_refstruct():
            arena.zero_()
            for v in mine:
                wt = wts_dev[v]
                mv.cuda_view_fwd_bwd(gauss, settings(cams_dev[v]), lambda c, wt=wt: wt, arena, flags=_C.FLAG_REFERENCE, capacity=0)
        for _ in range(2):
            step_refstruct()
        rs_steps = max(2, min(args.steps, 5))
        ms_rs = timed(step_refstruct, rs_steps)
        ref_struct = {"value": n_views * rs_steps / (ms_rs / 1000.0), "unit": UNIT, "ms_per_view": ms_rs / (n_views * rs_steps),
                      "steps": rs_steps, "flags": _C.FLAG_REFERENCE,
                      "note": "reference-STRUCTURE stand-in kernels of this repo (csrc/refstruct.cu), not the reference's binary"}


    # ---- the whole optimisation step (SURVEY 8f rows 1 + 4), N = 1: getters -> rasterizer -> L1+SSIM loss -> backward ->
    #      Adam, fused (GaussianParamArena / ViewLoss / one Adam launch) against the torch structure the reference runs
    #      around the same rasterizer (torch getters + cat, F.conv2d SSIM, autograd, torch.optim.Adam over six groups;
    #      gs-simp/train.py:86-128).  Both do `views_per_rank` views and ONE optimizer step per iteration. ----
    train_step = None
    if world == 1 and not args.no_train_step:
        try:
            train_step = bench_train_step(args, torch, _C, mv, S, sc, gauss, [settings(cams_dev[v]) for v in mine], dev, timed,
                                          av, mine, pipe, workspaces)
        except Exception as ex:   # an extra leg: never a reason to lose the bench line
            train_step = {"failed": repr(ex)}

    views_total = n_views * args.steps
    value = views_total / (ms_total / 1000.0)
    e2e_value = views_total / (ms_e2e / 1000.0)

    if rank != 0:
        if world > 1:
            if own_pg:
                dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peak, peak_kind = peaks()
    nv_shared = 1 if args.no_batched else min(len(mine), 8)      # views per batched launch (GSR_MAX_BATCH)
    alg = algorithmic_bytes(P, V, N, G, W, H, M, nv_shared, clear_in_k1=not args.no_batched)
    alg_survey = survey_bytes(P, V, N, G, W, H, M)
    n_view_steps = max(len(mine) * args.steps, 1)
    per_launch = {k: stage_ms[k] / max(stage_cnt[k], 1) for k in stage_ms}
    per_view = {k: stage_ms[k] / n_view_steps for k in stage_ms}   # batched stages launch once per step
    views_per_launch = {k: (n_view_steps / stage_cnt[k] if stage_cnt[k] else 1.0) for k in stage_ms}
    dom = max((k for k in per_launch if k in alg), key=lambda k: stage_ms[k])
    launch_bytes = alg[dom] * views_per_launch[dom]                # algorithmic bytes of ONE launch of the dominant kernel
    ach = launch_bytes / (per_launch[dom] / 1000.0) / 1e9 if per_launch[dom] > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get(dom)
            if traffic is not None:
                traffic = traffic * views_per_launch[dom]          # the capture is per view
        except Exception:
            traffic = None
    ctr = issue_counters().get(dom, {})
    issue_bound = dom.startswith("blend")
    b_view = sum(alg.values())
    ms_view = ms_total / views_total * world   # per-GPU time per view
    def stage_row(k):
        row = {"ms_per_view": round(per_view[k], 4), "ms_per_launch": round(per_launch[k], 4),
               "alg_bytes_per_view": (int(alg[k]) if k in alg else None), "survey_bytes_per_view": alg_survey.get(k),
               "gbps": None, "frac_hbm": None}
        if k in alg and per_view[k] > 0:
            g = alg[k] / (per_view[k] / 1000.0) / 1e9
            if g <= peak:
                row["gbps"], row["frac_hbm"] = round(g, 1), round(g / peak, 3)
            else:   # only the blend stages at shapes that saturate early: their model charges every binned instance,
                row["note"] = "byte model is an upper bound here (the walk stops at saturation): no bandwidth figure"   # the walk stops sooner
        if k in ctr_all:
            row["issue_active_frac"] = ctr_all[k].get("issue_active_frac")
        return row
    ctr_all = issue_counters() if args.workload == "headline" else {}
    stages = {k: stage_row(k) for k in per_launch}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **cfg, "views_per_rank": (None if strong else vpr), "views_per_step": n_views, "view_balance": balance, "orbit_deg": args.orbit_deg, "streams": args.streams, "batched_front_end": not args.no_batched, "batched_geom_backward": not args.per_view_backward, "ar_chunks": args.ar_chunks, "P": P, "V": V, "N": N,
                   "G": G, "M": M, "flags": args.flags, "parallelism": f"views sharded over {world} rank(s), fp32 grad-arena all-reduce", "allreduce": comm,
                   "allreduce_method": comm.get("method"),
                   "l2": "inputs (>= 700 MB of Gaussians per view) larger than the 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "alloc": {"cudaMalloc_in_timed_region": int(device_allocs), "reserved_gb": round(torch.cuda.memory_reserved(dev) / 2**30, 2)},
        "clocks": sampler.summary(),
        "profiled_ms_per_step": ms_prof / args.steps,
        # `achieved / peak / frac` are the contract's HBM figures for the dominant kernel (algorithmic bytes of one launch /
        # its CUDA-event time / the measured copy peak).  The blend kernels gather 48-byte records that the 126 MB L2 serves
        # (`traffic` << algorithmic bytes) and are bound by instruction ISSUE: their roof is the issue-slot utilisation
        # `issue_active_frac` measured by ncu, not the HBM fraction, which is low by construction.
        "roofline": {"bound": "issue" if issue_bound else "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic, "peak_kind": "of " + peak_kind, "alg_bytes_per_launch": int(launch_bytes),
                     "views_per_launch": views_per_launch[dom], "ms_per_launch": per_launch[dom],
                     "issue_active_frac": ctr.get("issue_active_frac"), "warp_inst_per_launch": ctr.get("warp_inst"),
                     "counters_from": ctr.get("source"),
                     "view": view_roofline(alg, N, W, H, ms_view, peak, survey=alg_survey)},
        "stages": stages,
        # SURVEY 8d timing protocol: forward, backward and forward+backward reported separately (stage-profiler pass,
        # one stream, per view)
        "split": {"fwd_ms_per_view": round(sum(per_view[k] for k in FWD_STAGES if k in per_view), 4),
                  "bwd_ms_per_view": round(sum(per_view[k] for k in BWD_STAGES if k in per_view), 4),
                  "fwd_bwd_ms_per_view_serial": round(sum(per_view[k] for k in FWD_STAGES + BWD_STAGES if k in per_view), 4),
                  "fwd_bwd_ms_per_view_timed_loop": round(ms_view, 4)},
    }
    if dropin is not None:
        line["dropin"] = dropin
    if ref_struct is not None:
        line["reference_structure"] = ref_struct
    if train_step is not None:
        line["train_step"] = train_step

    # ---- cpu_baseline: rank 0, N = 1 only ----
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle as O
            O.build()
            tile_step = args.cpu_tile_step or 1
            cpu_view_seconds(sc, cams_cpu[0], wts_cpu[mine[0]].numpy(), 8)          # warm-up (page-in, thread pool)
            ests, meas, parts, t_start = [], 0.0, {}, time.perf_counter()
            kept = {}
            while len(ests) < 8 and (time.perf_counter() - t_start < 15.0 or len(ests) < 2):
                v = mine[len(ests) % len(mine)]
                est, m, parts = cpu_view_seconds(sc, cams_cpu[v], wts_cpu[v].numpy(), tile_step, keep=(kept if not ests else None))
                ests.append(est)
                meas += m
            # the oracle has just computed view mine[0] in full: compare it with the CUDA path instead of throwing it away
            try:
                line["parity_headline"] = parity_against_oracle(torch, _C, gauss, settings(cams_dev[mine[0]]), wts_dev[mine[0]], kept,
                                                                P, W, H, tile_step, args.flags)
            except Exception as ex:
                line["parity_headline"] = {"failed": repr(ex)[:300]}
            est = sum(ests) / len(ests)
            line["cpu_baseline"] = {"value": 1.0 / est, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
                                    "sample": f"{len(ests)} full fwd+bwd view(s) of '{args.workload}' through the C oracle (OpenMP), blend on "
                                              f"every {tile_step}-th of {G} tiles" + (f" scaled x{tile_step}" if tile_step > 1 else "")
                                              + f"; {meas:.1f} s of CPU work measured",
                                    **parts}
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the bench line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(line), flush=True)
    if world > 1 and own_pg:
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------
# inference mode (BASELINE config 4: render.py + render_depth.py loops): forward only, views sharded, no collective
# -------------------------------------------------------------------------------------------------
def run_infer(args):
    import torch
    import torch.distributed as dist
    from multiview_inpaint_b200 import _C, multiview as mv, scenes as S
    from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    own_pg = False
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        if not dist.is_initialized():   # tools/exp_configs_multi.py runs several configurations in one process group
            dist.init_process_group("nccl", device_id=dev)
            own_pg = True
    cfg = dict(S.CONFIGS[args.workload])
    sc = S.make_config_scene(args.workload)
    P, W, H, M, D = sc["P"], sc["W"], sc["H"], sc["M"], sc["sh_degree"]
    G = ((W + 15) // 16) * ((H + 15) // 16)
    gauss = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    vpr = args.views_per_rank
    n_views = vpr * world
    cams_cpu = S.orbit_cameras(n_views, W, H, max_deg=args.orbit_deg)
    mine = mv.shard_views(n_views, rank, world)
    bg = torch.zeros(3, device=dev)
    cams_dev = {v: cams_cpu[v].to(dev) for v in mine}

    def settings(cam):
        return GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                             bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
                                             projmatrix=cam.full_proj_transform, sh_degree=D, campos=cam.camera_center,
                                             prefiltered=False)
    av = mv.AsyncViews(n_views)
    stats = {}
    for v in mine:   # exact-size pass: learns each view's num_rendered for the capacity hints
        r = mv.cuda_views_render(gauss, [settings(cams_dev[v])], flags=args.flags)[0]
        av.learn(v, r.num_rendered)
        stats["N"], stats["radii"] = r.num_rendered, r.radii
    pipe = mv.ViewPipeline(dev, depth=args.streams) if args.streams > 1 else None
    workspaces = [_C.Workspace(dev) for _ in mine]
    throttle = mv.StepThrottle(2)

    def step_resident(pipe=pipe):
        mv.cuda_views_render(gauss, [settings(cams_dev[v]) for v in mine], flags=args.flags,
                             capacities=[av.capacity(v) for v in mine], async_results=[av.slot(v) for v in mine],
                             pipeline=pipe, workspaces=workspaces, batched=not args.no_batched)
        throttle.tick(dev)

    # e2e: camera from pinned host memory in, colour + depth to pinned host memory out (what the PNG / NPY writer reads)
    cam_pinned = {v: torch.cat([cams_cpu[v].world_view_transform.reshape(-1), cams_cpu[v].full_proj_transform.reshape(-1),
                                cams_cpu[v].camera_center.reshape(-1)]).pin_memory() for v in mine}
    cam_stage = {v: torch.empty(35, device=dev) for v in mine}
    host_color = [torch.empty(3, H, W).pin_memory() for _ in mine]
    host_depth = [torch.empty(1, H, W).pin_memory() for _ in mine]

    def step_e2e():
        def stage(v):
            def f():
                cam_stage[v].copy_(cam_pinned[v], non_blocking=True)
                c = cams_cpu[v]
                return GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                                                     scale_modifier=1.0, viewmatrix=cam_stage[v][:16].view(4, 4),
                                                     projmatrix=cam_stage[v][16:32].view(4, 4), sh_degree=D,
                                                     campos=cam_stage[v][32:35], prefiltered=False)
            return f

        def sink(k, color, depth, radii):
            host_color[k].copy_(color, non_blocking=True)
            host_depth[k].copy_(depth, non_blocking=True)
        mv.cuda_views_render(gauss, [stage(v) for v in mine], flags=args.flags, capacities=[av.capacity(v) for v in mine],
                             async_results=[av.slot(v) for v in mine], pipeline=pipe, workspaces=workspaces, sink=sink,
                             batched=not args.no_batched)
        torch.cuda.current_stream(dev).synchronize()     # the step's result is on the host
        assert not av.check(mine), "capacity overflow inside the timed region"
    h2d = len(mine) * 35 * 4
    d2h = len(mine) * (4 * H * W * 4 + 8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    assert not av.check(mine)
    V, N = int((stats["radii"] > 0).sum().item()), int(stats["N"])
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _C.kernel_launches()
    ms_total = timed(step_resident, args.steps)
    launches = _C.kernel_launches() - l0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert not av.check(mine), "capacity overflow inside the timed region"
    step_resident(None)
    torch.cuda.synchronize()
    _C.profile_enable(True)
    _C.profile_collect()
    ms_prof = timed(lambda: step_resident(None), args.steps)
    _C.profile_enable(False)
    stage_ms, stage_cnt = _C.profile_collect()
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    views_total = n_views * args.steps
    if rank != 0:
        if world > 1:
            if own_pg:
                dist.destroy_process_group()
        return
    peak, peak_kind = peaks()
    alg = algorithmic_bytes(P, V, N, G, W, H, M, 1 if args.no_batched else min(len(mine), 8))
    fwd_stages = ("preprocess", "depth_sort", "duplicate", "tile_sort", "tile_ranges", "blend_forward")
    per_launch = {k: stage_ms[k] / max(stage_cnt[k], 1) for k in stage_ms if stage_cnt[k] > 0}
    dom = max((k for k in per_launch if k in fwd_stages), key=lambda k: stage_ms[k])
    vpl = (len(mine) * args.steps) / max(stage_cnt[dom], 1)           # views per launch of the dominant stage
    ach = alg[dom] * vpl / (per_launch[dom] / 1000.0) / 1e9
    b_view = sum(alg.get(k, 0) for k in fwd_stages)
    ms_view = ms_total / views_total * world
    line = {
        "metric": "fwd views/s (colour + depth)", "value": views_total / (ms_total / 1000.0), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "mode": "infer", **cfg, "views_per_rank": vpr, "views_per_step": n_views,
                   "orbit_deg": args.orbit_deg, "streams": args.streams, "batched_front_end": not args.no_batched, "P": P, "V": V, "N": N, "G": G, "M": M, "flags": args.flags,
                   "parallelism": f"views sharded over {world} rank(s), no collective",
                   "l2": "inputs larger than the 126 MB L2; no flush needed"},
        "e2e": {"value": views_total / (ms_e2e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": sampler.summary(), "profiled_ms_per_step": ms_prof / args.steps,
        "roofline": {"bound": "issue" if dom.startswith("blend") else "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": None, "peak_kind": "of " + peak_kind, "alg_bytes_per_launch": int(alg[dom] * vpl),
                     "issue_active_frac": issue_counters().get(dom, {}).get("issue_active_frac"),
                     "ms_per_launch": per_launch[dom],
                     "view": view_roofline({k: v for k, v in alg.items() if k in fwd_stages}, N, W, H, ms_view, peak)},
        "stages": {k: {"ms_per_launch": round(per_launch[k], 4), "ms_per_view": round(stage_ms[k] / max(len(mine) * args.steps, 1), 4),
                       "alg_bytes_per_view": (int(alg[k]) if k in alg else None)} for k in per_launch},
    }
    to_disk = None
    if args.files > 0 and world == 1:
        # render.py:32-39 all the way to PNG files: batched render + AsyncImageWriter (GPU quantisation, 3 B/pixel D2H,
        # encode on worker threads) against the structure of the reference (per view: blocking float D2H, 8-bit
        # conversion on the CPU, PIL encode on the calling thread -- what torchvision.utils.save_image does).
        import shutil
        import tempfile
        import numpy as np
        from multiview_inpaint_b200.imagewriter import AsyncImageWriter
        try:
            from PIL import Image
        except Exception:
            Image = None
        tmp = tempfile.mkdtemp(prefix="gsr_png_")
        try:
            writer = AsyncImageWriter(dev, slots=2 * len(mine), workers=args.files, use_pil=True)   # same encoder as the sync leg

            def step_async():
                mv.cuda_views_render(gauss, [settings(cams_dev[v]) for v in mine], flags=args.flags,
                                     capacities=[av.capacity(v) for v in mine], async_results=[av.slot(v) for v in mine],
                                     pipeline=pipe, workspaces=workspaces,
                                     sink=lambda k, color, depth, radii: writer.submit_png(os.path.join(tmp, f"a{k:05d}.png"), color))
            step_async()
            writer.flush()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_async()
            writer.flush()
            t_async = time.perf_counter() - t0
            writer.close()

            def step_sync():
                for k, v in enumerate(mine):
                    r = mv.cuda_views_render(gauss, [settings(cams_dev[v])], flags=args.flags)[0]
                    arr = r.color.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8).numpy()
                    if Image is not None:
                        Image.fromarray(arr).save(os.path.join(tmp, f"s{k:05d}.png"))
                    else:
                        np.save(os.path.join(tmp, f"s{k:05d}.npy"), arr)
            step_sync()
            n_sync = max(1, min(args.steps, 3))
            t0 = time.perf_counter()
            for _ in range(n_sync):
                step_sync()
            t_sync = time.perf_counter() - t0
            to_disk = {"async_writer_views_per_s": len(mine) * args.steps / t_async, "writer_threads": args.files,
                       "sync_structure_views_per_s": len(mine) * n_sync / t_sync, "host_cores": os.cpu_count(),
                       "encoder": "PIL" if Image is not None else "zlib", "d2h_bytes_per_view": 3 * H * W,
                       "note": "wall clock, files on the box's /tmp; both legs are bound by PNG encoding on the host"}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    if to_disk is not None:
        line["to_disk"] = to_disk
    print(json.dumps(line), flush=True)
    if world > 1 and own_pg:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.mode == "infer":
        run_infer(a)
    else:
        run_ours(a)
