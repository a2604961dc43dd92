"""The training step around the rasterizer, fused -- opt-in; `render()` and the
drop-in `diff_gaussian_rasterization` API stay as they are.

What every training script of the reference does per iteration (gs-simp/train.py:86-128, same skeleton in
sds_train.py and inpaint_rec.py) around the rasterizer call:

    getters      get_scaling = exp(_scaling), get_rotation = normalize(_rotation), get_opacity = sigmoid(_opacity),
                 get_features = cat((_features_dc, _features_rest), dim=1)       scene/gaussian_model.py:95-115
    loss         (1 - l) * l1_loss(image, gt) + l * (1 - ssim(image, gt))        train.py:91-92, utils/loss_utils.py
    backward     autograd through all of the above                               train.py:93
    optimizer    Adam over six parameter groups, eps 1e-15                       train.py:127, gaussian_model.py:154-165

Here:
  * `GaussianParamArena`  raw parameters, Adam moments and activated values in flat allocations laid out exactly
                          like `multiview.GradArena`; `_features_dc` / `_features_rest` are views of ONE (P,M,3)
                          tensor, so `get_features` is that tensor (no cat, no split in backward);
  * `activate()`          one kernel per STEP (the activated values are shared by all views of a multi-view step);
  * `l1_ssim_loss()`      autograd Function over the two fused kernels of csrc/loss.cu;
  * `apply_gradients()`   the activations' chain rule in place in the gradient arena + ONE Adam launch for all groups.
  * `densify_and_prune()` / `prune_points()` / `reset_opacity()`  the model-size changes of train.py:117-123 on the
                          arenas: one row-gather launch (densify.py, csrc/densify.cu);
  * `fused_train_step()`  the whole iteration for a batch of views: activate -> per view (K1..K6, loss fwd+bwd,
                          K7) -> batched K8+K9 (-> all-reduce) -> chain rule -> Adam.
Everything runs on the CUDA library (include/gsrast_b200.h); there is no torch fallback.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import _C
from .multiview import GradArena, ViewPipeline, cuda_views_fwd_bwd

#: the six groups of gaussian_model.py:154-161, in arena order; f_dc and f_rest share the (P,M,3) slice
GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def _numel(shp):
    n = 1
    for s in shp:
        n *= s
    return n


def expon_lr_func(lr_init: float, lr_final: float, lr_delay_steps: int = 0, lr_delay_mult: float = 1.0,
                  max_steps: int = 1000000):
    """The position learning-rate schedule of the reference (`get_expon_lr_func`, utils/general_utils.py:31-64, applied
    every iteration by `update_learning_rate`, gaussian_model.py:169-175, train.py:68): log-linear interpolation from
    lr_init (step 0) to lr_final (step max_steps), optionally eased in over lr_delay_steps.  Returns step -> lr; feed
    it to `fused_train_step` as lrs["xyz"]."""
    import numpy as np

    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
        else:
            delay_rate = 1.0
        t = np.clip(step / max_steps, 0, 1)
        return delay_rate * np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
    return helper


def learning_rates(iteration: int, xyz_schedule, feature_lr: float = 0.0025, opacity_lr: float = 0.05,
                   scaling_lr: float = 0.005, rotation_lr: float = 0.001) -> dict:
    """The six group learning rates of `training_setup` (gaussian_model.py:154-161; defaults arguments/__init__.py:83-86)
    at `iteration`, in the form `fused_train_step` / `apply_gradients` take them."""
    return dict(xyz=float(xyz_schedule(iteration)), f_dc=feature_lr, f_rest=feature_lr / 20.0, opacity=opacity_lr,
                scaling=scaling_lr, rotation=rotation_lr)


class GaussianParamArena:
    """Raw (pre-activation) Gaussian parameters in ONE flat fp32 allocation
        [ _xyz (P,3) | _features (P,M,3) | _opacity (P,1) | _scaling (P,3) | _rotation (P,4) ]
    with every slice starting on a 16-byte boundary -- slice for slice the layout of `GradArena.flat`, so a
    gradient slice, its parameter slice and its two Adam moment slices have identical offsets.  Attribute names
    follow GaussianModel (scene/gaussian_model.py:43-57)."""

    def __init__(self, P: int, M: int, device):
        self.P, self.M = int(P), int(M)
        self.device = torch.device(device)
        sizes = [("_xyz", (P, 3)), ("_features", (P, M, 3)), ("_opacity", (P, 1)), ("_scaling", (P, 3)),
                 ("_rotation", (P, 4))]
        offs, off = {}, 0
        for name, shp in sizes:
            offs[name] = off
            off += (_numel(shp) + 3) // 4 * 4
        self._offs, self._shapes, self.n_flat = offs, dict(sizes), off
        z = lambda: torch.zeros(max(off, 4), dtype=torch.float32, device=self.device)
        self.param, self.exp_avg, self.exp_avg_sq = z(), z(), z()
        self.step_count = 0
        # activated values handed to the rasterizer (what render() gets from the getters)
        self.scales = torch.empty(P, 3, dtype=torch.float32, device=self.device)
        self.rotations = torch.empty(P, 4, dtype=torch.float32, device=self.device)
        self.opacities = torch.empty(P, 1, dtype=torch.float32, device=self.device)
        for name, shp in sizes:
            setattr(self, name, self._slice(self.param, name))
        self._features_dc = self._features[:, :1, :]      # views: gaussian_model.py:142-143 keeps them as two tensors
        self._features_rest = self._features[:, 1:, :]
        self.max_sh_degree = int(round(self.M ** 0.5)) - 1   # M = (max_sh_degree + 1)^2
        self.active_sh_degree = 0                            # gaussian_model.py:45; raised by oneupSHdegree()

    def _slice(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        shp = self._shapes[name]
        return flat[self._offs[name]:self._offs[name] + _numel(shp)].view(*shp)

    @classmethod
    def from_tensors(cls, xyz, features_dc, features_rest, opacity, scaling, rotation) -> "GaussianParamArena":
        """From GaussianModel's six parameter tensors (gaussian_model.py:140-145): (P,3), (P,1,3), (P,M-1,3), (P,1),
        (P,3), (P,4), all RAW (pre-activation)."""
        P, M = xyz.shape[0], 1 + features_rest.shape[1]
        a = cls(P, M, xyz.device)
        with torch.no_grad():
            a._xyz.copy_(xyz)
            a._features[:, :1].copy_(features_dc)
            if M > 1:
                a._features[:, 1:].copy_(features_rest)
            a._opacity.copy_(opacity.reshape(P, 1))
            a._scaling.copy_(scaling)
            a._rotation.copy_(rotation)
        return a

    def oneupSHdegree(self):
        """gaussian_model.py:120-122 (train.py:71-72: every 1000 iterations): the degree handed to the rasterizer as
        `sh_degree` grows up to max_sh_degree; the (P,M,3) tensor keeps its full size throughout."""
        if self.active_sh_degree < self.max_sh_degree:
            self.active_sh_degree += 1

    @classmethod
    def create_from_pcd(cls, points, colors, sh_degree: int, device, dist2: torch.Tensor | None = None) -> "GaussianParamArena":
        """`GaussianModel.create_from_pcd` (gaussian_model.py:124-147): the initial model from an SfM point cloud --
        positions = points, DC coefficient = RGB2SH(colour) (utils/sh_utils.py:114-115), higher SH zero, isotropic
        scale = log sqrt(mean squared distance to the three nearest neighbours) from `distCUDA2` (here
        gsr_knn3_mean_dist2, csrc/knn.cu) clamped at 1e-7, identity rotations, opacity inverse_sigmoid(0.1).
        `points` / `colors`: (N,3) arrays or tensors (pcd.points / pcd.colors, colours in 0..1).  `dist2` lets a caller
        supply the neighbour distances (tests do, to run the host logic without a GPU); by default they come from the
        CUDA kernel, which needs a CUDA `device`."""
        import numpy as np
        device = torch.device(device)
        pts = torch.as_tensor(np.asarray(points)).float().to(device)
        rgb = torch.as_tensor(np.asarray(colors)).float().to(device)
        N, M = pts.shape[0], (sh_degree + 1) ** 2
        C0 = 0.28209479177387814
        fused_color = (rgb - 0.5) / C0                                                # RGB2SH
        if dist2 is None:
            dist2 = _C.dist_cuda2(pts.contiguous())                                   # raises on CPU tensors: no fallback
        dist2 = torch.clamp_min(dist2.to(device), 0.0000001)
        scales = torch.log(torch.sqrt(dist2))[..., None].repeat(1, 3)
        rots = torch.zeros((N, 4), device=device)
        rots[:, 0] = 1
        x = 0.1 * torch.ones((N, 1), dtype=torch.float, device=device)
        opacities = torch.log(x / (1 - x))                                            # inverse_sigmoid
        f_dc = fused_color.reshape(N, 1, 3)                                           # features[:, :3, 0] -> (N, 1, 3)
        f_rest = torch.zeros((N, M - 1, 3), device=device)
        a = cls.from_tensors(pts, f_dc, f_rest, opacities, scales, rots)
        a.active_sh_degree = 0
        return a

    # the getters of gaussian_model.py:95-115
    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_features(self):
        return self._features

    def activate(self) -> dict:
        """exp / normalize / sigmoid in one kernel -> the dict the multi-view entry points take."""
        _C.activate_forward(self._scaling, self._rotation, self._opacity, self.scales, self.rotations, self.opacities)
        return {"means3D": self._xyz, "shs": self._features, "opacities": self.opacities, "scales": self.scales,
                "rotations": self.rotations}

    def moments(self, name: str):
        return self._slice(self.exp_avg, name), self._slice(self.exp_avg_sq, name)

    def apply_gradients(self, arena: GradArena, lrs: dict, beta1: float = 0.9, beta2: float = 0.999,
                        eps: float = 1e-15):
        """`arena` holds dL/d(means3D, shs, opacities, scales, rotations) -- gradients with respect to the ACTIVATED
        values, as the rasterizer's backward produces them.  Runs the activations' chain rule in place in the arena
        (after which it holds what autograd would put in `.grad` of the six raw parameters), then one Adam launch.
        `lrs`: learning rate per group of GROUPS (the xyz one is scheduled by the caller, gaussian_model.py:167-173)."""
        assert arena.P == self.P and arena.M == self.M
        g = arena.views
        _C.activate_backward(self._scaling, self._rotation, self._opacity, g["dL_dscales"], g["dL_drotations"],
                             g["dL_dopacity"])
        self.step_count += 1
        pairs = (("_xyz", "dL_dmeans3D", dict(lr=lrs["xyz"])),
                 ("_features", "dL_dsh", dict(lr=lrs["f_dc"], lr_rest=lrs["f_rest"], row_len=3 * self.M, row_split=3)),
                 ("_opacity", "dL_dopacity", dict(lr=lrs["opacity"])),
                 ("_scaling", "dL_dscales", dict(lr=lrs["scaling"])),
                 ("_rotation", "dL_drotations", dict(lr=lrs["rotation"])))
        segs = []
        for pname, gname, kw in pairs:
            m, v = self.moments(pname)
            segs.append(dict(param=getattr(self, pname), grad=g[gname], exp_avg=m, exp_avg_sq=v, **kw))
        _C.adam_step(segs, self.step_count, beta1, beta2, eps)


    # ---- model-size changes (gaussian_model.py:263-266, :365-383, :467-480), see densify.py ----
    def densify_and_prune(self, stats, max_grad: float, min_opacity: float, extent: float, max_screen_size,
                          percent_dense: float = 0.01, generator=None, sync_ranks: bool = True):
        """`gaussians.densify_and_prune(opt.densify_grad_threshold, 0.005, scene.cameras_extent, size_threshold)`
        (train.py:120) from the statistics in `stats` (grad_norm_accum = xyz_gradient_accum, visible_count = denom):
        a `densify.DensificationStats` accumulated over the steps since the last densification (train.py:115-116),
        or a step's `GradArena`.  Clone / split / prune are composed into one index map and the parameter arena and
        both moment arenas move in ONE gather launch.  In place; returns the DensifyPlan (counts for logging).
        The model size changes: the caller replaces `stats` and its GradArena by `.resized(self.P)` (zeroed
        statistics, which is what densification_postfix leaves, :423-425) and drops per-view workspaces / capacity
        hints sized for the old P.  Multi-rank: every rank holds the same model and the same all-reduced statistics;
        the split's random samples are made identical by broadcasting rank 0's generator state first (`sync_ranks`)."""
        from . import densify
        assert stats.P == self.P
        if sync_ranks:
            densify.sync_rng(self.device, generator, getattr(stats, "group", None))
        plan = densify.plan_densify_and_prune(self._xyz, self._scaling, self._rotation, self._opacity,
                                              stats.grad_norm_accum, stats.visible_count, max_grad, min_opacity, extent,
                                              max_screen_size, percent_dense=percent_dense, generator=generator)
        densify.apply_plan(self, plan)
        return plan

    def prune_points(self, mask: torch.Tensor):
        """`gaussians.prune_points(mask)` (:365-383): Gaussians with mask == True are removed, order and optimizer
        state of the rest kept.  The caller carries the statistics over with `stats.pruned(mask)`."""
        from . import densify
        plan = densify.plan_prune(mask.to(self.device))
        assert plan.n_src == self.P
        densify.apply_plan(self, plan)
        return plan

    def reset_opacity(self):
        """`gaussians.reset_opacity()` (:263-266, train.py:122-123): opacities capped at 0.01, their Adam moments
        zeroed (replace_tensor_to_optimizer :331-344).  P floats: the reference's own torch expression."""
        from . import densify
        with torch.no_grad():
            self._opacity.copy_(densify.reset_opacity_values(self._opacity))
            for m in self.moments("_opacity"):
                m.zero_()


class _L1SSIMLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        image_c, gt_c = image.contiguous(), gt.contiguous()
        out3, temp = _C.loss_l1_ssim_forward(image_c, gt_c, lambda_dssim)
        ctx.save_for_backward(image_c, gt_c, temp)
        ctx.lambda_dssim = lambda_dssim
        l1, ss, loss = (t.clone() for t in out3.unbind(0))
        ctx.mark_non_differentiable(l1, ss)
        return loss, l1, ss

    @staticmethod
    def backward(ctx, g_loss, _g_l1, _g_ss):
        image, gt, temp = ctx.saved_tensors
        g = _C.loss_l1_ssim_backward(image, gt, ctx.lambda_dssim, temp, dL_dloss=g_loss.contiguous())
        return g, None, None


def l1_ssim_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float = 0.2):
    """-> (loss, Ll1, ssim), 0-dim CUDA tensors; `loss` = (1 - l) * l1_loss(image, gt) + l * (1 - ssim(image, gt))
    (train.py:91-92) and is differentiable with respect to `image`; Ll1 is what training_report logs (train.py:107)."""
    return _L1SSIMLoss.apply(image, gt, float(lambda_dssim))


class ViewLoss:
    """dL_dcolor callback of the multi-view entry points for one view: fused loss forward + backward against `gt`,
    with preallocated temp / output buffers (no allocator traffic in the loop).  After the step, `out3` holds
    {Ll1, ssim, loss} of the view on the device.  `weight`: dL/dloss seed (1/len(views) averages the batch)."""

    def __init__(self, gt: torch.Tensor, lambda_dssim: float = 0.2, weight: float = 1.0):
        self.gt = gt.contiguous()
        self.lambda_dssim = float(lambda_dssim)
        dev = gt.device
        self.temp = torch.empty(_C.loss_temp_bytes(*gt.shape), dtype=torch.uint8, device=dev)
        self.out3 = torch.zeros(3, dtype=torch.float32, device=dev)
        self.dL = torch.empty_like(self.gt)
        self.seed = None if weight == 1.0 else torch.full((1,), float(weight), dtype=torch.float32, device=dev)

    def __call__(self, color: torch.Tensor) -> torch.Tensor:
        _C.loss_l1_ssim_forward(color, self.gt, self.lambda_dssim, out3=self.out3, temp=self.temp)
        return _C.loss_l1_ssim_backward(color, self.gt, self.lambda_dssim, self.temp, dL_dloss=self.seed, out=self.dL)


def fused_train_step(params: GaussianParamArena, settings_list: Sequence, losses: Sequence[ViewLoss], arena: GradArena,
                     lrs: dict, flags: int | None = None, pipeline: ViewPipeline | None = None, all_reduce: bool = False,
                     capacities=None, async_results=None, workspaces=None, chunks: int = 4, apply: bool | None = None,
                     batched: bool = False):
    """One optimisation step on this rank's views (train.py:86-128 for a batch of views; with `all_reduce` the
    gradient arena is summed over ranks first).  Returns the per-view ViewState list; the densification
    statistics of the step are in `arena` (grad_norm_accum / visible_count / max_radii).

    `apply`: whether Adam runs inside this call.  Default: yes for the synchronous path, NO for the asynchronous one
    (`async_results`): there the host has not seen the views' (num_rendered, overflow) words yet, and a view that
    outgrew its binning capacity has a truncated instance list -- wrong image, wrong gradients -- which must not reach
    the Adam moments.  The asynchronous caller therefore finishes the step itself:

        states = fused_train_step(..., async_results=slots)          # gradients only, nothing blocks
        torch.cuda.current_stream().synchronize()                    # (or the step's natural sync point)
        if async_views.check(views):  <raise capacities, redo the step>   # the parameters are still untouched
        else:                         params.apply_gradients(arena, lrs)

    `apply=True` with `async_results` is allowed for callers that size the capacities so that an overflow cannot
    happen (bench.py: 25 % margin over a learnt high-water mark, checked after the timed region).

    Deviation from train.py:112-128, stated: the reference runs densify_and_prune / reset_opacity BEFORE
    optimizer.step(); on those iterations the re-created parameters have no .grad and Adam skips them.  Callers that
    need that trajectory pass apply=False on densification iterations, densify, and skip apply_gradients."""
    g = params.activate()
    states = cuda_views_fwd_bwd(g, settings_list, losses, arena, flags=flags, capacities=capacities,
                                async_results=async_results, pipeline=pipeline, all_reduce=all_reduce, chunks=chunks,
                                workspaces=workspaces, batched=batched)
    if apply is None:
        apply = async_results is None
    if apply:
        params.apply_gradients(arena, lrs)
    return states
