"""Densification for the fused training step: what `GaussianModel.densify_and_prune`,
`prune_points` and `reset_opacity` (gs-simp/scene/gaussian_model.py:467-480, :365-383, :263-266; called from
train.py:117-123, sds_train.py:156-163, inpaint_rec.py:151-158) do to the model, for `trainstep.GaussianParamArena`.

The reference changes the model size in four rounds -- clone (`densification_postfix` -> `cat_tensors_to_optimizer`),
split (the same again), removal of the split parents and the final prune (`prune_points` -> `_prune_optimizer`) --
and every round re-materialises all six parameter tensors plus their two Adam moments with `torch.cat` / mask
indexing: 72 whole-model copies per densification.  Here the four rounds are composed on INDICES only
(`plan_densify_and_prune`: a handful of P-sized torch ops, device-agnostic, checked on CPU against golden vectors
produced by the reference's own class, tests/golden/make_densify_golden.py), and the model is then moved ONCE:
`gsr_gather_rows` (csrc/densify.cu) writes every surviving row of the parameter arena and of both moment arenas in a
single launch.  The rows that are new values rather than copies -- the N = 2 children of a split Gaussian get a
sampled position and a shrunk scale (:433-440) -- are a few percent of the model and are written afterwards.

Row order of the result is the reference's: surviving originals (ascending), then clones, then split children.
Statistics (`xyz_gradient_accum`, `denom`, `max_radii2D`) come back zeroed, as `densification_postfix` leaves them
(:423-425) -- which also means the reference's `big_points_vs` test (:474) always sees zeros; reproduced, not fixed.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch


def build_rotation(r: torch.Tensor) -> torch.Tensor:
    """Raw quaternion (w,x,y,z) rows -> rotation matrices, in the operation order of
    gs-simp/utils/general_utils.py:80-101 (the split samples must land where the reference puts them)."""
    q = r / torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.zeros((q.size(0), 3, 3), dtype=r.dtype, device=r.device)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - w * z)
    R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y)
    R[:, 2, 1] = 2 * (y * z + w * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


@dataclass
class DensifyPlan:
    """Where every row of the densified model comes from."""
    src_row: torch.Tensor          # int32 (P_new,): row of the OLD model each new row copies
    n_src: int                     # P of the old model
    n_keep_state: int              # leading rows that keep their Adam moments (surviving originals)
    child_rows: torch.Tensor       # int64 (n_child,): new-model rows that are split children ...
    child_xyz: torch.Tensor        # (n_child, 3): ... their sampled positions (:439) ...
    child_scaling: torch.Tensor    # (n_child, 3): ... and raw (log) scales (:440)
    counts: dict = field(default_factory=dict)   # cloned / split / pruned, for logging

    @property
    def n_dst(self) -> int:
        return int(self.src_row.numel())


def plan_prune(mask: torch.Tensor) -> DensifyPlan:
    """`prune_points(mask)` (:365-383): rows with mask == True disappear, the rest keep order and optimizer state."""
    keep = (~mask.reshape(-1).bool()).nonzero(as_tuple=True)[0]
    e = torch.empty(0, 3, dtype=torch.float32, device=mask.device)
    return DensifyPlan(keep.to(torch.int32), int(mask.numel()), int(keep.numel()), keep[:0], e, e.clone(),
                       dict(cloned=0, split=0, pruned=int(mask.numel() - keep.numel())))


def plan_densify_and_prune(xyz: torch.Tensor, scaling_raw: torch.Tensor, rotation_raw: torch.Tensor,
                           opacity_raw: torch.Tensor, xyz_gradient_accum: torch.Tensor, denom: torch.Tensor,
                           max_grad: float, min_opacity: float, extent: float, max_screen_size,
                           percent_dense: float = 0.01, N: int = 2, generator=None) -> DensifyPlan:
    """`densify_and_prune(max_grad, min_opacity, extent, max_screen_size)` (:467-480) on indices.  Inputs are the RAW
    parameters (`_xyz`, `_scaling`, `_rotation`, `_opacity`) and the accumulated statistics
    (`GradArena.grad_norm_accum` / `visible_count`).  Device-agnostic torch; the only random draw is the reference's
    own `torch.normal(mean=0, std=scale)` of densify_and_split (:438), same shape and call, so a generator in the
    same state yields the same children."""
    P = xyz.shape[0]
    dev = xyz.device
    grads = xyz_gradient_accum.reshape(P, 1).float() / denom.reshape(P, 1).float()            # :468
    grads[grads.isnan()] = 0.0                                                                # :469
    scal = torch.exp(scaling_raw)                                                             # get_scaling
    smax = torch.max(scal, dim=1).values
    thr = percent_dense * extent
    # ---- densify_and_clone (:451-465): small Gaussians with a large view-space gradient are duplicated ----
    mask_c = torch.logical_and(torch.where(torch.norm(grads, dim=-1) >= max_grad, True, False), smax <= thr)
    idx_c = mask_c.nonzero(as_tuple=True)[0]
    n1 = P + idx_c.numel()
    src1 = torch.cat((torch.arange(P, device=dev), idx_c))                                    # old row of each of the n1 rows
    # ---- densify_and_split (:427-449): large ones are replaced by N samples of themselves ----
    padded_grad = torch.zeros(n1, device=dev)
    padded_grad[:P] = grads.reshape(-1)
    smax1 = smax[src1]
    mask_s = torch.logical_and(torch.where(padded_grad >= max_grad, True, False), smax1 > thr)
    par = src1[mask_s.nonzero(as_tuple=True)[0]]                                              # old rows of the parents
    stds = scal[par].repeat(N, 1)
    means = torch.zeros((stds.size(0), 3), device=dev)
    samples = torch.normal(mean=means, std=stds, generator=generator)
    rots = build_rotation(rotation_raw[par]).repeat(N, 1, 1)
    child_xyz = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + xyz[par].repeat(N, 1)
    child_scaling = torch.log(scal[par].repeat(N, 1) / (0.8 * N))
    keep1 = (~mask_s).nonzero(as_tuple=True)[0]                                               # parents leave (:448-449)
    n_child = par.numel() * N
    src2 = torch.cat((src1[keep1], par.repeat(N)))
    is_child = torch.cat((torch.zeros(keep1.numel(), dtype=torch.bool, device=dev),
                          torch.ones(n_child, dtype=torch.bool, device=dev)))
    from_original = torch.cat((keep1 < P, torch.zeros(n_child, dtype=torch.bool, device=dev)))
    # ---- final prune (:472-478) on the n2 = n1 - n_split + N n_split rows ----
    prune = (torch.sigmoid(opacity_raw.reshape(-1, 1)[src2]) < min_opacity).reshape(-1)
    if max_screen_size:
        max_radii2D = torch.zeros(src2.numel(), device=dev)                                   # densification_postfix :425
        big_vs = max_radii2D > max_screen_size
        smax2 = torch.cat((smax1[keep1], torch.exp(child_scaling).max(dim=1).values))
        big_ws = smax2 > 0.1 * extent
        prune = torch.logical_or(torch.logical_or(prune, big_vs), big_ws)
    keep2 = (~prune).nonzero(as_tuple=True)[0]
    child_sel = is_child[keep2]
    child_src = keep2[child_sel] - keep1.numel()
    return DensifyPlan(src2[keep2].to(torch.int32), P, int(from_original[keep2].sum()),
                       child_sel.nonzero(as_tuple=True)[0], child_xyz[child_src], child_scaling[child_src],
                       dict(cloned=int(idx_c.numel()), split=int(par.numel()), pruned=int(prune.sum())))


class DensificationStats:
    """`xyz_gradient_accum` / `denom` / `max_radii2D` of GaussianModel (gaussian_model.py:151-152, :66): the
    statistics summed over the iterations BETWEEN two densifications (train.py:115-116 every iteration, consumed at
    :120 every `densification_interval`).  `GradArena` holds the statistics of ONE step (they travel with the
    gradients through the all-reduce, so they are overwritten per step); `add_step` folds them in.  Same attribute
    names as GradArena, so either can be handed to `GaussianParamArena.densify_and_prune`."""

    def __init__(self, P: int, device, group=None):
        self.P, self.group = int(P), group
        self.grad_norm_accum = torch.zeros(P, dtype=torch.float32, device=device)
        self.visible_count = torch.zeros(P, dtype=torch.int32, device=device)
        self.max_radii = torch.zeros(P, dtype=torch.int32, device=device)

    def add_step(self, arena):
        """after a (multi-view) step, once its all-reduce is done: train.py:115-116 for all views of the step"""
        assert arena.P == self.P
        self.grad_norm_accum += arena.grad_norm_accum
        self.visible_count += arena.visible_count
        torch.maximum(self.max_radii, arena.max_radii, out=self.max_radii)

    def resized(self, P: int) -> "DensificationStats":
        """zeroed statistics for the densified model (densification_postfix, gaussian_model.py:423-425)"""
        return DensificationStats(P, self.grad_norm_accum.device, self.group)

    def pruned(self, mask: torch.Tensor) -> "DensificationStats":
        """prune_points carries the survivors' statistics over (gaussian_model.py:379-383)"""
        keep = ~mask.reshape(-1).bool().to(self.grad_norm_accum.device)
        new = DensificationStats(int(keep.sum()), self.grad_norm_accum.device, self.group)
        new.grad_norm_accum.copy_(self.grad_norm_accum[keep])
        new.visible_count.copy_(self.visible_count[keep])
        new.max_radii.copy_(self.max_radii[keep])
        return new


def sync_rng(device, generator=None, group=None) -> bool:
    """Every rank holds a replica of the model and the same all-reduced statistics, so every rank derives the same
    clone / split / prune masks; the one thing that could differ is the random draw of densify_and_split
    (gaussian_model.py:438).  Rank 0's generator state (the given generator, else the default generator of `device`)
    is broadcast so that all replicas sample identical children.  Returns True if a broadcast happened."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return False
    device = torch.device(device)
    cuda = device.type == "cuda"
    if generator is not None:
        state = generator.get_state()
    else:
        state = torch.cuda.get_rng_state(device) if cuda else torch.get_rng_state()
    state = state.to(device)
    dist.broadcast(state, 0, group=group)
    if generator is not None:
        generator.set_state(state.cpu())
    elif cuda:
        torch.cuda.set_rng_state(state.cpu(), device)
    else:
        torch.set_rng_state(state.cpu())
    return True


def reset_opacity_values(opacity_raw: torch.Tensor) -> torch.Tensor:
    """`reset_opacity` (:263-266): inverse_sigmoid(min(sigmoid(o), 0.01)), utils/general_utils.py:18-19."""
    o = torch.sigmoid(opacity_raw)
    x = torch.min(o, torch.ones_like(o) * 0.01)
    return torch.log(x / (1 - x))


def apply_plan(params, plan: DensifyPlan):
    """Moves `params` (a trainstep.GaussianParamArena on a CUDA device) to the model `plan` describes: one
    gsr_gather_rows launch over the five slices of the parameter arena and of both Adam moment arenas, then the split
    children's positions and scales.  In place: `params` keeps its identity, its tensors are new allocations."""
    from . import _C
    from .trainstep import GaussianParamArena
    assert plan.n_src == params.P, f"plan was made for a model of {plan.n_src} Gaussians, this one has {params.P}"
    new = GaussianParamArena(plan.n_dst, params.M, params.device)
    src_row = plan.src_row.to(device=params.device, dtype=torch.int32).contiguous()
    if plan.n_dst:
        lo, hi = int(src_row.min()), int(src_row.max())
        if lo < 0 or hi >= plan.n_src:
            raise RuntimeError(f"DensifyPlan.src_row out of range [{lo}, {hi}] for {plan.n_src} source rows")
        segs = []
        for name in ("_xyz", "_features", "_opacity", "_scaling", "_rotation"):
            segs.append(dict(src=getattr(params, name), dst=getattr(new, name), zero_new=False))
            for old_m, new_m in zip(params.moments(name), new.moments(name)):
                segs.append(dict(src=old_m, dst=new_m, zero_new=True))
        _C.gather_rows(src_row, plan.n_src, segs, n_keep_state=plan.n_keep_state)
        if plan.child_rows.numel():
            rows = plan.child_rows.to(params.device)
            new._xyz[rows] = plan.child_xyz.to(params.device)
            new._scaling[rows] = plan.child_scaling.to(params.device)
    step, active = params.step_count, getattr(params, "active_sh_degree", 0)
    params.__dict__.update(new.__dict__)
    params.step_count, params.active_sh_degree = step, active
    return params
