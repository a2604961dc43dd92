"""View-sharded multi-view step: the one place this path shards.

The reference renders one camera per iteration in one process (gs-simp/train.py:78,86;
inpaint_rec.py:100,108) and fans scenes out over GPUs with shell jobs (gs-simp/train.sh:1), no
communication.  Its multi-view consumers -- SVD's 14/25-frame orbits (scene/__init__.py:129-198),
SDS view batches, test-set renders (render.py:32-39) -- iterate views that are independent given
the Gaussians.  Here one process per GPU holds a replica of the Gaussians, view v of a batch goes
to rank v mod R, every rank accumulates its views' parameter gradients into ONE flat fp32 arena
(the kernels add in place: GSR_FLAG_ACCUMULATE), and a single all-reduce (NCCL over NVLink, or
gloo in the CPU tests) sums the arena.  Densification statistics keep the reference's per-view
semantics (scene/gaussian_model.py:482-484, train.py:115): sum over views of ||means2D.grad[:, :2]||,
sum of visibility counts, max of radii -- reduced with SUM / SUM / MAX.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Sequence

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> list[int]:
    """Round-robin: view v -> rank v mod R (25 views on 8 ranks -> 4/3 per rank)."""
    return list(range(rank, n_views, world))


class GradArena:
    """Flat fp32 buffer [dL_dmeans3D | dL_dsh | dL_dopacity | dL_dscales | dL_drotations] =
    3 + 3M + 1 + 3 + 4 floats per Gaussian (14 / 23 / 59 at M = 1 / 4 / 16), plus typed views, plus the
    densification statistics.  Everything lives in ONE allocation
        [ flat | grad_norm_accum f32[P] | visible_count i32[P] | max_radii i32[P] ]
    so that one collective covers it.  With `symmetric=True` (CUDA, world > 1) the allocation is
    symmetric memory with an NVSwitch multicast mapping and all_reduce() runs the in-switch kernel
    gsr_nvls_all_reduce (multimem.ld_reduce / multimem.st) between two cross-rank barriers; without
    multicast support (or on CPU / gloo) it falls back to torch.distributed all-reduces."""

    def __init__(self, P: int, M: int, device, symmetric: bool = False, group=None):
        self.P, self.M = P, M
        device = torch.device(device)
        sizes = [("dL_dmeans3D", (P, 3)), ("dL_dsh", (P, M, 3)), ("dL_dopacity", (P, 1)),
                 ("dL_dscales", (P, 3)), ("dL_drotations", (P, 4))]
        # each slice starts on a 16-byte boundary (float4 stores in the kernels)
        offs, off = [], 0
        for _, shp in sizes:
            offs.append(off)
            off += (_numel(shp) + 3) // 4 * 4
        n_flat, Pp = off, (P + 3) // 4 * 4
        total = n_flat + 3 * Pp
        self._handle, self._mc = None, 0
        self.group = group
        storage = None
        if symmetric and device.type == "cuda" and dist.is_available() and dist.is_initialized() \
                and dist.get_world_size(group) > 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                storage = symm_mem.empty(max(total, 4), dtype=torch.float32, device=device)
                self._handle = symm_mem.rendezvous(storage, group if group is not None else dist.group.WORLD)
                self._mc = int(getattr(self._handle, "multicast_ptr", 0) or 0)
            except Exception as ex:   # no symmetric memory on this system / build: plain allocation + NCCL
                self._handle, self._mc, storage = None, 0, None
                self.symmetric_error = repr(ex)
        if storage is None:
            storage = torch.empty(max(total, 4), dtype=torch.float32, device=device)
        storage.zero_()
        if self._handle is not None:
            # every rank's replica (and the allocation's signal pads) initialised before anybody's first in-switch
            # reduction can touch it: one host-level rendezvous per arena, at creation only
            torch.cuda.synchronize(device)
            dist.barrier(group)
        self.storage = storage
        self._n_f32 = n_flat + Pp                     # float SUM segment: flat + grad_norm_accum
        self._sh_first = offs[1]                      # first float of the (P, M, 3) SH gradient inside it
        self.sparse = True                            # skip SH rows of Gaussians no rank saw (visible_count == 0)
        self._off_cnt, self._off_max = 4 * (n_flat + Pp), 4 * (n_flat + 2 * Pp)   # byte offsets of the int segments
        self._offs = {name: o for (name, _), o in zip(sizes, offs)}          # first float of each slice
        self._row_w = {name: _numel(shp) // max(P, 1) for name, shp in sizes}  # floats per Gaussian
        self._n_flat, self._Pp = n_flat, Pp
        self._comm_stream = None
        self.flat = storage[:n_flat]
        self.views = {name: self.flat[o:o + _numel(shp)].view(*shp) for (name, shp), o in zip(sizes, offs)}
        # densification statistics (different reductions: SUM / SUM / MAX)
        self.grad_norm_accum = storage[n_flat:n_flat + P]
        self.visible_count = storage[n_flat + Pp:n_flat + Pp + P].view(torch.int32)
        self.max_radii = storage[n_flat + 2 * Pp:n_flat + 2 * Pp + P].view(torch.int32)

    @property
    def uses_nvls(self) -> bool:
        """the exchange is one of this repository's kernels over symmetric memory (in-switch, or peer-to-peer at two ranks)
        -- the ones that can be issued per Gaussian range and pipelined with the backward"""
        return self._mc != 0 and (self.method == "nvls" or (self.method == "p2p" and self._peer_ptr() != 0))

    def _peer_ptr(self) -> int:
        """the peer's replica of this arena as mapped here (two ranks only), 0 when unavailable"""
        h = self._handle
        if h is None or h.world_size != 2:
            return 0
        try:
            return int(h.buffer_ptrs[1 - h.rank]) + int(getattr(h, "offset", 0) or 0)
        except Exception:
            return 0

    method = "nvls"   # preferred collective when a multicast mapping exists; see calibrate()
    taper = False     # chunk_ranges(): half-length first and last range of the pipelined exchange
    # CTAs of the in-switch kernel (512 threads each; 0 = the library's default of two per SM).  The reduction is bound by
    # the links, not by the SMs, and when it overlaps the chunked per-Gaussian backward every CTA it holds is taken from
    # that kernel: tools/exp_scale8.py sweeps this at 8 GPUs.
    nvls_blocks = int(__import__("os").environ.get("GSR_NVLS_BLOCKS", "0"))

    def calibrate(self, iters: int = 3) -> dict:
        """Times the in-switch kernel against the NCCL all-reduce on this arena (contents are summed
        repeatedly: call it before the arena holds anything of value) and keeps the faster one.  All ranks
        agree because the decision uses the max over ranks.  At 2 GPUs NCCL's direct peer copies win; from
        4 GPUs on the switch reduction does."""
        if not self._mc:
            self.method = "nccl"
            return {"method": "nccl", "reason": "no multicast mapping"}
        res = {}
        dev = self.storage.device
        self.visible_count.fill_(1)   # worst case for the row-sparse kernel: every Gaussian seen by somebody
        for m in ("nccl", "nvls"):
            self.method = m
            self.all_reduce()
            torch.cuda.synchronize(dev)
            dist.barrier(self.group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                self.all_reduce()
            e1.record()
            torch.cuda.synchronize(dev)
            t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            res[m + "_ms"] = float(t.item())
        self.method = "nvls" if res["nvls_ms"] < res["nccl_ms"] else "nccl"
        res["method"] = self.method
        self.zero_()
        return res

    def zero_(self):
        self.storage.zero_()

    def add_view_stats(self, dL_dmeans2D: torch.Tensor, radii: torch.Tensor):
        """Densification statistics of one view (gaussian_model.py:483-484, train.py:115) added into the arena by
        one fused kernel (gsr_accumulate_view_stats) instead of six torch launches.  CUDA only: there is no CPU
        path (the gloo tests of the sharding logic accumulate their oracle views themselves)."""
        if not radii.is_cuda:
            raise RuntimeError("GradArena.add_view_stats needs CUDA tensors: the statistics kernel has no CPU path")
        from . import _C
        _C.accumulate_view_stats(radii, dL_dmeans2D, self.grad_norm_accum, self.visible_count, self.max_radii)

    def all_reduce(self, group=None):
        group = group if group is not None else self.group
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        if self._mc and self.method == "p2p" and self._peer_ptr():
            self.all_reduce_range(0, self.P)      # the plan form covers the whole arena; barriers inside
            return
        if self._mc and self.method == "nvls":
            from . import _C
            h = self._handle
            h.barrier()      # every replica is complete (stream-ordered after this rank's kernels)
            rows = self.P if (self.sparse and (3 * self.M) % 4 == 0) else 0
            _C.nvls_all_reduce(self._mc, self.storage.device, 0, self._n_f32, self._off_cnt, self.P, self._off_max, self.P,
                               h.rank, h.world_size, self.nvls_blocks, self._sh_first, rows, 3 * self.M)
            h.barrier()      # every slice has been written back everywhere
            return
        dist.all_reduce(self.storage[:self._n_f32], op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(self.visible_count, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=group)

    def all_reduce_range(self, g0: int, g1: int, post_barrier: bool = True):
        """In-switch all-reduce of the arena rows (and statistics) of Gaussians [g0, g1) only, on the current
        stream; g0 must be a multiple of 4.  NVLS only (see cuda_views_geom_backward_allreduce).  The barrier
        AFTER the kernel (every rank's slice written back everywhere) is only needed before somebody reads the
        result or overwrites the rows again: a caller reducing several ranges back to back passes
        post_barrier=False and ends with one `arena.barrier()`."""
        assert self.uses_nvls and g0 % 4 == 0 and 0 <= g0 <= g1 <= self.P
        from . import _C
        h = self._handle
        r4 = lambda n: (n + 3) // 4 * 4     # slices are padded to 4 floats: rounding a tail up stays inside the slice
        dense, rows = [], None
        for name in ("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations"):
            w, first = self._row_w[name], self._offs[name]
            if name == "dL_dsh" and self.sparse and w in (12, 48):
                rows = (4 * (first + g0 * w), g1 - g0, w, self._off_cnt + 4 * g0)
            else:
                dense.append((4 * (first + g0 * w), r4((g1 - g0) * w)))
        dense.append((4 * (self._n_flat + g0), r4(g1 - g0)))                      # grad_norm_accum
        h.barrier()
        if self.method == "p2p":
            _C.p2p_all_reduce_plan(self.storage.data_ptr(), self._peer_ptr(), self.storage.device, h.rank, dense=dense, rows=rows,
                                   add_s32=(self._off_cnt + 4 * g0, g1 - g0), max_s32=(self._off_max + 4 * g0, g1 - g0),
                                   blocks=self.nvls_blocks)
        else:
            _C.nvls_all_reduce_plan(self._mc, self.storage.device, h.rank, h.world_size, dense=dense, rows=rows,
                                    add_s32=(self._off_cnt + 4 * g0, g1 - g0), max_s32=(self._off_max + 4 * g0, g1 - g0),
                                    blocks=self.nvls_blocks)
        if post_barrier:
            h.barrier()

    def resized(self, P: int) -> "GradArena":
        """A zeroed arena for a model of P Gaussians on the same device / group / memory kind -- what the caller
        swaps in after `GaussianParamArena.densify_and_prune` (densification_postfix zeroes the statistics,
        gaussian_model.py:423-425).  With symmetric memory this is a collective call (rendezvous)."""
        new = GradArena(P, self.M, self.storage.device, symmetric=self._handle is not None, group=self.group)
        new.sparse = self.sparse
        new.nvls_blocks = self.nvls_blocks
        new.method = self.method if new._mc else "nccl"
        return new

    def pruned(self, mask: torch.Tensor) -> "GradArena":
        """The arena after `prune_points(mask)`: statistics of the surviving Gaussians carried over
        (gaussian_model.py:379-383), gradients zero."""
        keep = ~mask.reshape(-1).bool().to(self.storage.device)
        new = self.resized(int(keep.sum()))
        new.grad_norm_accum.copy_(self.grad_norm_accum[keep])
        new.visible_count.copy_(self.visible_count[keep])
        new.max_radii.copy_(self.max_radii[keep])
        return new

    def barrier(self):
        """Cross-rank barrier on the current stream (symmetric-memory signal pads)."""
        self._handle.barrier()

    def comm_stream(self):
        if self._comm_stream is None:
            # high priority: the reduction's CTAs must get SM slots while the (grid-filling) backward of the next
            # chunk is running, otherwise the two kernels simply run one after the other
            self._comm_stream = torch.cuda.Stream(self.storage.device, priority=-1)
        return self._comm_stream

    @property
    def bytes(self) -> int:
        return self.flat.numel() * 4


def _numel(shp):
    n = 1
    for s in shp:
        n *= s
    return n


@dataclass
class ViewResult:
    color: torch.Tensor
    depth: torch.Tensor
    radii: torch.Tensor
    num_rendered: int


class ViewPipeline:
    """Consecutive views of a step alternate between `depth` CUDA streams, so that the latency-bound
    front end of view v+1 (K1, the sorts, binning: small grids that leave most of the GPU idle) runs
    under the issue-bound blend kernels of view v.  Views stay independent up to the gradient arena:
    the per-Gaussian backward of view v+1 is ordered after that of view v with an event (its plain
    read-modify-writes of the arena must not interleave), everything before it overlaps freely.

        with pipe.step():                       # side streams wait for the caller's stream ...
            for v in views:
                cuda_view_fwd_bwd(..., pipeline=pipe)
        # ... and the caller's stream waits for them here (arena, outputs and stats are ready in stream order)

    `dL_dcolor_fn` runs on the view's stream; anything else the caller does with a view's outputs
    belongs after the `with` block."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(max(depth, 1))]
        self._k = 0
        self._prev_bwd = None
        self.slot = 0

    def step(self):
        return _PipelineStep(self)

    def _begin(self):
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        for st in self.streams:
            st.wait_event(ev)
        self._prev_bwd = None

    def _end(self):
        main = torch.cuda.current_stream(self.device)
        for st in self.streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)

    def next_stream(self):
        self.slot = self._k % len(self.streams)
        self._k += 1
        return torch.cuda.stream(self.streams[self.slot])

    def order_backward(self):
        if self._prev_bwd is not None:
            torch.cuda.current_stream(self.device).wait_event(self._prev_bwd)

    def mark_backward_done(self):
        self._prev_bwd = torch.cuda.Event()
        self._prev_bwd.record(torch.cuda.current_stream(self.device))


class _PipelineStep:
    def __init__(self, pipe):
        self.pipe = pipe

    def __enter__(self):
        self.pipe._begin()
        return self.pipe

    def __exit__(self, *exc):
        self.pipe._end()
        return False


def cuda_view_fwd_bwd(gaussians: dict, settings, dL_dcolor_fn: Callable[[torch.Tensor], torch.Tensor],
                      arena: GradArena, flags: int | None = None, capacity: int | None = None,
                      async_result: torch.Tensor | None = None, pipeline: ViewPipeline | None = None) -> ViewResult:
    """One view through the CUDA path: forward, loss gradient, backward ADDING into `arena`.
    `gaussians`: means3D, shs, opacities, scales, rotations (post-activation, as render() passes them);
    `settings`: GaussianRasterizationSettings (a callable returning them is evaluated on the view's
    stream, e.g. to stage a camera from pinned memory).  With `async_result` (pinned int64[2]) +
    `capacity` nothing in this call blocks the host; check the result with `AsyncViews.check()` after
    the step.  With `pipeline` the view runs on the pipeline's next stream (see ViewPipeline)."""
    import contextlib
    from . import _C
    flags = _C.resolve_flags(flags)
    with (pipeline.next_stream() if pipeline is not None else contextlib.nullcontext()):
        rs = settings() if callable(settings) else settings
        e = torch.empty(0, device=gaussians["means3D"].device)
        n, color, radii, geom, binning, img, depth = _C.rasterize_gaussians(
            rs.bg, gaussians["means3D"], e, gaussians["opacities"], gaussians["scales"], gaussians["rotations"],
            rs.scale_modifier, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
            rs.image_width, gaussians["shs"], rs.sh_degree, rs.campos, rs.prefiltered, flags=flags,
            capacity=capacity, async_result=async_result)
        dL = dL_dcolor_fn(color)
        if pipeline is not None:
            pipeline.order_backward()
        out = _C.rasterize_gaussians_backward(
            rs.bg, gaussians["means3D"], radii, e, gaussians["scales"], gaussians["rotations"], rs.scale_modifier, e,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, dL, gaussians["shs"], rs.sh_degree, rs.campos,
            geom, n, binning, img, flags=flags | _C.FLAG_ACCUMULATE, out=arena.views)
        arena.add_view_stats(out[0], radii)
        if pipeline is not None:
            pipeline.mark_backward_done()
    return ViewResult(color, depth, radii, n)


@dataclass
class ViewState:
    """What a view leaves behind between its forward + blend backward (K1..K7) and the batched
    per-Gaussian backward (cuda_views_geom_backward)."""
    color: torch.Tensor
    depth: torch.Tensor
    radii: torch.Tensor
    num_rendered: int
    geom: torch.Tensor
    scratch: torch.Tensor
    settings: object


def cuda_view_fwd_blend(gaussians: dict, settings, dL_dcolor_fn: Callable[[torch.Tensor], torch.Tensor], flags: int | None = None,
                        capacity: int | None = None, async_result: torch.Tensor | None = None,
                        pipeline: ViewPipeline | None = None, workspace=None) -> ViewState:
    """Forward (K1..K6), loss gradient, blend backward (K7) of one view; the per-Gaussian chain rule is
    left to cuda_views_geom_backward(), which runs it ONCE for all views of the step.  With `pipeline`
    the view runs on the pipeline's next stream; views need no mutual ordering here."""
    import contextlib
    from . import _C
    flags = _C.resolve_flags(flags)
    with (pipeline.next_stream() if pipeline is not None else contextlib.nullcontext()):
        rs = settings() if callable(settings) else settings
        e = torch.empty(0, device=gaussians["means3D"].device)
        n, color, radii, geom, binning, img, depth = _C.rasterize_gaussians(
            rs.bg, gaussians["means3D"], e, gaussians["opacities"], gaussians["scales"], gaussians["rotations"],
            rs.scale_modifier, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
            rs.image_width, gaussians["shs"], rs.sh_degree, rs.campos, rs.prefiltered, flags=flags,
            capacity=capacity, async_result=async_result, workspace=workspace)
        dL = dL_dcolor_fn(color)
        scratch = _C.backward_blend(rs.bg, dL, geom, binning, img, gaussians["means3D"].shape[0], flags=flags,
                                    workspace=workspace)
    return ViewState(color, depth, radii, n, geom, scratch, rs)


def cuda_views_fwd_blend_batched(gaussians: dict, settings_list: Sequence, dL_dcolor_fns: Sequence[Callable], flags: int,
                                 capacities: Sequence[int], async_results: Sequence, workspaces: Sequence | None = None
                                 ) -> list[ViewState]:
    """Forward (K1..K6) of ALL views with one launch per stage (gsr_forward_views), the loss gradients, then the blend
    backward (K7) of all views in one launch (gsr_backward_blend_views).  The views share the Gaussians, so K1's
    parameter reads come from HBM once; the sorts and the expansion run as segmented kernels that fill the machine
    where a single view's grids (1.7 waves per sort pass at 3 M Gaussians) left it latency-bound.  Asynchronous only:
    needs capacities + pinned async_results (AsyncViews)."""
    from . import _C
    rss = [s() if callable(s) else s for s in settings_list]
    if not rss:
        return []
    rs0 = rss[0]
    e = torch.empty(0, device=gaussians["means3D"].device)
    cleared = []   # the blend backward's accumulators, cleared by K1 on the way (no memset of their own)
    outs = _C.forward_views(rs0.bg, gaussians["means3D"], e, gaussians["opacities"], gaussians["scales"], gaussians["rotations"],
                            rs0.scale_modifier, e, rss, gaussians["shs"], rs0.sh_degree, rs0.prefiltered, capacities,
                            async_results, workspaces=workspaces, flags=flags, scratches_out=cleared)
    dLs = [fn(o[1]) for fn, o in zip(dL_dcolor_fns, outs)]
    scratches = _C.backward_blend_views(rs0.bg, dLs, [o[3] for o in outs], [o[4] for o in outs], [o[5] for o in outs],
                                        gaussians["means3D"].shape[0], flags=flags, workspaces=workspaces, scratches=cleared)
    return [ViewState(o[1], o[6], o[2], -1, o[3], sc, rs) for o, sc, rs in zip(outs, scratches, rss)]


def cuda_views_geom_backward(gaussians: dict, states: Sequence[ViewState], arena: GradArena, accumulate: bool = False,
                             flags: int | None = None, want_means2D: bool = False):
    """Batched K8+K9 over `states` (one launch per four views): parameter gradients and densification
    statistics are WRITTEN into `arena` (added with `accumulate`), so the arena needs no zeroing."""
    from . import _C
    flags = _C.resolve_flags(flags)
    if not states:
        # a rank without views (fewer views than ranks): the batched kernel WRITES the arena, so nothing else would
        # clear last step's gradients and statistics -- this rank must contribute zeros to the all-reduce
        if not accumulate:
            arena.zero_()
        return None
    views = [dict(radii=s.radii, geom=s.geom, scratch=s.scratch, viewmatrix=s.settings.viewmatrix,
                  projmatrix=s.settings.projmatrix, campos=s.settings.campos, tanfovx=s.settings.tanfovx,
                  tanfovy=s.settings.tanfovy, width=s.settings.image_width, height=s.settings.image_height) for s in states]
    rs = states[0].settings
    cur = torch.cuda.current_stream(gaussians["means3D"].device)
    for st in states:   # produced on the pipeline's streams, consumed here: keep the allocator from recycling them early
        for t in (st.radii, st.geom, st.scratch):
            t.record_stream(cur)
    return _C.backward_geom_multi(gaussians["means3D"], gaussians["shs"], gaussians["scales"], gaussians["rotations"],
                                  rs.scale_modifier, rs.sh_degree, views, arena.views,
                                  stats=(arena.grad_norm_accum, arena.visible_count, arena.max_radii),
                                  flags=flags | (_C.FLAG_ACCUMULATE if accumulate else 0), want_means2D=want_means2D)


def chunk_ranges(P: int, chunks: int, taper: bool = False) -> list[tuple[int, int]]:
    """Gaussian ranges [g0, g1) of the pipelined backward + exchange, boundaries on multiples of 32.  `taper`: the first
    and the last range are half as long as the others (weights 1/2, 1, ..., 1, 1/2).  The exchange is the longer stage
    of the pipeline (the links, not the SMs), so what stays exposed is the wait for the FIRST chunk's backward before
    any reduction can start plus whatever the reductions lag behind at the end: a short first chunk starts the links
    earlier, a short last one shortens the tail nobody overlaps."""
    if chunks <= 1 or P <= 0:
        return [(0, P)] if P > 0 else []
    if not taper or chunks < 3:
        step = ((P + chunks - 1) // chunks + 31) // 32 * 32
        return [(g0, min(P, g0 + step)) for g0 in range(0, P, step)]
    unit = P / (chunks - 1)                       # weights sum to chunks - 1
    edges, acc = [0], 0.0
    for k in range(chunks):
        acc += unit * (0.5 if k in (0, chunks - 1) else 1.0)
        e = P if k == chunks - 1 else min(P, (int(acc) + 31) // 32 * 32)
        if e > edges[-1]:
            edges.append(e)
    if edges[-1] != P:
        edges.append(P)
    return list(zip(edges[:-1], edges[1:]))


def cuda_views_geom_backward_allreduce(gaussians: dict, states: Sequence[ViewState], arena: GradArena, flags: int | None = None,
                                       chunks: int = 4):
    """Batched K8+K9 + gradient all-reduce of a multi-rank step, pipelined over Gaussian-range chunks: while the
    switch reduces the arena rows of chunk c (comm stream: barrier, gsr_nvls_all_reduce_plan, barrier), the
    per-Gaussian backward of chunk c + 1 runs on the caller's stream.  The backward is HBM-bound, the reduction
    NVLink-bound, so the backward disappears behind the collective.  Without a multicast mapping (or with NCCL
    selected) it is the plain sequence: backward, then arena.all_reduce()."""
    from . import _C
    flags = _C.resolve_flags(flags)
    P = arena.P
    if not (arena.uses_nvls and dist.is_initialized() and dist.get_world_size(arena.group) > 1) or chunks <= 1 \
            or P < 4096:
        cuda_views_geom_backward(gaussians, states, arena, flags=flags)
        arena.all_reduce()
        return
    views = [dict(radii=s.radii, geom=s.geom, scratch=s.scratch, viewmatrix=s.settings.viewmatrix,
                  projmatrix=s.settings.projmatrix, campos=s.settings.campos, tanfovx=s.settings.tanfovx,
                  tanfovy=s.settings.tanfovy, width=s.settings.image_width, height=s.settings.image_height) for s in states]
    rs = states[0].settings if states else None
    dev = gaussians["means3D"].device
    main, comm = torch.cuda.current_stream(dev), arena.comm_stream()
    if not states:
        arena.zero_()   # a rank without views contributes zeros, and still takes part in every barrier and reduction below
    for st in states:
        for t in (st.radii, st.geom, st.scratch):
            t.record_stream(main)
    for g0, g1 in chunk_ranges(P, chunks, getattr(arena, "taper", False)):
        if states:
            _C.backward_geom_multi(gaussians["means3D"], gaussians["shs"], gaussians["scales"], gaussians["rotations"],
                                   rs.scale_modifier, rs.sh_degree, views, arena.views,
                                   stats=(arena.grad_norm_accum, arena.visible_count, arena.max_radii), flags=flags, g_range=(g0, g1))
        ev = torch.cuda.Event()
        ev.record(main)
        comm.wait_event(ev)
        with torch.cuda.stream(comm):
            arena.all_reduce_range(g0, g1, post_barrier=False)
    with torch.cuda.stream(comm):
        arena.barrier()
    main.wait_stream(comm)


def cuda_views_fwd_bwd(gaussians: dict, settings_list: Sequence, dL_dcolor_fns: Sequence[Callable], arena: GradArena,
                       flags: int | None = None, capacities: Sequence[int] | None = None, async_results: Sequence | None = None,
                       pipeline: ViewPipeline | None = None, accumulate: bool = False, all_reduce: bool = False,
                       chunks: int = 4, workspaces: Sequence | None = None, batched: bool = False) -> list[ViewState]:
    """A rank's share of a multi-view step: every view's forward + blend backward (two views in flight
    with `pipeline`), then ONE batched per-Gaussian backward that writes the arena.  Falls back to the
    per-view accumulate path when the batched kernel does not cover the configuration (M not in 1/4/16).
    With `all_reduce` the arena is also summed over ranks, pipelined with the per-Gaussian backward
    (cuda_views_geom_backward_allreduce).  `workspaces`: one _C.Workspace per view (kept by the caller across
    steps) -- the steady-state loop then never touches the caching allocator."""
    import contextlib
    from . import _C
    flags = _C.resolve_flags(flags)
    M = gaussians["shs"].shape[1]
    if not _C.backward_geom_multi_supported(M):
        if not accumulate:
            arena.zero_()
        out = []
        with (pipeline.step() if pipeline else contextlib.nullcontext()):
            for k, rs in enumerate(settings_list):
                r = cuda_view_fwd_bwd(gaussians, rs, dL_dcolor_fns[k], arena, flags=flags,
                                      capacity=capacities[k] if capacities else None,
                                      async_result=async_results[k] if async_results else None, pipeline=pipeline)
                out.append(ViewState(r.color, r.depth, r.radii, r.num_rendered, None, None, rs))
        if all_reduce:
            arena.all_reduce()
        return out
    states = []
    if batched and capacities and async_results and all(c > 0 for c in capacities) and not (flags & (_C.FLAG_BINNING_KEY64 | _C.FLAG_REFERENCE)):
        # one launch per stage for all views of a group (`batched`): every launch fills the machine.  With a pipeline the
        # views are split into one group per stream, so that the memory- / latency-bound front end of one group runs
        # under the issue-bound blend kernels of the other.
        n = len(settings_list)
        n_groups = min(len(pipeline.streams), n) if pipeline is not None else 1
        bounds = [n * g // n_groups for g in range(n_groups + 1)]
        states = []
        with (pipeline.step() if pipeline else contextlib.nullcontext()):
            for g in range(n_groups):
                sl = slice(bounds[g], bounds[g + 1])
                with (pipeline.next_stream() if pipeline is not None else contextlib.nullcontext()):
                    states += cuda_views_fwd_blend_batched(gaussians, settings_list[sl], dL_dcolor_fns[sl], flags, capacities[sl],
                                                           async_results[sl], workspaces[sl] if workspaces else None)
        settings_list = []
    with (pipeline.step() if pipeline else contextlib.nullcontext()):
        for k, rs in enumerate(settings_list):
            states.append(cuda_view_fwd_blend(gaussians, rs, dL_dcolor_fns[k], flags=flags,
                                              capacity=capacities[k] if capacities else None,
                                              async_result=async_results[k] if async_results else None, pipeline=pipeline,
                                              workspace=workspaces[k] if workspaces else None))
    if all_reduce and not accumulate:
        cuda_views_geom_backward_allreduce(gaussians, states, arena, flags=flags, chunks=chunks)
    else:
        cuda_views_geom_backward(gaussians, states, arena, accumulate=accumulate, flags=flags)
        if all_reduce:
            arena.all_reduce()
    return states


class StepThrottle:
    """Bounds how far the host may run ahead of the GPU when steps never block (GSR_FLAG_ASYNC): at most
    `max_inflight` steps queued.  Every queued step pins ~1 GB of per-view scratch (geometry, binning and
    image buffers) until it has run, so an unbounded run-ahead makes the caching allocator fall back to
    cudaMalloc -- a device synchronisation -- in the middle of the loop.  Waiting on the event of the step
    before the previous one keeps the GPU fed (a whole step is still queued behind the running one)."""

    def __init__(self, max_inflight: int = 2):
        self.max_inflight = max(1, max_inflight)
        self._events: list = []

    def tick(self, device=None):
        """Call once per step, after queueing it."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self._events.append(ev)
        if len(self._events) > self.max_inflight:
            self._events.pop(0).synchronize()


class AsyncViews:
    """Host-side bookkeeping for fully asynchronous view steps: one pinned (N, status) slot and one
    capacity per view.  Usage per step:  for v: cuda_view_fwd_bwd(..., capacity=a.capacity(v),
    async_result=a.slot(v));  <sync / all-reduce>;  redo = a.check(views)  -> views whose capacity
    overflowed (their contribution is invalid: zero the arena and redo the step)."""

    def __init__(self, n_views: int, margin: float = 1.25):
        self.slots = torch.zeros(n_views, 2, dtype=torch.int64)
        if torch.cuda.is_available():
            self.slots = self.slots.pin_memory()
        self.cap = [0] * n_views
        self.margin = margin

    def slot(self, v: int) -> torch.Tensor:
        return self.slots[v]

    def capacity(self, v: int) -> int:
        """One capacity for all views (the largest seen, rounded up to 2^20 instances): every view's binning
        buffer then has the same size, so the caching allocator recycles blocks between views instead of
        growing a pool per size."""
        c = max(self.cap)
        return (c + (1 << 20) - 1) >> 20 << 20 if c > 0 else 0

    def learn(self, v: int, n: int):
        self.cap[v] = max(self.cap[v], int(n * self.margin) + 65536)

    def check(self, views) -> list[int]:
        """Call after the stream is synchronised.  Returns the views that must be redone."""
        bad = []
        for v in views:
            n, status = int(self.slots[v, 0]), int(self.slots[v, 1])
            if status & 0xffffffff:
                raise RuntimeError("rasterize_gaussians: point filtered by culling but 'prefiltered' was set")
            if n > self.capacity(v) or (status >> 32):
                bad.append(v)
            self.learn(v, n)
        return bad


def sharded_step(view_fwd_bwd: Callable[[int], None], n_views: int, arena: GradArena, group=None,
                 rank: int | None = None, world: int | None = None) -> list[int]:
    """Runs `view_fwd_bwd(v)` for this rank's share of the views (each call accumulates into
    `arena`), then all-reduces.  Returns the view indices this rank processed."""
    if rank is None:
        rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    mine = shard_views(n_views, rank, world)
    for v in mine:
        view_fwd_bwd(v)
    arena.all_reduce(group)
    return mine


def cuda_views_render(gaussians: dict, settings_list: Sequence, flags: int | None = None, capacities: Sequence[int] | None = None,
                      async_results: Sequence | None = None, pipeline: ViewPipeline | None = None,
                      workspaces: Sequence | None = None,
                      sink: Callable[[int, torch.Tensor, torch.Tensor, torch.Tensor], None] | None = None,
                      batched: bool = False) -> list[ViewResult]:
    """Batched inference: the forward of every view of `settings_list` -- the loops of
    render.py:32-39, render_depth.py:31-39 and gen_seq.py:36-58, which call render() once per camera under
    no_grad and then block on an image write.  Views alternate over the pipeline's streams (front end of view
    v+1 under the blend of view v), each with its own caller-owned Workspace, and with `capacities` +
    `async_results` nothing blocks the host.  `sink(k, color, depth, radii)` runs on view k's stream right after
    its forward -- the place for the device-to-host copy into pinned memory that an asynchronous PNG / NPY
    writer drains; with workspaces the three tensors are views that the next call on that workspace overwrites."""
    import contextlib
    from . import _C
    flags = _C.resolve_flags(flags)
    out = []
    if batched and capacities and async_results and all(c > 0 for c in capacities) and settings_list \
            and not (flags & (_C.FLAG_BINNING_KEY64 | _C.FLAG_REFERENCE)):
        # one launch per stage for up to eight views at a time (gsr_forward_views); groups alternate over the pipeline's streams
        with torch.no_grad(), (pipeline.step() if pipeline else contextlib.nullcontext()):
            for k0 in range(0, len(settings_list), _C.MAX_BATCH):
                sl = slice(k0, min(k0 + _C.MAX_BATCH, len(settings_list)))
                with (pipeline.next_stream() if pipeline is not None else contextlib.nullcontext()):
                    rss = [s() if callable(s) else s for s in settings_list[sl]]
                    e = torch.empty(0, device=gaussians["means3D"].device)
                    outs = _C.forward_views(rss[0].bg, gaussians["means3D"], e, gaussians["opacities"], gaussians["scales"],
                                            gaussians["rotations"], rss[0].scale_modifier, e, rss, gaussians["shs"], rss[0].sh_degree,
                                            rss[0].prefiltered, capacities[sl], async_results[sl],
                                            workspaces=workspaces[sl] if workspaces else None, flags=flags)
                    for j, o in enumerate(outs):
                        if sink is not None:
                            sink(k0 + j, o[1], o[6], o[2])
                        out.append(ViewResult(o[1], o[6], o[2], -1))
        return out
    with torch.no_grad(), (pipeline.step() if pipeline else contextlib.nullcontext()):
        for k, settings in enumerate(settings_list):
            with (pipeline.next_stream() if pipeline is not None else contextlib.nullcontext()):
                rs = settings() if callable(settings) else settings
                e = torch.empty(0, device=gaussians["means3D"].device)
                n, color, radii, _geom, _binning, _img, depth = _C.rasterize_gaussians(
                    rs.bg, gaussians["means3D"], e, gaussians["opacities"], gaussians["scales"], gaussians["rotations"],
                    rs.scale_modifier, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
                    rs.image_width, gaussians["shs"], rs.sh_degree, rs.campos, rs.prefiltered, flags=flags,
                    capacity=capacities[k] if capacities else None,
                    async_result=async_results[k] if async_results else None,
                    workspace=workspaces[k] if workspaces else None)
                if sink is not None:
                    sink(k, color, depth, radii)
                out.append(ViewResult(color, depth, radii, n))
    return out


def render_views_sharded(render_one: Callable[[int], Sequence[torch.Tensor]], n_views: int,
                         rank: int = 0, world: int = 1) -> dict[int, Sequence[torch.Tensor]]:
    """Inference (render.py / render_depth.py loops): no collective, rank r renders views r, r+R, ..."""
    with torch.no_grad():
        return {v: render_one(v) for v in shard_views(n_views, rank, world)}
