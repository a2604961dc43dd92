"""Torch-facing mirror of the reference extension's pybind module `diff_gaussian_rasterization._C`.

Same three entry points, same positional arguments and return tuples as the w-depth fork the
reference installs (README.md:26 of the reference; signatures in SURVEY.md section 8b):

    rasterize_gaussians(...)           -> (num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer, depth)
    rasterize_gaussians_backward(...)  -> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D,
                                           dL_dsh, dL_dscales, dL_drotations)
    mark_visible(means3D, viewmatrix, projmatrix) -> bool[P]

but implemented over the C ABI of include/gsrast_b200.h (libgsrast_b200.so, hand-written sm_100a
CUDA) through ctypes.  torch is only used for device memory and the current stream.  There is NO
CPU or eager fallback: if the library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsrast_b200.so")
if os.environ.get("GSR_LIB_VARIANT"):   # experiments only (tools/build_variant.py): the same sources compiled with extra -D switches
    LIB_PATH = os.path.join(_HERE, "variants", f"libgsrast_b200_{os.environ['GSR_LIB_VARIANT']}.so")

FLAG_BINNING_KEY64 = 1
FLAG_PRECISE = 2
FLAG_REFERENCE = 4     # reference-structure ablation baseline (cub sort, thread-per-pixel blend, 9 atomics/pair)
FLAG_ACCUMULATE = 8
FLAG_ASYNC = 16
FLAG_SCRATCH_CLEARED = 64  # backward_blend(_views): the forward cleared the accumulators (forward_views(scratches_out=...))
FLAG_TIGHT_BINNING = 32  # bin only the tiles the {alpha >= 1/255} bounding box reaches (same images / gradients, shorter lists)
MAX_BATCH = 8           # views per batched launch of the front end / blend kernels (GSR_MAX_BATCH)
GM_MAX_VIEWS = 4        # views per launch of the batched per-Gaussian backward (csrc/geom_backward_multi.cu)
NUM_STAGES = 10
STAGE_NAMES = ("preprocess", "depth_sort", "scan", "duplicate", "tile_sort", "tile_ranges", "blend_forward",
               "accum_clear", "blend_backward", "geom_backward")
#: default kernel flags of the Python layers (see include/gsrast_b200.h); override with GSR_FLAGS=<int>.  flags=0 gives
#: the reference's literal (tile, Gaussian) lists -- what the list-parity tests pin.
DEFAULT_FLAGS = int(os.environ.get("GSR_FLAGS", str(FLAG_TIGHT_BINNING)))


def resolve_flags(flags) -> int:
    return DEFAULT_FLAGS if flags is None else int(flags)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the B200 rasterizer has no fallback path. Build it with "
        "`python -m multiview_inpaint_b200.build` (needs nvcc, sm_100a).")

_lib = C.CDLL(LIB_PATH)

_ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class GsrLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "rec", "depths", "clamped", "tiles_touched", "point_offsets", "order", "geom_bytes",
        "final_T", "n_contrib", "ranges", "image_bytes", "point_list", "binning_bytes")]


_vp, _i, _f, _i64, _u32, _sz = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_uint32, C.c_size_t
_lib.gsr_last_error.restype = C.c_char_p
_lib.gsr_version.restype = _i
_lib.gsr_forward.restype = _i
_lib.gsr_forward.argtypes = [_vp, _ALLOC_FN, _vp, _ALLOC_FN, _vp, _ALLOC_FN, _vp, _i, _i, _i, _vp, _i, _i,
                             _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f, _i,
                             _vp, _vp, _vp, _vp, _i64, _u32]
_lib.gsr_backward_scratch_bytes.restype = _sz
_lib.gsr_backward_scratch_bytes.argtypes = [_i]
_lib.gsr_backward.restype = _i
_lib.gsr_backward.argtypes = [_vp, _i, _i, _i, _i64, _vp, _i, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp,
                              _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                              _vp, _vp, _sz, _u32]
_lib.gsr_mark_visible.restype = _i
_lib.gsr_mark_visible.argtypes = [_vp, _i, _vp, _vp, _vp, _vp]
class GsrViewGrad(C.Structure):
    _fields_ = [("radii", _vp), ("geom_buffer", _vp), ("scratch", _vp), ("viewmatrix", _vp), ("projmatrix", _vp),
                ("cam_pos", _vp), ("dL_dmean2D", _vp), ("tan_fovx", _f), ("tan_fovy", _f), ("width", _i), ("height", _i)]


class GsrViewForward(C.Structure):
    _fields_ = [("background", _vp), ("viewmatrix", _vp), ("projmatrix", _vp), ("cam_pos", _vp), ("tan_fovx", _f), ("tan_fovy", _f),
                ("width", _i), ("height", _i), ("out_color", _vp), ("out_depth", _vp), ("radii", _vp), ("geom_buffer", _vp),
                ("binning_buffer", _vp), ("image_buffer", _vp), ("capacity", _i64), ("result_host", _vp), ("backward_scratch", _vp)]


class GsrViewBackward(C.Structure):
    _fields_ = [("background", _vp), ("width", _i), ("height", _i), ("geom_buffer", _vp), ("binning_buffer", _vp),
                ("image_buffer", _vp), ("dL_dpix", _vp), ("scratch", _vp), ("scratch_bytes", _sz)]


_lib.gsr_forward_views.restype = _i
_lib.gsr_forward_views.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, C.POINTER(GsrViewForward), _i, _u32]
_lib.gsr_backward_blend_views.restype = _i
_lib.gsr_backward_blend_views.argtypes = [_vp, _i, C.POINTER(GsrViewBackward), _i, _u32]
_lib.gsr_backward_blend.restype = _i
_lib.gsr_backward_blend.argtypes = [_vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _u32]
_lib.gsr_backward_geom_multi.restype = _i
_lib.gsr_backward_geom_multi.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp, _f, _vp, C.POINTER(GsrViewGrad), _i,
                                         _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32]
_lib.gsr_backward_geom_multi_range.restype = _i
_lib.gsr_backward_geom_multi_range.argtypes = _lib.gsr_backward_geom_multi.argtypes + [_i, _i]


class GsrNvlsPlan(C.Structure):
    _fields_ = [("dense_off", _sz * 6), ("dense_n_f32", _sz * 6), ("n_dense", _i), ("rows_off", _sz), ("rows", _sz),
                ("row_f32", _i), ("rows_count_off", _sz), ("add_s32_off", _sz), ("n_add_s32", _sz),
                ("max_s32_off", _sz), ("n_max_s32", _sz)]


_lib.gsr_nvls_all_reduce_plan.restype = _i
_lib.gsr_nvls_all_reduce_plan.argtypes = [_vp, _vp, C.POINTER(GsrNvlsPlan), _i, _i, _i]
_lib.gsr_p2p_all_reduce_plan.restype = _i
_lib.gsr_p2p_all_reduce_plan.argtypes = [_vp, _vp, _vp, C.POINTER(GsrNvlsPlan), _i, _i]
_lib.gsr_nvls_all_reduce.restype = _i
_lib.gsr_nvls_all_reduce.argtypes = [_vp, _vp, _sz, _sz, _sz, _sz, _sz, _sz, _i, _i, _i, _sz, _sz, _i]
_lib.gsr_accumulate_view_stats.restype = _i
_lib.gsr_accumulate_view_stats.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp]
_lib.gsr_knn_temp_bytes.restype = _sz
_lib.gsr_knn_temp_bytes.argtypes = [_i]
_lib.gsr_knn3_mean_dist2.restype = _i
_lib.gsr_knn3_mean_dist2.argtypes = [_vp, _i, _vp, _vp, _vp, _sz]
_lib.gsr_sort_temp_bytes.restype = _sz
_lib.gsr_sort_temp_bytes.argtypes = [_i64, _i, _i]
_lib.gsr_sort_pairs_u64.restype = _i
_lib.gsr_sort_pairs_u64.argtypes = [_vp, _i64, _vp, _vp, _vp, _vp, _i, _vp, _sz]
_lib.gsr_sort_pairs_u32.restype = _i
_lib.gsr_sort_pairs_u32.argtypes = [_vp, _i64, _vp, _vp, _vp, _vp, _i, _vp, _sz]
_lib.gsr_scan_temp_bytes.restype = _sz
_lib.gsr_scan_temp_bytes.argtypes = [_i64]
_lib.gsr_inclusive_scan_u32.restype = _i
_lib.gsr_inclusive_scan_u32.argtypes = [_vp, _i64, _vp, _vp, _vp, _vp, _sz]
_lib.gsr_get_layout.restype = _i
_lib.gsr_get_layout.argtypes = [_i, _i, _i, _i64, _u32, C.POINTER(GsrLayout)]
_lib.gsr_debug_set.restype = _i
_lib.gsr_debug_set.argtypes = [_i, _i]
_lib.gsr_kernel_launches.restype = C.c_uint64
_lib.gsr_profile_enable.restype = None
_lib.gsr_profile_enable.argtypes = [_i]
_lib.gsr_profile_collect.restype = _i
_lib.gsr_profile_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(_i64)]


class GsrAdamSegment(C.Structure):
    _fields_ = [("param", _vp), ("grad", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp), ("n", C.c_uint64),
                ("lr", C.c_double), ("lr_rest", C.c_double), ("row_len", _i), ("row_split", _i)]


_lib.gsr_loss_temp_bytes.restype = _sz
_lib.gsr_loss_temp_bytes.argtypes = [_i, _i, _i]
_lib.gsr_loss_l1_ssim_forward.restype = _i
_lib.gsr_loss_l1_ssim_forward.argtypes = [_vp, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _sz]
_lib.gsr_loss_l1_ssim_backward.restype = _i
_lib.gsr_loss_l1_ssim_backward.argtypes = [_vp, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _sz, _vp]
_lib.gsr_activate_forward.restype = _i
_lib.gsr_activate_forward.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]
_lib.gsr_activate_backward.restype = _i
_lib.gsr_activate_backward.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]
_lib.gsr_quantize_rgb8.restype = _i
_lib.gsr_quantize_rgb8.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp]
_lib.gsr_adam_step.restype = _i
_lib.gsr_adam_step.argtypes = [_vp, C.POINTER(GsrAdamSegment), _i, _i64, C.c_double, C.c_double, C.c_double]

class GsrGatherSegment(C.Structure):
    _fields_ = [("src", _vp), ("dst", _vp), ("row_f32", _i), ("zero_new", _i)]


GATHER_MAX_SEGS = 16
_lib.gsr_gather_rows.restype = _i
_lib.gsr_gather_rows.argtypes = [_vp, _i64, _i64, _i64, _vp, C.POINTER(GsrGatherSegment), _i]

EXPORTED_SYMBOLS = ("gsr_forward", "gsr_backward", "gsr_backward_scratch_bytes", "gsr_mark_visible",
                    "gsr_accumulate_view_stats", "gsr_loss_temp_bytes", "gsr_loss_l1_ssim_forward",
                    "gsr_loss_l1_ssim_backward", "gsr_activate_forward", "gsr_activate_backward", "gsr_adam_step", "gsr_quantize_rgb8", "gsr_gather_rows",
                    "gsr_knn_temp_bytes", "gsr_knn3_mean_dist2", "gsr_forward_views", "gsr_backward_blend_views", "gsr_backward_blend", "gsr_backward_geom_multi", "gsr_backward_geom_multi_range", "gsr_nvls_all_reduce", "gsr_nvls_all_reduce_plan", "gsr_p2p_all_reduce_plan",
                    "gsr_sort_temp_bytes", "gsr_sort_pairs_u64", "gsr_sort_pairs_u32",
                    "gsr_scan_temp_bytes", "gsr_inclusive_scan_u32", "gsr_get_layout",
                    "gsr_profile_enable", "gsr_profile_collect", "gsr_kernel_launches", "gsr_debug_approx_units",
                    "gsr_last_error", "gsr_version", "gsr_debug_set")


def _check(rc: int, what: str):
    if rc != 0:
        msg = _lib.gsr_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def _ptr(t: torch.Tensor | None):
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    """float32, CUDA, contiguous (camera_center is a row-slice view in the reference,
    scene/cameras.py:63, so contiguity must be enforced here)."""
    if t.numel() == 0:
        return t
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (this rasterizer has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Workspace:
    """Caller-owned scratch of ONE view slot: the geometry / binning / image buffers, the blend-backward
    accumulator and the colour / depth / radii outputs live in tensors that are kept between calls and only
    ever grow, so a steady-state multi-view loop never reaches the caching allocator (with several steps in
    flight on several streams its cross-stream reuse rules otherwise end in cudaMalloc -- a device
    synchronisation -- inside the loop).  Contract: a workspace may be handed to the next call only when the
    work that used it last is stream-ordered before that call (ViewPipeline guarantees this for one workspace
    per view of a step); tensors returned from a call with a workspace are views into it and are overwritten
    by the next call."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.bufs: dict = {}

    def bytes(self, name: str, nbytes: int) -> torch.Tensor:
        t = self.bufs.get(name)
        if t is None or t.numel() < nbytes:
            t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self.bufs[name] = t
        return t[:nbytes]

    def tensor(self, name: str, shape, dtype) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        return self.bytes(name, n * torch.empty(0, dtype=dtype).element_size()).view(dtype).view(*shape)

    def reserved_bytes(self) -> int:
        return sum(t.numel() for t in self.bufs.values())


class _Grower:
    """The reference's resizeFunctional(): a callback that (re)allocates a uint8 tensor.

    The ctypes thunk references a bound method of this object, i.e. a reference cycle that would
    keep the (hundreds of MB) buffer alive until the next cyclic GC pass -- every view would then
    force the caching allocator into a fresh cudaMalloc.  `take()` breaks the cycle as soon as the
    native call has returned, so the buffers are freed by reference counting."""

    def __init__(self, device, workspace=None, name=None):
        self.device = device
        self.workspace, self.name = workspace, name
        self.tensor = torch.empty(0, dtype=torch.uint8, device=device)
        self.cb = _ALLOC_FN(self._alloc)

    def _alloc(self, _user, nbytes):
        try:
            if self.workspace is not None:
                self.tensor = self.workspace.bytes(self.name, int(nbytes))
            else:
                self.tensor = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            return self.tensor.data_ptr()
        except Exception:  # surfaces as GSR_E_ALLOC
            return 0

    def take(self) -> torch.Tensor:
        t, self.tensor, self.cb, self.workspace = self.tensor, None, None, None
        return t


#: running high-water mark of num_rendered per device -> capacity hint of the next call, so that the
#: binning + blend kernels are queued before N reaches the host (no GPU bubble).  GSR_SPECULATE=0
#: restores the reference's exact-size behaviour (one mid-pipeline host round trip).
_SPECULATE = os.environ.get("GSR_SPECULATE", "1") != "0"
_hwm: dict = {}


def capacity_hint(device) -> int:
    h = _hwm.get(device.index, 0)
    return int(h * 1.25) + 65536 if (h > 0 and _SPECULATE) else 0


def _note_rendered(device, n: int):
    _hwm[device.index] = max(int(n), int(_hwm.get(device.index, 0) * 0.97))


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                        cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                        image_width, sh, degree, campos, prefiltered, debug=False, flags=None,
                        capacity=None, async_result=None, workspace=None):
    """`workspace` (extension): a Workspace that provides every buffer and output of this call (see there).
    `capacity` / `async_result` (extensions): with `async_result` (a pinned int64[2] tensor) the
    call never blocks the host -- GSR_FLAG_ASYNC -- and returns num_rendered = -1; the caller reads
    async_result after a stream sync ([0] = N must be <= capacity, [1] must be 0)."""
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    flags = DEFAULT_FLAGS if flags is None else int(flags)
    P, H, W = means3D.shape[0], int(image_height), int(image_width)
    dev = means3D.device
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor (this rasterizer has no CPU path)")
    M = 0
    if sh.numel() != 0:
        M = sh.shape[1]
    with torch.cuda.device(dev):
        if P == 0:
            # the reference returns zero-filled outputs and rendered = 0 without launching
            z = torch.zeros
            e = torch.empty(0, dtype=torch.uint8, device=dev)
            return (0, z(3, H, W, device=dev), z(0, dtype=torch.int32, device=dev), e, e.clone(), e.clone(),
                    z(1, H, W, device=dev))
        background = _f32c(background, "background")
        means3D = _f32c(means3D, "means3D")
        colors, opacity = _f32c(colors, "colors_precomp"), _f32c(opacity, "opacities")
        scales, rotations = _f32c(scales, "scales"), _f32c(rotations, "rotations")
        cov3D_precomp, sh = _f32c(cov3D_precomp, "cov3D_precomp"), _f32c(sh, "shs")
        viewmatrix, projmatrix, campos = _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix"), _f32c(campos, "campos")
        if workspace is not None:
            out_color = workspace.tensor("color", (3, H, W), torch.float32)
            out_depth = workspace.tensor("depth", (1, H, W), torch.float32)
            radii = workspace.tensor("radii", (P,), torch.int32)
        else:
            out_color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
            out_depth = torch.empty(1, H, W, dtype=torch.float32, device=dev)
            radii = torch.empty(P, dtype=torch.int32, device=dev)
        geom, binning, img = _Grower(dev, workspace, "geom"), _Grower(dev, workspace, "binning"), _Grower(dev, workspace, "image")
        n = _i64(0)
        if async_result is not None:
            assert async_result.is_pinned() and async_result.dtype == torch.int64 and async_result.numel() >= 2
            assert capacity is not None and capacity > 0
            flags |= FLAG_ASYNC
            n_ptr = async_result.data_ptr()
        else:
            n_ptr = C.addressof(n)
            if capacity is None:
                capacity = capacity_hint(dev)
        rc = _lib.gsr_forward(
            _stream(dev), geom.cb, None, binning.cb, None, img.cb, None, P, int(degree), M,
            _ptr(background), W, H, _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity), _ptr(scales),
            float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix),
            _ptr(projmatrix), _ptr(campos), float(tan_fovx), float(tan_fovy), int(bool(prefiltered)),
            _ptr(out_color), _ptr(out_depth), _ptr(radii), n_ptr, int(capacity), flags)
        geom, binning, img = geom.take(), binning.take(), img.take()
        _check(rc, "rasterize_gaussians")
    if async_result is not None:
        return -1, out_color, radii, geom, binning, img, out_depth
    _note_rendered(dev, n.value)
    return int(n.value), out_color, radii, geom, binning, img, out_depth


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                 dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer,
                                 imageBuffer, debug=False, flags=None, return_conic=False, out=None):
    """`out` (optional): dict of preallocated gradient tensors keyed dL_dmeans3D, dL_dsh, dL_dcolors,
    dL_dopacity, dL_dcov3D, dL_dscales, dL_drotations -- with FLAG_ACCUMULATE they are added into."""
    flags = DEFAULT_FLAGS if flags is None else int(flags)
    P = means3D.shape[0]
    H, W = dL_dout_color.shape[1], dL_dout_color.shape[2]
    dev = means3D.device
    M = sh.shape[1] if sh.numel() != 0 else 0
    with torch.cuda.device(dev):
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        out = out or {}
        acc = bool(flags & FLAG_ACCUMULATE)

        def o(name, *shape):
            t = out.get(name)
            if t is None:
                return torch.zeros(*shape, dtype=torch.float32, device=dev) if acc else e(*shape)
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == math.prod(shape), name
            return t
        dL_dmeans2D, dL_dconic = e(P, 3), (e(P, 2, 2) if return_conic else None)
        dL_dmeans3D, dL_dcolors = o("dL_dmeans3D", P, 3), o("dL_dcolors", P, 3)
        dL_dopacity, dL_dcov3D = o("dL_dopacity", P, 1), o("dL_dcov3D", P, 6)
        dL_dsh, dL_dscales, dL_drotations = o("dL_dsh", P, M, 3), o("dL_dscales", P, 3), o("dL_drotations", P, 4)
        if P != 0:
            background = _f32c(background, "background")
            means3D = _f32c(means3D, "means3D")
            colors, scales, rotations = _f32c(colors, "colors_precomp"), _f32c(scales, "scales"), _f32c(rotations, "rotations")
            cov3D_precomp, sh = _f32c(cov3D_precomp, "cov3D_precomp"), _f32c(sh, "shs")
            viewmatrix, projmatrix, campos = _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix"), _f32c(campos, "campos")
            dL_dout_color = _f32c(dL_dout_color, "dL_dout_color")
            nscratch = int(_lib.gsr_backward_scratch_bytes(P))
            scratch = torch.empty(nscratch, dtype=torch.uint8, device=dev)
            rc = _lib.gsr_backward(
                _stream(dev), P, int(degree), M, max(int(R), 0), _ptr(background), W, H, _ptr(means3D), _ptr(sh),
                _ptr(colors), _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp),
                _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), float(tan_fovx), float(tan_fovy),
                _ptr(radii), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer), _ptr(dL_dout_color),
                _ptr(dL_dmeans2D), _ptr(dL_dconic), _ptr(dL_dopacity), _ptr(dL_dcolors), _ptr(dL_dmeans3D),
                _ptr(dL_dcov3D), _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drotations), _ptr(scratch),
                nscratch, flags)
            _check(rc, "rasterize_gaussians_backward")
    out = (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)
    return out + (dL_dconic,) if return_conic else out


def mark_visible(means3D, viewmatrix, projmatrix):
    P = means3D.shape[0]
    dev = means3D.device
    present = torch.zeros(P, dtype=torch.bool, device=dev)
    if P != 0:
        with torch.cuda.device(dev):
            means3D, viewmatrix, projmatrix = _f32c(means3D, "means3D"), _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix")
            _check(_lib.gsr_mark_visible(_stream(dev), P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix),
                                         present.data_ptr()), "mark_visible")
    return present


def dist_cuda2(points):
    """`simple_knn._C.distCUDA2(points) -> float32[P]`: mean squared distance of every point to its three
    nearest other points (reference call sites gs-simp/scene/gaussian_model.py:134,546,623)."""
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    P = points.shape[0]
    dev = points.device
    points = _f32c(points, "points")
    out = torch.empty(P, dtype=torch.float32, device=dev)
    if P != 0:
        with torch.cuda.device(dev):
            nb = int(_lib.gsr_knn_temp_bytes(P))
            temp = torch.empty(nb, dtype=torch.uint8, device=dev)
            _check(_lib.gsr_knn3_mean_dist2(_stream(dev), P, _ptr(points), _ptr(out), _ptr(temp), nb), "knn3_mean_dist2")
    return out


def backward_blend(background, dL_dout_color, geomBuffer, binningBuffer, imageBuffer, P, flags=None, workspace=None):
    """K7 only (gsr_backward_blend): returns the packed per-Gaussian accumulator (uint8 scratch tensor)
    that backward_geom_multi() consumes."""
    flags = DEFAULT_FLAGS if flags is None else int(flags)
    dev = dL_dout_color.device
    H, W = dL_dout_color.shape[1], dL_dout_color.shape[2]
    with torch.cuda.device(dev):
        nscratch = int(_lib.gsr_backward_scratch_bytes(P))
        scratch = workspace.bytes("scratch", nscratch) if workspace is not None else torch.empty(nscratch, dtype=torch.uint8, device=dev)
        if P != 0:
            background, dL_dout_color = _f32c(background, "background"), _f32c(dL_dout_color, "dL_dout_color")
            _check(_lib.gsr_backward_blend(_stream(dev), P, _ptr(background), W, H, _ptr(geomBuffer), _ptr(binningBuffer),
                                           _ptr(imageBuffer), _ptr(dL_dout_color), _ptr(scratch), nscratch, flags),
                   "backward_blend")
    return scratch


def forward_views(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, views, sh, degree,
                  prefiltered, capacities, async_results, workspaces=None, flags=None, scratches_out=None):
    """Batched forward (gsr_forward_views): every pipeline stage is ONE launch for all `views` of the same Gaussians.
    views: list of dicts / settings objects with viewmatrix, projmatrix, campos, tanfovx, tanfovy, image_height,
    image_width (and optionally bg).  capacities[k] > 0 and async_results[k] (pinned int64[2]) per view; nothing
    blocks the host.  -> list of (-1, color, radii, geom, binning, img, depth) like rasterize_gaussians.
    scratches_out: pass an empty list when a backward follows -- it receives one accumulator buffer per view that K1 has
    already cleared; hand it to backward_blend_views(scratches=...), which then skips its own clear."""
    flags = DEFAULT_FLAGS if flags is None else int(flags)
    P = means3D.shape[0]
    dev = means3D.device
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor (this rasterizer has no CPU path)")
    M = sh.shape[1] if sh.numel() != 0 else 0
    nv = len(views)
    get = lambda v, k: v[k] if isinstance(v, dict) else getattr(v, k)
    keep, out = [], []
    with torch.cuda.device(dev):
        background = _f32c(background, "background")
        means3D = _f32c(means3D, "means3D")
        colors, opacity = _f32c(colors, "colors_precomp"), _f32c(opacity, "opacities")
        scales, rotations = _f32c(scales, "scales"), _f32c(rotations, "rotations")
        cov3D_precomp, sh = _f32c(cov3D_precomp, "cov3D_precomp"), _f32c(sh, "shs")
        arr = (GsrViewForward * max(nv, 1))()
        for k, v in enumerate(views):
            H, W = int(get(v, "image_height")), int(get(v, "image_width"))
            cap = int(capacities[k])
            res = async_results[k]
            assert res.is_pinned() and res.dtype == torch.int64 and res.numel() >= 2 and cap > 0
            lay = get_layout(P, W, H, cap, flags)
            ws = workspaces[k] if workspaces is not None else None
            if ws is not None:
                color, depth = ws.tensor("color", (3, H, W), torch.float32), ws.tensor("depth", (1, H, W), torch.float32)
                radii = ws.tensor("radii", (P,), torch.int32)
                geom, binning, img = ws.bytes("geom", lay.geom_bytes), ws.bytes("binning", lay.binning_bytes), ws.bytes("image", lay.image_bytes)
            else:
                color, depth = torch.empty(3, H, W, device=dev), torch.empty(1, H, W, device=dev)
                radii = torch.empty(P, dtype=torch.int32, device=dev)
                geom, binning, img = (torch.empty(int(n), dtype=torch.uint8, device=dev)
                                      for n in (lay.geom_bytes, lay.binning_bytes, lay.image_bytes))
            vm, pm, cp = _f32c(get(v, "viewmatrix"), "viewmatrix"), _f32c(get(v, "projmatrix"), "projmatrix"), _f32c(get(v, "campos"), "campos")
            bgk = background
            keep.append((vm, pm, cp, bgk))
            g = arr[k]
            g.background, g.viewmatrix, g.projmatrix, g.cam_pos = _ptr(bgk), _ptr(vm), _ptr(pm), _ptr(cp)
            g.tan_fovx, g.tan_fovy, g.width, g.height = float(get(v, "tanfovx")), float(get(v, "tanfovy")), W, H
            g.out_color, g.out_depth, g.radii = color.data_ptr(), depth.data_ptr(), (radii.data_ptr() if P else None)
            g.geom_buffer, g.binning_buffer, g.image_buffer = geom.data_ptr(), binning.data_ptr(), img.data_ptr()
            g.capacity, g.result_host = cap, res.data_ptr()
            g.backward_scratch = None
            if scratches_out is not None and P:
                nscratch = int(_lib.gsr_backward_scratch_bytes(P))
                sc = ws.bytes("scratch", nscratch) if ws is not None else torch.empty(nscratch, dtype=torch.uint8, device=dev)
                scratches_out.append(sc)
                g.backward_scratch = sc.data_ptr()
            out.append((-1, color, radii, geom, binning, img, depth))
        if nv:
            _check(_lib.gsr_forward_views(_stream(dev), P, int(degree), M, _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity),
                                          _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp),
                                          int(bool(prefiltered)), arr, nv, flags), "forward_views")
    return out


def backward_blend_views(background, dL_dout_colors, geoms, binnings, imgs, P, flags=None, workspaces=None, scratches=None):
    """K7 of several views in one launch (gsr_backward_blend_views) -> list of packed accumulators (uint8 scratch).
    scratches: the buffers forward_views(scratches_out=...) handed out (already cleared by K1)."""
    flags = DEFAULT_FLAGS if flags is None else int(flags)
    if scratches:
        assert len(scratches) == len(dL_dout_colors)
        flags |= FLAG_SCRATCH_CLEARED
    nv = len(dL_dout_colors)
    if nv == 0:
        return []
    dev = dL_dout_colors[0].device
    with torch.cuda.device(dev):
        nscratch = int(_lib.gsr_backward_scratch_bytes(P))
        background = _f32c(background, "background")
        arr = (GsrViewBackward * nv)()
        given = scratches if scratches else None
        keep, scratches = [], []
        for k in range(nv):
            dL = _f32c(dL_dout_colors[k], "dL_dout_color")
            keep.append(dL)
            if given:
                sc = given[k]
            else:
                sc = workspaces[k].bytes("scratch", nscratch) if workspaces is not None else torch.empty(nscratch, dtype=torch.uint8, device=dev)
            scratches.append(sc)
            g = arr[k]
            g.background, g.width, g.height = _ptr(background), int(dL.shape[2]), int(dL.shape[1])
            g.geom_buffer, g.binning_buffer, g.image_buffer = _ptr(geoms[k]), _ptr(binnings[k]), _ptr(imgs[k])
            g.dL_dpix, g.scratch, g.scratch_bytes = dL.data_ptr(), sc.data_ptr(), nscratch
        if P != 0:
            _check(_lib.gsr_backward_blend_views(_stream(dev), P, arr, nv, flags), "backward_blend_views")
    return scratches


def backward_geom_multi_supported(M: int) -> bool:
    return M in (1, 4, 16)


def backward_geom_multi(means3D, sh, scales, rotations, scale_modifier, degree, views, out, stats=None, flags=0,
                        want_means2D=False, g_range=None):
    """Batched K8+K9 over the views of one set of Gaussians (gsr_backward_geom_multi).
    views: list of dicts with radii, geom, scratch, viewmatrix, projmatrix, campos, tanfovx, tanfovy, width, height.
    out:   dict of preallocated dL_dmeans3D (P,3), dL_dsh (P,M,3), dL_dopacity (P,1), dL_dscales (P,3),
           dL_drotations (P,4) -- written (or added to with FLAG_ACCUMULATE).
    stats: optional (grad_norm_accum f32[P], visible_count i32[P], max_radii i32[P]).
    g_range: optional (g_begin, g_end), g_begin % 4 == 0: only those Gaussians (gsr_backward_geom_multi_range).
    Returns the list of per-view dL_dmeans2D (P,3) tensors if want_means2D else None."""
    P, M = means3D.shape[0], sh.shape[1]
    g0, g1 = (0, P) if g_range is None else (int(g_range[0]), int(g_range[1]))
    dev = means3D.device
    keep = []   # contiguous copies must outlive the launch

    def c(t, name):
        t = _f32c(t, name)
        keep.append(t)
        return t
    with torch.cuda.device(dev):
        means3D, sh, scales, rotations = c(means3D, "means3D"), c(sh, "shs"), c(scales, "scales"), c(rotations, "rotations")
        arr = (GsrViewGrad * len(views))()
        m2d = []
        for k, v in enumerate(views):
            g = arr[k]
            g.radii, g.geom_buffer, g.scratch = _ptr(v["radii"]), _ptr(v["geom"]), _ptr(v["scratch"])
            g.viewmatrix, g.projmatrix, g.cam_pos = (_ptr(c(v["viewmatrix"], "viewmatrix")), _ptr(c(v["projmatrix"], "projmatrix")),
                                                     _ptr(c(v["campos"], "campos")))
            g.tan_fovx, g.tan_fovy, g.width, g.height = float(v["tanfovx"]), float(v["tanfovy"]), int(v["width"]), int(v["height"])
            if want_means2D:
                t = torch.empty(P, 3, dtype=torch.float32, device=dev)
                m2d.append(t)
                g.dL_dmean2D = t.data_ptr()
            else:
                g.dL_dmean2D = None
        for name in ("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations"):
            t = out[name]
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), name
        st = stats or (None, None, None)
        if P != 0 and len(views) != 0:
            _check(_lib.gsr_backward_geom_multi_range(
                _stream(dev), P, int(degree), M, _ptr(means3D), _ptr(sh), _ptr(scales), float(scale_modifier), _ptr(rotations),
                arr, len(views), _ptr(out["dL_dopacity"]), _ptr(out["dL_dmeans3D"]), _ptr(out["dL_dsh"]), _ptr(out["dL_dscales"]),
                _ptr(out["dL_drotations"]), _ptr(st[0]), _ptr(st[1]), _ptr(st[2]), int(flags), g0, g1), "backward_geom_multi")
    return m2d if want_means2D else None


def nvls_all_reduce(multicast_ptr: int, device, off_f32: int, n_f32: int, off_add_s32: int, n_add_s32: int,
                    off_max_s32: int, n_max_s32: int, rank: int, world: int, blocks: int = 0,
                    sparse_first_f32: int = 0, sparse_rows: int = 0, sparse_row_f32: int = 0):
    """gsr_nvls_all_reduce on the current stream of `device` (see include/gsrast_b200.h)."""
    with torch.cuda.device(device):
        _check(_lib.gsr_nvls_all_reduce(_stream(device), multicast_ptr, off_f32, n_f32, off_add_s32, n_add_s32,
                                        off_max_s32, n_max_s32, rank, world, blocks, sparse_first_f32, sparse_rows,
                                        sparse_row_f32), "nvls_all_reduce")


def nvls_all_reduce_plan(multicast_ptr: int, device, rank: int, world: int, dense=(), rows=None, add_s32=(0, 0),
                         max_s32=(0, 0), blocks: int = 0):
    """gsr_nvls_all_reduce_plan on the current stream.  dense: [(byte offset, n floats)], rows: (byte offset, n rows,
    floats per row, byte offset of row 0's int32 count) or None, add_s32 / max_s32: (byte offset, n)."""
    pl = GsrNvlsPlan()
    pl.n_dense = len(dense)
    for k, (off, n) in enumerate(dense):
        pl.dense_off[k], pl.dense_n_f32[k] = int(off), int(n)
    if rows is not None:
        pl.rows_off, pl.rows, pl.row_f32, pl.rows_count_off = int(rows[0]), int(rows[1]), int(rows[2]), int(rows[3])
    pl.add_s32_off, pl.n_add_s32 = int(add_s32[0]), int(add_s32[1])
    pl.max_s32_off, pl.n_max_s32 = int(max_s32[0]), int(max_s32[1])
    with torch.cuda.device(device):
        _check(_lib.gsr_nvls_all_reduce_plan(_stream(device), multicast_ptr, C.byref(pl), rank, world, blocks), "nvls_all_reduce_plan")


def p2p_all_reduce_plan(local_ptr: int, peer_ptr: int, device, rank: int, dense=(), rows=None, add_s32=(0, 0), max_s32=(0, 0),
                        blocks: int = 0):
    """gsr_p2p_all_reduce_plan on the current stream: the plan of nvls_all_reduce_plan between exactly two ranks, by peer
    loads / stores over NVLink (local_ptr / peer_ptr: this rank's and the peer's replica of the arena)."""
    pl = GsrNvlsPlan()
    pl.n_dense = len(dense)
    for k, (off, n) in enumerate(dense):
        pl.dense_off[k], pl.dense_n_f32[k] = int(off), int(n)
    if rows is not None:
        pl.rows_off, pl.rows, pl.row_f32, pl.rows_count_off = int(rows[0]), int(rows[1]), int(rows[2]), int(rows[3])
    pl.add_s32_off, pl.n_add_s32 = int(add_s32[0]), int(add_s32[1])
    pl.max_s32_off, pl.n_max_s32 = int(max_s32[0]), int(max_s32[1])
    with torch.cuda.device(device):
        _check(_lib.gsr_p2p_all_reduce_plan(_stream(device), int(local_ptr), int(peer_ptr), C.byref(pl), rank, blocks), "p2p_all_reduce_plan")


def accumulate_view_stats(radii, dL_dmeans2D, grad_norm_accum=None, visible_count=None, max_radii=None):
    """One fused kernel for the densification statistics of a view (gaussian_model.py:482-484,
    train.py:115): grad_norm_accum[vis] += ||dL_dmeans2D[vis, :2]||, visible_count[vis] += 1,
    max_radii[vis] = max(max_radii[vis], radii[vis]) with vis = radii > 0."""
    P = radii.shape[0]
    dev = radii.device
    for t, dt in ((radii, torch.int32), (dL_dmeans2D, torch.float32), (grad_norm_accum, torch.float32),
                  (visible_count, torch.int32), (max_radii, torch.int32)):
        if t is not None and not (t.is_cuda and t.dtype == dt and t.is_contiguous()):
            raise RuntimeError("accumulate_view_stats: tensors must be contiguous CUDA int32 / float32")
    if P != 0:
        with torch.cuda.device(dev):
            _check(_lib.gsr_accumulate_view_stats(_stream(dev), P, _ptr(radii), _ptr(dL_dmeans2D), _ptr(grad_norm_accum),
                                                  _ptr(visible_count), _ptr(max_radii)), "accumulate_view_stats")


# ---- measurement hooks ---------------------------------------------------------------------------

def debug_set(knob: int, value: int):
    """Experiment switches of the library (gsr_debug_set); never used by the product path."""
    _check(_lib.gsr_debug_set(int(knob), int(value)), "gsr_debug_set")


if os.environ.get("GSR_DEBUG_KNOBS"):   # experiments only (tools/gpu_knobs.sh): "knob=value,knob=value"
    for _kv in os.environ["GSR_DEBUG_KNOBS"].split(","):
        debug_set(*[int(x) for x in _kv.split("=")])


def kernel_launches() -> int:
    return int(_lib.gsr_kernel_launches())


def debug_approx_units(x: torch.Tensor) -> torch.Tensor:
    """(n,) float32 CUDA -> (n, 2): [rcp.approx.ftz(x), ex2.approx.ftz(x)] as the default blend evaluates them."""
    x = x.contiguous().float()
    out = torch.empty(x.numel(), 2, dtype=torch.float32, device=x.device)
    _lib.gsr_debug_approx_units.restype = _i
    _lib.gsr_debug_approx_units.argtypes = [C.c_void_p, _i, C.c_void_p, C.c_void_p]
    _check(_lib.gsr_debug_approx_units(_ptr(x), x.numel(), _ptr(out), _stream(x.device)), "gsr_debug_approx_units")
    return out


def profile_enable(on: bool):
    _lib.gsr_profile_enable(int(bool(on)))


def profile_collect():
    """-> ({stage: total ms}, {stage: brackets}) accumulated since the last collect."""
    ms = (C.c_double * NUM_STAGES)()
    cnt = (_i64 * NUM_STAGES)()
    _check(_lib.gsr_profile_collect(ms, cnt), "gsr_profile_collect")
    return ({STAGE_NAMES[i]: ms[i] for i in range(NUM_STAGES)}, {STAGE_NAMES[i]: int(cnt[i]) for i in range(NUM_STAGES)})


# ---- stage-level access for parity tests (not part of the reference's _C surface) ----------------

def get_layout(P, W, H, num_rendered, flags=None) -> GsrLayout:
    flags = DEFAULT_FLAGS if flags is None else int(flags)
    lay = GsrLayout()
    _check(_lib.gsr_get_layout(int(P), int(W), int(H), int(num_rendered), flags, C.byref(lay)), "gsr_get_layout")
    return lay


def _view(buf: torch.Tensor, off: int, count: int, dtype) -> torch.Tensor:
    nbytes = count * torch.empty(0, dtype=dtype).element_size()
    return buf[off:off + nbytes].view(dtype)


def unpack_state(P, W, H, num_rendered, geomBuffer, binningBuffer, imgBuffer, flags=None) -> dict:
    """Typed views of the opaque buffers (Appendix A.8 names)."""
    lay = get_layout(P, W, H, num_rendered, flags)
    G = ((W + 15) // 16) * ((H + 15) // 16)
    rec = _view(geomBuffer, lay.rec, 12 * P, torch.float32).view(P, 12)
    out = dict(
        means2D=rec[:, 0:2], conic_opacity=torch.stack([rec[:, 2], rec[:, 3], rec[:, 4], rec[:, 5]], 1),
        rgb=rec[:, 8:11], cull=torch.stack([rec[:, 6], rec[:, 7], rec[:, 11]], 1),
        depths=_view(geomBuffer, lay.depths, P, torch.float32),
        clamped=_view(geomBuffer, lay.clamped, P, torch.uint8),
        tiles_touched=_view(geomBuffer, lay.tiles_touched, P, torch.int32),
        point_offsets=_view(geomBuffer, lay.point_offsets, P, torch.int32),
        final_T=_view(imgBuffer, lay.final_T, H * W, torch.float32).view(H, W),
        n_contrib=_view(imgBuffer, lay.n_contrib, H * W, torch.int32).view(H, W),
        ranges=_view(imgBuffer, lay.ranges, 2 * G, torch.int32).view(G, 2),
        point_list=_view(binningBuffer, lay.point_list, num_rendered, torch.int32),
    )
    if lay.order != C.c_size_t(-1).value:
        out["order"] = _view(geomBuffer, lay.order, P, torch.int32)
    return out


def sort_pairs(keys: torch.Tensor, vals: torch.Tensor, end_bit: int):
    """gsr_sort_pairs_u64 / _u32.  keys: int64 (bit pattern of u64) or int32 (u32); vals int32."""
    n = keys.numel()
    dev = keys.device
    kb = 8 if keys.dtype == torch.int64 else 4
    keys_out, vals_out = torch.empty_like(keys), torch.empty_like(vals)
    with torch.cuda.device(dev):
        nt = int(_lib.gsr_sort_temp_bytes(n, kb, end_bit))
        temp = torch.empty(nt, dtype=torch.uint8, device=dev)
        fn = _lib.gsr_sort_pairs_u64 if kb == 8 else _lib.gsr_sort_pairs_u32
        _check(fn(_stream(dev), n, _ptr(keys), _ptr(vals), _ptr(keys_out), _ptr(vals_out), int(end_bit),
                  _ptr(temp), nt), "sort_pairs")
    return keys_out, vals_out


def inclusive_scan(x: torch.Tensor, gather: torch.Tensor | None = None) -> torch.Tensor:
    n = x.numel() if gather is None else gather.numel()
    dev = x.device
    out = torch.empty(n, dtype=x.dtype, device=dev)
    with torch.cuda.device(dev):
        nt = int(_lib.gsr_scan_temp_bytes(n))
        temp = torch.empty(max(nt, 1), dtype=torch.uint8, device=dev)
        _check(_lib.gsr_inclusive_scan_u32(_stream(dev), n, _ptr(x), _ptr(gather), _ptr(out), _ptr(temp), nt),
               "inclusive_scan")
    return out


# ---- the training step either side of the rasterizer -------------------------
def _req_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (no CPU path)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous float32")
    return t


def loss_temp_bytes(C_: int, H: int, W: int) -> int:
    return int(_lib.gsr_loss_temp_bytes(int(C_), int(H), int(W)))


def loss_l1_ssim_forward(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float, out3: torch.Tensor | None = None,
                         temp: torch.Tensor | None = None):
    """-> (out3, temp): out3 = device float32[3] {Ll1, ssim, loss} (utils/loss_utils.py:17-18,33-62 composed as
    train.py:91-92); temp holds the derivative maps loss_l1_ssim_backward needs."""
    image, gt = _req_cuda_f32(image, "image"), _req_cuda_f32(gt, "gt")
    if image.ndim != 3 or image.shape != gt.shape:
        raise RuntimeError(f"image and gt must both be (C,H,W), got {tuple(image.shape)} and {tuple(gt.shape)}")
    Cc, H, W = image.shape
    dev = image.device
    nb = loss_temp_bytes(Cc, H, W)
    with torch.cuda.device(dev):
        if temp is None or temp.numel() < nb:
            temp = torch.empty(nb, dtype=torch.uint8, device=dev)
        if out3 is None:
            out3 = torch.empty(3, dtype=torch.float32, device=dev)
        _check(_lib.gsr_loss_l1_ssim_forward(_stream(dev), Cc, H, W, image.data_ptr(), gt.data_ptr(), float(lambda_dssim),
                                             out3.data_ptr(), temp.data_ptr(), temp.numel()), "loss_l1_ssim_forward")
    return out3, temp


def loss_l1_ssim_backward(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float, temp: torch.Tensor,
                          dL_dloss: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """d loss / d image (C,H,W), scaled by the DEVICE scalar dL_dloss (None = 1)."""
    image, gt = _req_cuda_f32(image, "image"), _req_cuda_f32(gt, "gt")
    Cc, H, W = image.shape
    dev = image.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty_like(image)
        g = None
        if dL_dloss is not None:
            g = _req_cuda_f32(dL_dloss.reshape(-1), "dL_dloss")
            assert g.numel() == 1
        _check(_lib.gsr_loss_l1_ssim_backward(_stream(dev), Cc, H, W, image.data_ptr(), gt.data_ptr(), float(lambda_dssim),
                                              _ptr(g), temp.data_ptr(), temp.numel(), _req_cuda_f32(out, "out").data_ptr()),
               "loss_l1_ssim_backward")
    return out


def activate_forward(raw_scales, raw_rotations, raw_opacities, scales=None, rotations=None, opacities=None):
    """exp / normalize / sigmoid of GaussianModel's getters (gaussian_model.py:95-115) in one kernel.
    Any raw_* may be None.  -> (scales, rotations, opacities)"""
    ref = next(t for t in (raw_scales, raw_rotations, raw_opacities) if t is not None)
    dev, P = ref.device, ref.shape[0]
    outs = []
    for raw, o, name in ((raw_scales, scales, "scales"), (raw_rotations, rotations, "rotations"), (raw_opacities, opacities, "opacities")):
        if raw is None:
            outs.append(None)
            continue
        _req_cuda_f32(raw, "raw_" + name)
        assert raw.shape[0] == P
        outs.append(_req_cuda_f32(o, name) if o is not None else torch.empty_like(raw))
    with torch.cuda.device(dev):
        _check(_lib.gsr_activate_forward(_stream(dev), P, _ptr(raw_scales), _ptr(raw_rotations), _ptr(raw_opacities),
                                         _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2])), "activate_forward")
    return tuple(outs)


def activate_backward(raw_scales, raw_rotations, raw_opacities, g_scales, g_rotations, g_opacities):
    """In place: g_* = dL/d(activated) -> dL/d(raw)."""
    ref = next(t for t in (raw_scales, raw_rotations, raw_opacities) if t is not None)
    dev, P = ref.device, ref.shape[0]
    for raw, g, name in ((raw_scales, g_scales, "scales"), (raw_rotations, g_rotations, "rotations"), (raw_opacities, g_opacities, "opacities")):
        if raw is not None:
            _req_cuda_f32(raw, "raw_" + name), _req_cuda_f32(g, "g_" + name)
            assert raw.numel() == g.numel() and raw.shape[0] == P
    with torch.cuda.device(dev):
        _check(_lib.gsr_activate_backward(_stream(dev), P, _ptr(raw_scales), _ptr(raw_rotations), _ptr(raw_opacities),
                                          _ptr(g_scales if raw_scales is not None else None),
                                          _ptr(g_rotations if raw_rotations is not None else None),
                                          _ptr(g_opacities if raw_opacities is not None else None)), "activate_backward")


def adam_step(segments, step: int, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-15):
    """One launch of Adam over `segments`: dicts with param, grad, exp_avg, exp_avg_sq (same numel, contiguous CUDA
    float32), lr, and optionally row_len / row_split / lr_rest (see gsr_adam_step).  Defaults are those of
    gaussian_model.py:165 (eps=1e-15) and torch.optim.Adam (betas)."""
    segs = (GsrAdamSegment * max(len(segments), 1))()
    dev = None
    keep = []
    for k, sg in enumerate(segments):
        p, g, m, v = (_req_cuda_f32(sg[n], n) for n in ("param", "grad", "exp_avg", "exp_avg_sq"))
        assert p.numel() == g.numel() == m.numel() == v.numel(), "segment tensors differ in size"
        dev = p.device if dev is None else dev
        keep.append((p, g, m, v))
        segs[k] = GsrAdamSegment(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(sg["lr"]),
                                 float(sg.get("lr_rest", sg["lr"])), int(sg.get("row_len", 0)), int(sg.get("row_split", 0)))
    if dev is None:
        return
    with torch.cuda.device(dev):
        _check(_lib.gsr_adam_step(_stream(dev), segs, len(segments), int(step), float(beta1), float(beta2), float(eps)),
               "adam_step")


def quantize_rgb8(image: torch.Tensor, out: torch.Tensor | None = None, affine: torch.Tensor | None = None) -> torch.Tensor:
    """(C,H,W) float32, C in {1,3} -> (H,W,3) uint8 exactly as torchvision.utils.save_image quantises
    (gsr_quantize_rgb8); `affine` = device float[2] {lo, inv_range} applies (x - lo) * inv_range first."""
    image = _req_cuda_f32(image, "image")
    if image.ndim != 3 or image.shape[0] not in (1, 3):
        raise RuntimeError(f"image must be (1|3,H,W), got {tuple(image.shape)}")
    Cc, H, W = image.shape
    dev = image.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty(H, W, 3, dtype=torch.uint8, device=dev)
        assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and out.numel() == H * W * 3
        if affine is not None:
            affine = _req_cuda_f32(affine, "affine")
            assert affine.numel() == 2
        _check(_lib.gsr_quantize_rgb8(_stream(dev), Cc, H, W, image.data_ptr(), _ptr(affine), out.data_ptr()), "quantize_rgb8")
    return out


def gather_rows(src_row: torch.Tensor, n_src: int, segments, n_keep_state: int | None = None):
    """ONE launch of gsr_gather_rows: for every segment dict(src=(n_src, ...), dst=(n_dst, ...), zero_new=bool) sets
    dst[i] = src[src_row[i]], or zeros for rows i >= n_keep_state of a zero_new segment (the Adam moments of cloned /
    split Gaussians).  src_row: CUDA int32 (n_dst,), entries in [0, n_src) -- the caller guarantees the range
    (densify.py builds it from masks).  All tensors contiguous CUDA float32; src and dst must not alias."""
    if not src_row.is_cuda or src_row.dtype != torch.int32 or not src_row.is_contiguous():
        raise RuntimeError("src_row must be a contiguous CUDA int32 tensor (this library has no CPU path)")
    dev, n_dst = src_row.device, src_row.numel()
    n_keep_state = n_dst if n_keep_state is None else int(n_keep_state)
    if len(segments) > GATHER_MAX_SEGS:
        raise RuntimeError(f"at most {GATHER_MAX_SEGS} segments per launch")
    segs = (GsrGatherSegment * max(len(segments), 1))()
    for k, sg in enumerate(segments):
        src, dst = _req_cuda_f32(sg["src"], "src"), _req_cuda_f32(sg["dst"], "dst")
        rs = src.numel() // n_src if n_src else 0
        rd = dst.numel() // n_dst if n_dst else rs
        if src.numel() != rs * n_src or dst.numel() != rd * n_dst or (n_src and n_dst and rs != rd):
            raise RuntimeError(f"segment {k}: src {tuple(src.shape)} / dst {tuple(dst.shape)} do not have {n_src} / {n_dst} "
                               "rows of equal length")
        segs[k] = GsrGatherSegment(_ptr(src), _ptr(dst), rd, 1 if sg.get("zero_new") else 0)
    with torch.cuda.device(dev):
        _check(_lib.gsr_gather_rows(_stream(dev), n_dst, int(n_src), n_keep_state, _ptr(src_row), segs, len(segments)),
               "gather_rows")
