"""ctypes front-end of the CPU oracle (oracle/gsplat_oracle.c).

TEST INFRASTRUCTURE ONLY -- "parity unpinned" for the whole pipeline (see gsplat_oracle.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product package never does.

The call structure mirrors what the reference drives through
gs-simp/gaussian_renderer/__init__.py:85-93 (forward) and loss.backward() (train.py:93):
`forward()` returns every intermediate of SURVEY.md Appendix A so tests can compare stage by stage.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgsplat_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "gsplat_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.gso_inclusive_scan.restype = C.c_int64
        _lib.gso_higher_msb.restype = C.c_uint32
        _lib.gso_num_threads.restype = C.c_int
        _lib.gso_preprocess.restype = C.c_int
    return _lib


def _p(a):
    if a is None:
        return C.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def num_threads() -> int:
    return int(lib().gso_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the oracle from here on (bench.py's reference arm: all host cores, whatever
    OMP_NUM_THREADS the launcher exported).  Returns the count in effect."""
    lib().gso_set_num_threads(C.c_int(int(n)))
    return num_threads()


def higher_msb(n: int) -> int:
    return int(lib().gso_higher_msb(C.c_uint32(n)))


@dataclass
class Forward:
    """All intermediates of one forward pass (SURVEY Appendix A.8 names)."""
    P: int
    W: int
    H: int
    radii: np.ndarray = None
    means2D: np.ndarray = None
    depths: np.ndarray = None
    cov3D: np.ndarray = None
    rgb: np.ndarray = None
    conic_opacity: np.ndarray = None
    tiles_touched: np.ndarray = None
    clamped: np.ndarray = None
    point_offsets: np.ndarray = None
    num_rendered: int = 0
    keys_unsorted: np.ndarray = None
    values_unsorted: np.ndarray = None
    keys_sorted: np.ndarray = None
    point_list: np.ndarray = None
    ranges: np.ndarray = None
    color: np.ndarray = None
    depth: np.ndarray = None
    final_T: np.ndarray = None
    n_contrib: np.ndarray = None
    inputs: dict = field(default_factory=dict)


def preprocess(means3D, opacities, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy,
               sh_degree=0, shs=None, colors_precomp=None, scales=None, rotations=None,
               cov3D_precomp=None, scale_modifier=1.0, prefiltered=False) -> Forward:
    means3D = _f32(means3D)
    P = means3D.shape[0]
    shs, colors_precomp = _f32(shs), _f32(colors_precomp)
    scales, rotations, cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
    opacities = _f32(opacities).reshape(-1)
    viewmatrix, projmatrix, campos = _f32(viewmatrix), _f32(projmatrix), _f32(campos)
    M = 0 if shs is None else shs.shape[1]
    f = Forward(P=P, W=W, H=H)
    f.inputs = dict(means3D=means3D, opacities=opacities, shs=shs, colors_precomp=colors_precomp,
                    scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp,
                    viewmatrix=viewmatrix, projmatrix=projmatrix, campos=campos,
                    tanfovx=float(tanfovx), tanfovy=float(tanfovy), sh_degree=int(sh_degree), M=M,
                    scale_modifier=float(scale_modifier))
    f.radii = np.zeros(P, np.int32)
    f.means2D = np.zeros((P, 2), np.float32)
    f.depths = np.zeros(P, np.float32)
    f.cov3D = np.zeros((P, 6), np.float32) if cov3D_precomp is None else cov3D_precomp.copy()
    f.rgb = np.zeros((P, 3), np.float32)
    f.conic_opacity = np.zeros((P, 4), np.float32)
    f.tiles_touched = np.zeros(P, np.uint32)
    f.clamped = np.zeros((P, 3), np.uint8)
    rc = lib().gso_preprocess(
        C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(scales),
        C.c_float(scale_modifier), _p(rotations), _p(opacities), _p(shs), _p(cov3D_precomp),
        _p(colors_precomp), _p(viewmatrix), _p(projmatrix), _p(campos), C.c_int(W), C.c_int(H),
        C.c_float(tanfovx), C.c_float(tanfovy), C.c_int(int(prefiltered)),
        _p(f.radii), _p(f.means2D), _p(f.depths), _p(f.cov3D), _p(f.rgb), _p(f.conic_opacity),
        _p(f.tiles_touched), _p(f.clamped))
    if rc != 0:
        raise RuntimeError("prefiltered set but a point was culled (the reference traps here)")
    return f


def bin_and_sort(f: Forward) -> Forward:
    P, W, H = f.P, f.W, f.H
    gx, gy = (W + 15) // 16, (H + 15) // 16
    f.point_offsets = np.zeros(P, np.uint32)
    f.num_rendered = int(lib().gso_inclusive_scan(C.c_int(P), _p(f.tiles_touched), _p(f.point_offsets))) if P else 0
    N = f.num_rendered
    f.keys_unsorted = np.zeros(N, np.uint64)
    f.values_unsorted = np.zeros(N, np.uint32)
    f.keys_sorted = np.zeros(N, np.uint64)
    f.point_list = np.zeros(N, np.uint32)
    f.ranges = np.zeros((gx * gy, 2), np.uint32)
    if N:
        lib().gso_duplicate_with_keys(C.c_int(P), _p(f.means2D), _p(f.depths), _p(f.point_offsets),
                                      _p(f.radii), C.c_int(W), C.c_int(H), _p(f.keys_unsorted),
                                      _p(f.values_unsorted))
        bit = higher_msb(gx * gy)
        lib().gso_sort_pairs(C.c_int64(N), _p(f.keys_unsorted), _p(f.values_unsorted),
                             _p(f.keys_sorted), _p(f.point_list), C.c_int(32 + bit))
        lib().gso_identify_tile_ranges(C.c_int64(N), _p(f.keys_sorted), C.c_int(gx * gy), _p(f.ranges))
    return f


def blend(f: Forward, bg, tile_start: int = 0, tile_step: int = 1) -> Forward:
    W, H = f.W, f.H
    bg = _f32(bg)
    f.inputs["bg"] = bg
    f.color = np.zeros((3, H, W), np.float32)
    f.depth = np.zeros((1, H, W), np.float32)
    f.final_T = np.zeros((H, W), np.float32)
    f.n_contrib = np.zeros((H, W), np.uint32)
    lib().gso_blend_forward_tiles(C.c_int(W), C.c_int(H), _p(f.ranges), _p(f.point_list), _p(f.means2D),
                                  _p(f.rgb), _p(f.depths), _p(f.conic_opacity), _p(bg), _p(f.color),
                                  _p(f.depth), _p(f.final_T), _p(f.n_contrib), C.c_int(tile_start),
                                  C.c_int(tile_step))
    return f


def forward(bg, **kw) -> Forward:
    """Full forward: K1 -> K2 -> K3 -> K4 -> K5 -> K6."""
    return blend(bin_and_sort(preprocess(**kw)), bg)


def backward(f: Forward, dL_dcolor, tile_start: int = 0, tile_step: int = 1) -> dict:
    """K7 -> K8 -> K9.  Returns the eight gradient tensors of
    `rasterize_gaussians_backward` plus dL_dconic."""
    P, W, H = f.P, f.W, f.H
    i = f.inputs
    M = i["M"]
    dL_dcolor = _f32(dL_dcolor).reshape(3, H, W)
    g = dict(
        dL_dmeans2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 4), np.float32),
        dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolors=np.zeros((P, 3), np.float32),
        dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
        dL_dsh=np.zeros((P, M, 3), np.float32), dL_dscales=np.zeros((P, 3), np.float32),
        dL_drotations=np.zeros((P, 4), np.float32))
    if P == 0:
        return g
    lib().gso_blend_backward_tiles(C.c_int(W), C.c_int(H), _p(f.ranges), _p(f.point_list), _p(f.means2D),
                                   _p(f.rgb), _p(f.conic_opacity), _p(i["bg"]), _p(f.final_T),
                                   _p(f.n_contrib), _p(dL_dcolor), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]),
                                   _p(g["dL_dopacity"]), _p(g["dL_dcolors"]), C.c_int(tile_start),
                                   C.c_int(tile_step))
    lib().gso_preprocess_backward(
        C.c_int(P), C.c_int(i["sh_degree"]), C.c_int(M), _p(i["means3D"]), _p(f.radii), _p(i["shs"]),
        _p(f.clamped), _p(i["scales"]), _p(i["rotations"]), C.c_float(i["scale_modifier"]),
        _p(f.cov3D), _p(i["viewmatrix"]), _p(i["projmatrix"]), _p(i["campos"]), C.c_int(W), C.c_int(H),
        C.c_float(i["tanfovx"]), C.c_float(i["tanfovy"]), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]),
        _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]),
        _p(g["dL_dscales"]), _p(g["dL_drotations"]))
    return g


def mark_visible(means3D, viewmatrix) -> np.ndarray:
    means3D, viewmatrix = _f32(means3D), _f32(viewmatrix)
    out = np.zeros(means3D.shape[0], np.uint8)
    lib().gso_mark_visible(C.c_int(means3D.shape[0]), _p(means3D), _p(viewmatrix), _p(out))
    return out.astype(bool)


def knn3_mean_dist2(points, queries=None) -> np.ndarray:
    """simple_knn.distCUDA2 restated by brute force (gaussian_model.py:134,546,623); `queries` = optional
    subset of point indices (for sampled checks at sizes where P*P is too much)."""
    pts = _f32(points).reshape(-1, 3)
    P = pts.shape[0]
    if queries is None:
        out = np.zeros(P, np.float32)
        lib().gso_knn3_mean_dist2(C.c_int(P), _p(pts), _p(None), C.c_int(0), _p(out))
        return out
    q = np.ascontiguousarray(np.asarray(queries, dtype=np.int32))
    out = np.zeros(q.shape[0], np.float32)
    lib().gso_knn3_mean_dist2(C.c_int(P), _p(pts), _p(q), C.c_int(q.shape[0]), _p(out))
    return out
