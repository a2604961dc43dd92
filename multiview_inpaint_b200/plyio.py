"""`point_cloud.ply` in and out of the parameter arena -- the data format on either side of the training path.

The reference stores a trained model as a binary little-endian PLY written by `GaussianModel.save_ply`
(gs-simp/scene/gaussian_model.py:192-210, called through `scene.save(iteration)`, train.py:108) and every render
script starts from `load_ply` (:268-312; render.py / render_depth.py / gen_seq.py via `Scene(..., load_iteration)`).
A model optimised with the fused step (trainstep.GaussianParamArena) must therefore leave a file the reference's
scripts load unchanged, and the arena must start from the reference's checkpoints.

Format (`construct_list_of_attributes`, :177-190): one `vertex` element, all properties `float`, in this order
    x y z | nx ny nz (zeros) | f_dc_0..2 | f_rest_0..3(M-1)-1 | opacity | scale_0..2 | rot_0..3
with the SH coefficients CHANNEL-major -- `_features_rest (P, M-1, 3)` is written as `transpose(1, 2).flatten(1)`
(:198), i.e. f_rest_{c (M-1) + k} = coefficient k+1 of colour channel c -- and all values raw (pre-activation).
The reference writes through the third-party `plyfile` package (not installed here, unpinned); the header it emits
for an all-`f4` structured array is restated below and `load_ply` accepts what `plyfile` accepts for this layout
(comments, `float` / `float32` spellings, properties in any order).  numpy only; no GPU work.
"""
from __future__ import annotations

import os

import numpy as np
import torch


def attribute_names(M: int) -> list[str]:
    """`construct_list_of_attributes` (gaussian_model.py:177-190) for (M-1) non-DC coefficients per channel."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(3)]
    names += [f"f_rest_{i}" for i in range(3 * (M - 1))]
    names += ["opacity"]
    names += [f"scale_{i}" for i in range(3)]
    names += [f"rot_{i}" for i in range(4)]
    return names


def _header(n: int, names: list[str]) -> bytes:
    lines = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"]
    lines += [f"property float {a}" for a in names]
    lines += ["end_header"]
    return ("\n".join(lines) + "\n").encode("ascii")


def save_ply(path: str, params) -> None:
    """`GaussianModel.save_ply(path)` for a trainstep.GaussianParamArena (any device)."""
    P, M = params.P, params.M
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)                                   # mkdir_p(os.path.dirname(path)), :193
    cpu = lambda t: t.detach().to("cpu", torch.float32).contiguous().numpy()
    feats = cpu(params._features)                                       # (P, M, 3)
    cols = [cpu(params._xyz), np.zeros((P, 3), np.float32),
            feats[:, :1, :].transpose(0, 2, 1).reshape(P, 3),           # f_dc: transpose(1, 2).flatten(1), :197
            feats[:, 1:, :].transpose(0, 2, 1).reshape(P, 3 * (M - 1)),  # f_rest: channel-major, :198
            cpu(params._opacity).reshape(P, 1), cpu(params._scaling), cpu(params._rotation)]
    table = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype="<f4")
    names = attribute_names(M)
    assert table.shape == (P, len(names))
    with open(path, "wb") as f:
        f.write(_header(P, names))
        f.write(table.tobytes())


_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
          "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
          "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


def read_vertex_table(path: str) -> dict:
    """The `vertex` element of a binary little-endian PLY as {property: array} (what
    `np.asarray(plydata.elements[0][name])` gives the reference's load_ply)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, n, props, in_vertex, seen_vertex = None, 0, [], False, False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: header without end_header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if seen_vertex:
                    in_vertex = False                     # later elements (faces ...) follow the vertex block: ignored
                elif tok[1] != "vertex":
                    raise ValueError(f"{path}: element {tok[1]!r} precedes 'vertex' (not a file the reference writes)")
                else:
                    n, seen_vertex, in_vertex = int(tok[2]), True, True
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list property in the vertex element")
                if tok[1] not in _TYPES:
                    raise ValueError(f"{path}: unknown property type {tok[1]}")
                props.append((tok[2], _TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt != "binary_little_endian":
            raise ValueError(f"{path}: format {fmt!r}; the reference writes binary_little_endian")
        if not seen_vertex:
            raise ValueError(f"{path}: no vertex element")
        dt = np.dtype(props)
        raw = f.read(n * dt.itemsize)
        if len(raw) != n * dt.itemsize:
            raise ValueError(f"{path}: truncated ({len(raw)} of {n * dt.itemsize} vertex bytes)")
    rec = np.frombuffer(raw, dtype=dt, count=n)
    return {name: np.asarray(rec[name]) for name, _ in props}


def load_ply(path: str, device, sh_degree: int | None = None):
    """`GaussianModel.load_ply(path)` (:268-312) into a trainstep.GaussianParamArena on `device`.  `sh_degree`
    (the model's max_sh_degree) is checked like the reference's assert (:282) when given, else inferred from the
    number of f_rest_* properties."""
    from .trainstep import GaussianParamArena
    v = read_vertex_table(path)
    need = ("x", "y", "z", "opacity", "f_dc_0", "f_dc_1", "f_dc_2")
    missing = [k for k in need if k not in v]
    if missing:
        raise ValueError(f"{path}: missing properties {missing}")
    P = v["x"].shape[0]
    by_index = lambda prefix: sorted((k for k in v if k.startswith(prefix)), key=lambda s: int(s.split("_")[-1]))   # :281, :290, :296
    rest, scales, rots = by_index("f_rest_"), by_index("scale_"), by_index("rot")
    if len(rest) % 3:
        raise ValueError(f"{path}: {len(rest)} f_rest_* properties (not a multiple of 3)")
    M = 1 + len(rest) // 3
    deg = int(round(M ** 0.5)) - 1
    if (deg + 1) ** 2 != M:
        raise ValueError(f"{path}: {len(rest)} f_rest_* properties do not form a full SH degree")
    if sh_degree is not None and len(rest) != 3 * (sh_degree + 1) ** 2 - 3:
        raise ValueError(f"{path}: {len(rest)} f_rest_* properties, expected {3 * (sh_degree + 1) ** 2 - 3} for SH degree {sh_degree}")
    if len(scales) != 3 or len(rots) != 4:
        raise ValueError(f"{path}: expected scale_0..2 and rot_0..3")
    f32 = lambda a: torch.from_numpy(np.array(a, dtype=np.float32, order="C"))       # a writable copy
    xyz = np.stack([v["x"], v["y"], v["z"]], axis=1)
    f_dc = np.stack([v["f_dc_0"], v["f_dc_1"], v["f_dc_2"]], axis=1).reshape(P, 3, 1)
    f_rest = np.stack([v[k] for k in rest], axis=1).reshape(P, 3, M - 1) if M > 1 else np.zeros((P, 3, 0))
    # (P, 3, coeffs) channel-major in the file -> (P, coeffs, 3) in the model: .transpose(1, 2), :304-305
    params = GaussianParamArena.from_tensors(
        f32(xyz).to(device), f32(f_dc.transpose(0, 2, 1)).to(device), f32(f_rest.transpose(0, 2, 1)).to(device),
        f32(v["opacity"].reshape(P, 1)).to(device), f32(np.stack([v[k] for k in scales], axis=1)).to(device),
        f32(np.stack([v[k] for k in rots], axis=1)).to(device))
    params.active_sh_degree = deg                                       # :312 active_sh_degree = max_sh_degree
    return params
