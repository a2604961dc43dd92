"""point_cloud.ply in and out of the parameter arena (multiview_inpaint_b200/plyio.py) against the layout the
reference writes (gs-simp/scene/gaussian_model.py:177-210) and reads (:268-312).  plyfile is not installed here, so
the checker is the format itself: the property list of construct_list_of_attributes, channel-major SH coefficients,
raw (pre-activation) values, binary little-endian floats."""
import struct

import numpy as np
import pytest
import torch

from multiview_inpaint_b200 import plyio
from multiview_inpaint_b200.trainstep import GaussianParamArena


def _arena(P, M, seed=0):
    g = torch.Generator().manual_seed(seed)
    pa = GaussianParamArena(P, M, "cpu")
    pa.param.normal_(generator=g)
    return pa


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_attribute_list_is_the_reference_one(deg):
    M = (deg + 1) ** 2
    names = plyio.attribute_names(M)
    assert names[:6] == ["x", "y", "z", "nx", "ny", "nz"] and names[6:9] == ["f_dc_0", "f_dc_1", "f_dc_2"]
    assert names[9:9 + 3 * (M - 1)] == [f"f_rest_{i}" for i in range(3 * M - 3)]
    assert names[-8:] == ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert len(names) == 6 + 3 * M + 8                      # 62 floats per Gaussian at degree 3


@pytest.mark.parametrize("P,deg", [(1, 0), (17, 1), (250, 3), (0, 2)])
def test_round_trip_is_bit_exact(tmp_path, P, deg):
    M = (deg + 1) ** 2
    pa = _arena(P, M, seed=P + deg)
    path = str(tmp_path / "point_cloud" / "iteration_7" / "point_cloud.ply")     # scene.save layout, scene/__init__.py
    plyio.save_ply(path, pa)
    back = plyio.load_ply(path, "cpu", sh_degree=deg)
    assert back.P == P and back.M == M and back.active_sh_degree == deg
    for name in ("_xyz", "_features", "_opacity", "_scaling", "_rotation"):
        assert torch.equal(getattr(back, name), getattr(pa, name)), name
    assert float(back.exp_avg.abs().max()) == 0.0 if P else True         # a loaded model starts with fresh optimizer state


def test_file_layout_byte_for_byte(tmp_path):
    """A 2-Gaussian degree-1 model written out and parsed by hand: header text, row stride, channel-major SH."""
    pa = GaussianParamArena(2, 4, "cpu")
    pa._xyz.copy_(torch.tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]))
    f = torch.arange(2 * 4 * 3, dtype=torch.float32).view(2, 4, 3)        # f[p, k, c] = 12 p + 3 k + c
    pa._features.copy_(f)
    pa._opacity.copy_(torch.tensor([[0.25], [-0.5]]))
    pa._scaling.copy_(torch.tensor([[-1.0, -2.0, -3.0], [-4.0, -5.0, -6.0]]))
    pa._rotation.copy_(torch.tensor([[1.0, 0.0, 0.0, 0.0], [0.5, 0.5, 0.5, 0.5]]))
    path = str(tmp_path / "m.ply")
    plyio.save_ply(path, pa)
    blob = open(path, "rb").read()
    head, body = blob.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 2"]
    assert lines[3:] == [f"property float {a}" for a in plyio.attribute_names(4)]
    n = len(plyio.attribute_names(4))
    assert n == 26 and len(body) == 2 * n * 4
    row0 = struct.unpack("<26f", body[:104])
    row1 = struct.unpack("<26f", body[104:])
    assert row0[:3] == (1.0, 2.0, 3.0) and row0[3:6] == (0.0, 0.0, 0.0)          # normals are zeros (:196)
    assert row0[6:9] == (0.0, 1.0, 2.0)                                           # f_dc_c = f[0, 0, c]
    # f_rest_{c * 3 + k} = f[0, k + 1, c]: channel-major (transpose(1, 2).flatten, :198)
    assert row0[9:18] == (3.0, 6.0, 9.0, 4.0, 7.0, 10.0, 5.0, 8.0, 11.0)
    assert row0[18] == 0.25 and row0[19:22] == (-1.0, -2.0, -3.0) and row0[22:26] == (1.0, 0.0, 0.0, 0.0)
    assert row1[6:9] == (12.0, 13.0, 14.0) and row1[9:12] == (15.0, 18.0, 21.0) and row1[18] == -0.5


def test_reads_what_plyfile_style_writers_emit(tmp_path):
    """Comments, float32 spelling, shuffled property order, a trailing face element: all legal for the reference's
    reader (it looks properties up by name and sorts f_rest_* / scale_* / rot_* by index, :281-300)."""
    names = plyio.attribute_names(4)
    rng = np.random.default_rng(3)
    table = rng.normal(size=(5, len(names))).astype("<f4")
    order = list(rng.permutation(len(names)))
    head = ["ply", "format binary_little_endian 1.0", "comment written by a test", "element vertex 5"]
    head += [f"property float32 {names[j]}" for j in order] + ["element face 0", "property list uchar int vertex_indices", "end_header"]
    path = str(tmp_path / "shuffled.ply")
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode())
        f.write(np.ascontiguousarray(table[:, order]).tobytes())
    pa = plyio.load_ply(path, "cpu")
    col = {n: table[:, k] for k, n in enumerate(names)}
    assert pa.M == 4 and torch.equal(pa._xyz, torch.from_numpy(np.stack([col["x"], col["y"], col["z"]], 1)))
    for c in range(3):
        assert torch.equal(pa._features[:, 0, c], torch.from_numpy(col[f"f_dc_{c}"]))
        for k in range(3):
            assert torch.equal(pa._features[:, k + 1, c], torch.from_numpy(col[f"f_rest_{c * 3 + k}"]))
    assert torch.equal(pa._rotation[:, 2], torch.from_numpy(col["rot_2"]))
    assert torch.equal(pa._opacity[:, 0], torch.from_numpy(col["opacity"]))


def test_rejects_what_the_reference_would_not_load(tmp_path):
    pa = _arena(3, 4)
    path = str(tmp_path / "a.ply")
    plyio.save_ply(path, pa)
    with pytest.raises(ValueError):
        plyio.load_ply(path, "cpu", sh_degree=3)                      # the assert at :282
    blob = open(path, "rb").read()
    open(path, "wb").write(blob[:-5])
    with pytest.raises(ValueError):
        plyio.load_ply(path, "cpu")                                   # truncated
    open(path, "wb").write(blob.replace(b"binary_little_endian", b"ascii"))
    with pytest.raises(ValueError):
        plyio.load_ply(path, "cpu")
    open(path, "wb").write(b"not a ply\n")
    with pytest.raises(ValueError):
        plyio.load_ply(path, "cpu")
