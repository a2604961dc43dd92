"""The reference's OWN render() and GaussianModel, executed unchanged against this repo's drop-in package.

`/root/reference/gs-simp/gaussian_renderer/__init__.py` is imported as it lies (with `diff_gaussian_rasterization`
and `simple_knn` resolving to this repo, which is the whole point of the drop-in), a reference `GaussianModel` is
filled with a small scene, and `render()` + `loss.backward()` run exactly as train.py:86-93 drives them.  There is
no GPU in this container and the product has no CPU path, so the two native entry points are replaced by recorders
that check what reaches the C-ABI binding -- 18 / 20 positional arguments in the order rasterize_points.h declares,
shapes, dtypes, the empty-tensor convention -- and hand back tensors of the documented shapes so that the reference
code keeps going.  Everything above the native call (settings tuple, module, autograd Function, the nine-value
backward wiring, means2D's (P,3) gradient, the result dict) is therefore exercised by the reference file itself,
not by a restatement of it.  Skipped where the reference tree is absent (the GPU box); tests/test_api_gpu.py runs a
line-by-line replay there, with names taken from tests/golden/render_contract.json."""
import json
import math
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/gs-simp"

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "gaussian_renderer", "__init__.py")),
                                reason="reference tree not present")


class _TorchNoCuda:
    """`torch` as the reference module sees it, minus the device="cuda" it hard-codes at :26 (no GPU here)."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def zeros_like(x, **kw):
        kw.pop("device", None)
        return torch.zeros_like(x, **kw)


@pytest.fixture()
def reference_modules(monkeypatch):
    if ROOT not in sys.path:
        monkeypatch.syspath_prepend(ROOT)
    monkeypatch.syspath_prepend(REF)
    if "plyfile" not in sys.modules:          # imported by scene/gaussian_model.py:19 for PLY io, unused on this path
        ply = types.ModuleType("plyfile")
        ply.PlyData = ply.PlyElement = object
        monkeypatch.setitem(sys.modules, "plyfile", ply)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("scene", "utils", "gaussian_renderer", "arguments")}
    import importlib
    gm = importlib.import_module("scene.gaussian_model")
    gr = importlib.import_module("gaussian_renderer")
    gr.torch = _TorchNoCuda()
    yield gr, gm
    for k in list(sys.modules):
        if k.split(".")[0] in ("scene", "utils", "gaussian_renderer", "arguments"):
            del sys.modules[k]
    sys.modules.update(saved)


def test_reference_render_runs_unchanged_on_the_dropin(reference_modules, monkeypatch):
    gr, gm = reference_modules
    from multiview_inpaint_b200 import rasterizer as R, scenes as S
    import diff_gaussian_rasterization as dgr
    assert gr.GaussianRasterizer is dgr.GaussianRasterizer and gr.GaussianRasterizationSettings is dgr.GaussianRasterizationSettings
    contract = json.load(open(os.path.join(ROOT, "tests", "golden", "render_contract.json")))

    P, W, H, deg = 500, 64, 48, 2
    sc = S.make_scene(P, W, H, deg, 11, mu_s=S.default_mu_s(W, 6.0))
    M = (deg + 1) ** 2
    pc = gm.GaussianModel(deg)                      # the reference's class, its own getters / activations
    pc.active_sh_degree = 1
    pc._xyz = torch.nn.Parameter(sc["means3D"].clone())
    pc._features_dc = torch.nn.Parameter(sc["shs"][:, :1].clone())
    pc._features_rest = torch.nn.Parameter(sc["shs"][:, 1:].clone())
    pc._scaling = torch.nn.Parameter(torch.log(sc["scales"]))
    pc._rotation = torch.nn.Parameter(sc["rotations"].clone())
    op = sc["opacities"].clamp(1e-4, 1 - 1e-4)
    pc._opacity = torch.nn.Parameter(torch.log(op / (1 - op)).reshape(P, 1))
    cam = sc["camera"]
    pipe = types.SimpleNamespace(compute_cov3D_python=False, convert_SHs_python=False, debug=False)
    bg = torch.rand(3)

    seen = {}

    def fake_forward(*args, flags=0, **kw):
        seen["fwd"] = args
        assert len(args) == 18
        (bg_, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D, viewmatrix, projmatrix, tanfovx, tanfovy,
         image_height, image_width, sh, degree, campos, prefiltered) = args
        assert bg_.shape == (3,) and means3D.shape == (P, 3) and opacity.shape == (P, 1)
        assert scales.shape == (P, 3) and rotations.shape == (P, 4) and sh.shape == (P, M, 3)
        assert colors.numel() == 0 and cov3D.numel() == 0            # "not provided" = empty tensors
        assert viewmatrix.shape == (4, 4) and projmatrix.shape == (4, 4) and campos.shape == (3,)
        assert (image_height, image_width, degree, prefiltered, scale_modifier) == (H, W, 1, False, 1.0)
        assert math.isclose(tanfovx, math.tan(cam.FoVx * 0.5)) and math.isclose(tanfovy, math.tan(cam.FoVy * 0.5))
        # post-activation values, as the reference's getters produce them
        torch.testing.assert_close(scales, sc["scales"], rtol=1e-6, atol=0)
        torch.testing.assert_close(rotations, torch.nn.functional.normalize(sc["rotations"]))
        torch.testing.assert_close(opacity.reshape(-1), op.reshape(-1), rtol=1e-5, atol=1e-7)
        color = torch.rand(3, H, W)
        radii = torch.arange(P, dtype=torch.int32) % 3            # some zero (culled), some positive
        buf = lambda n: torch.zeros(n, dtype=torch.uint8)
        return 1234, color, radii, buf(64), buf(32), buf(16), torch.full((1, H, W), 15.0)

    def fake_backward(*args, flags=0, **kw):
        seen["bwd"] = args
        assert len(args) == 20
        (bg_, means3D, radii, colors, scales, rotations, scale_modifier, cov3D, viewmatrix, projmatrix, tanfovx, tanfovy,
         dL_dout_color, sh, degree, campos, geom, num_rendered, binning, img) = args
        assert dL_dout_color.shape == (3, H, W) and radii.dtype == torch.int32 and num_rendered == 1234
        assert geom.numel() == 64 and binning.numel() == 32 and img.numel() == 16     # the buffers saved on ctx
        g2d = torch.zeros(P, 3)
        g2d[:, :2] = 1.0
        return (g2d, torch.zeros(0), torch.ones(P, 1), torch.ones(P, 3), torch.zeros(0), torch.ones(P, M, 3),
                torch.ones(P, 3), torch.ones(P, 4))

    monkeypatch.setattr(R._C, "rasterize_gaussians", fake_forward)
    monkeypatch.setattr(R._C, "rasterize_gaussians_backward", fake_backward)

    pkg = gr.render(cam, pc, pipe, bg)                                 # the reference's function, unchanged
    assert list(pkg) == contract["result_keys"]
    assert pkg["render"].shape == (3, H, W) and pkg["depth"].shape == (1, H, W)
    assert pkg["radii"].dtype == torch.int32 and pkg["visibility_filter"].dtype == torch.bool
    assert torch.equal(pkg["visibility_filter"], pkg["radii"] > 0)
    assert pkg["render"].requires_grad and not pkg["depth"].requires_grad
    gt = torch.rand(3, H, W)
    loss = (pkg["render"] - gt).abs().mean()                           # train.py:90
    loss.backward()                                                    # train.py:93
    vsp = pkg["viewspace_points"]
    assert vsp.grad is not None and vsp.grad.shape == (P, 3)           # gaussian_model.py:483 reads .grad[:, :2]
    for p in (pc._xyz, pc._features_dc, pc._features_rest, pc._scaling, pc._rotation, pc._opacity):
        assert p.grad is not None and p.grad.shape == p.shape and torch.isfinite(p.grad).all()
    # the reference's densification bookkeeping on these outputs (gaussian_model.py:482-484)
    pc.xyz_gradient_accum = torch.zeros(P, 1)
    pc.denom = torch.zeros(P, 1)
    pc.add_densification_stats(vsp, pkg["visibility_filter"])
    assert pc.denom.sum() == pkg["visibility_filter"].sum()

    # render.py:42 / render_depth.py:45: under no_grad, and with override_color (-> colors_precomp, sh empty)
    def fake_forward_override(*args, flags=0, **kw):
        assert args[2].shape == (P, 3) and args[14].numel() == 0
        return fake_forward(args[0], args[1], torch.zeros(0), *args[3:14], torch.zeros(P, M, 3), *args[15:])
    monkeypatch.setattr(R._C, "rasterize_gaussians", fake_forward_override)
    with torch.no_grad():
        pkg2 = gr.render(cam, pc, pipe, bg, override_color=torch.rand(P, 3))
    assert not pkg2["render"].requires_grad


def test_reference_render_without_recorders_reaches_the_cuda_check(reference_modules):
    """Without the recorders the same call must fail LOUDLY at the device check of the binding (there is no CPU
    fallback), i.e. every layer above it accepted what the reference passes."""
    gr, gm = reference_modules
    from multiview_inpaint_b200 import scenes as S
    sc = S.make_scene(64, 32, 32, 0, 3, mu_s=S.default_mu_s(32, 6.0))
    pc = gm.GaussianModel(0)
    pc._xyz = torch.nn.Parameter(sc["means3D"].clone())
    pc._features_dc = torch.nn.Parameter(sc["shs"][:, :1].clone())
    pc._features_rest = torch.nn.Parameter(sc["shs"][:, 1:].clone())
    pc._scaling = torch.nn.Parameter(torch.log(sc["scales"]))
    pc._rotation = torch.nn.Parameter(sc["rotations"].clone())
    pc._opacity = torch.nn.Parameter(torch.zeros(64, 1))
    pipe = types.SimpleNamespace(compute_cov3D_python=False, convert_SHs_python=False, debug=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        gr.render(sc["camera"], pc, pipe, torch.zeros(3))
