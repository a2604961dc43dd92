"""Drop-in contract: replays the body of the reference's render()
(gs-simp/gaussian_renderer/__init__.py:18-101) against `diff_gaussian_rasterization` with stub
`pc` / `viewpoint_camera` / `pipe` objects shaped like GaussianModel (scene/gaussian_model.py:95-118)
and Camera (scene/cameras.py:54-64), then the consumers: loss.backward() (train.py:93),
densification stats (gaussian_model.py:482-484), radii max (train.py:115), depth mask (gen_seq.py:50)."""
import math

import numpy as np
import pytest
import torch

from multiview_inpaint_b200 import scenes as S
from tests.util import oracle_forward, rel_err, small_scene

pytestmark = pytest.mark.gpu


class StubGaussians:
    """Pre-activation parameters as nn.Parameters + the getters render() reads."""

    def __init__(self, sc, max_sh_degree):
        dev = "cuda"
        self.max_sh_degree = max_sh_degree
        self.active_sh_degree = sc["sh_degree"]
        self._xyz = torch.nn.Parameter(sc["means3D"].to(dev))
        self._features_dc = torch.nn.Parameter(sc["shs"][:, :1].contiguous().to(dev))
        self._features_rest = torch.nn.Parameter(sc["shs"][:, 1:].contiguous().to(dev))
        self._scaling = torch.nn.Parameter(torch.log(sc["scales"]).to(dev))
        self._rotation = torch.nn.Parameter((sc["rotations"] * 1.7).to(dev))       # un-normalised
        op = sc["opacities"].clamp(1e-4, 1 - 1e-4)
        self._opacity = torch.nn.Parameter(torch.log(op / (1 - op)).to(dev))

    get_xyz = property(lambda s: s._xyz)
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s._rotation))
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    get_features = property(lambda s: torch.cat((s._features_dc, s._features_rest), dim=1))


class Pipe:
    compute_cov3D_python = False
    convert_SHs_python = False
    debug = False


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None):
    """Statement-for-statement replay of the reference render() call sequence (lines cited)."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer   # :14
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device="cuda") + 0  # :26
    try:
        screenspace_points.retain_grad()                                                          # :28
    except Exception:
        pass
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)                                                # :33
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(                                               # :36-49
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform, projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)                               # :51
    means3D, means2D, opacity = pc.get_xyz, screenspace_points, pc.get_opacity                     # :53-55
    scales, rotations, cov3D_precomp = pc.get_scaling, pc.get_rotation, None                       # :65-66
    shs, colors_precomp = (pc.get_features, None) if override_color is None else (None, override_color)  # :70-82
    rendered_image, radii, depth = rasterizer(                                                     # :85-93
        means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp, opacities=opacity,
        scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)
    return {"render": rendered_image, "depth": depth, "viewspace_points": screenspace_points,       # :97-101
            "visibility_filter": radii > 0, "radii": radii}


def test_render_replay_train_step_and_consumers(oracle):
    sc = small_scene(4000, 128, 96, 1, 41, 6.0, max_sh_degree=3)       # active degree 1, stride M = 16
    pc = StubGaussians(sc, 3)
    cam = sc["camera"].to("cuda")
    assert not cam.camera_center.is_contiguous() or cam.camera_center.storage_offset() >= 0
    bg = torch.rand(3, device="cuda")                                   # train.py:84 random background
    pkg = render(cam, pc, Pipe(), bg)
    # names extracted from the reference file by tests/golden/make_contract_golden.py (ast walk), not typed in here
    import inspect, json, os
    contract = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render_contract.json")))
    assert list(pkg) == contract["result_keys"]
    src = inspect.getsource(render)
    for kw in contract["settings_keywords"] + contract["call_keywords"] + contract["rasterizer_ctor_keywords"]:
        assert kw + "=" in src, kw     # the replay passes every keyword the reference passes
    image, depth, radii = pkg["render"], pkg["depth"], pkg["radii"]
    assert image.shape == (3, 96, 128) and depth.shape == (1, 96, 128) and radii.shape == (4000,)
    assert radii.dtype == torch.int32 and pkg["visibility_filter"].dtype == torch.bool
    assert image.requires_grad and not depth.requires_grad and not radii.requires_grad
    gt = torch.rand(3, 96, 128, device="cuda")
    loss = (image - gt).abs().mean()                                    # train.py:90
    loss.backward()                                                     # train.py:93
    vsp = pkg["viewspace_points"]
    assert vsp.grad is not None and vsp.grad.shape == (4000, 3)
    vis = pkg["visibility_filter"]
    # gaussian_model.py:482-484
    accum = torch.zeros(4000, 1, device="cuda")
    accum[vis] += torch.norm(vsp.grad[vis, :2], dim=-1, keepdim=True)
    assert accum.sum() > 0 and (vsp.grad[:, 2] == 0).all() and (vsp.grad[~vis] == 0).all()
    # train.py:115 (float32 buffer, int32 radii -> type promotion)
    max_radii2D = torch.zeros(4000, device="cuda")
    max_radii2D[vis] = torch.max(max_radii2D[vis], radii[vis])
    assert max_radii2D.max() == radii.max()
    for p in (pc._xyz, pc._features_dc, pc._features_rest, pc._scaling, pc._rotation, pc._opacity):
        assert p.grad is not None and torch.isfinite(p.grad).all()
    assert (pc._features_rest.grad[:, 3:] == 0).all()                   # coefficients above the active degree
    # gen_seq.py:50 depth mask arithmetic
    inter_t = torch.full((1, 96, 128), 5.0, device="cuda")
    mask = (inter_t > 0.) & ((inter_t < depth) | (depth == 15.))
    assert mask.dtype == torch.bool and (depth <= 15.).all() and (depth > 0.2).all()
    # against the oracle, through the same activations
    sc2 = dict(sc)
    sc2["rotations"] = pc.get_rotation.detach().cpu()      # exactly what the kernels received
    sc2["scales"] = pc.get_scaling.detach().cpu()
    sc2["opacities"] = pc.get_opacity.detach().cpu()
    f = oracle_forward(oracle, sc2, bg=bg.cpu().numpy())
    np.testing.assert_array_equal(radii.cpu().numpy(), f.radii)
    assert np.abs(image.detach().cpu().numpy() - f.color).max() < 1e-5
    dL = (torch.sign(image.detach() - gt) / image.numel()).cpu().numpy()
    g = oracle.backward(f, dL)
    assert rel_err(vsp.grad.cpu().numpy(), g["dL_dmeans2D"]) < 1e-3
    assert rel_err(pc._xyz.grad.cpu().numpy(), g["dL_dmeans3D"]) < 1e-3


def test_no_grad_inference_and_override_color():
    """render.py:42 / render_depth.py:45 run under no_grad; override_color feeds colors_precomp."""
    sc = small_scene(3000, 96, 64, 0, 43, 6.0)
    pc = StubGaussians(sc, 0)
    cam = sc["camera"].to("cuda")
    with torch.no_grad():
        a = render(cam, pc, Pipe(), torch.zeros(3, device="cuda"))
        b = render(cam, pc, Pipe(), torch.zeros(3, device="cuda"), override_color=torch.rand(3000, 3, device="cuda"))
    assert not a["render"].requires_grad and a["render"].shape == b["render"].shape
    assert torch.equal(a["radii"], b["radii"]) and torch.equal(a["depth"], b["depth"])
    disparity = 1. / torch.clamp_min(a["depth"], 0.001)                 # render_depth.py:37
    assert torch.isfinite(disparity).all()


def test_forward_is_deterministic_and_stateless():
    sc = small_scene(5000, 128, 80, 2, 45, 6.0)
    pc = StubGaussians(sc, 2)
    cam = sc["camera"].to("cuda")
    with torch.no_grad():
        r1 = render(cam, pc, Pipe(), torch.ones(3, device="cuda"))
        other = small_scene(700, 64, 48, 0, 46)
        render(other["camera"].to("cuda"), StubGaussians(other, 0), Pipe(), torch.zeros(3, device="cuda"))  # P changes between calls
        r2 = render(cam, pc, Pipe(), torch.ones(3, device="cuda"))
    assert torch.equal(r1["render"], r2["render"]) and torch.equal(r1["depth"], r2["depth"])


def test_empty_point_cloud_and_errors():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    cam = S.make_camera(64, 48).to("cuda")
    rs = GaussianRasterizationSettings(image_height=48, image_width=64, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                       bg=torch.ones(3, device="cuda"), scale_modifier=1.0,
                                       viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
                                       sh_degree=0, campos=cam.camera_center, prefiltered=False)
    r = GaussianRasterizer(raster_settings=rs)
    z = torch.zeros(0, 3, device="cuda")
    color, radii, depth = r(means3D=z, means2D=z, opacities=torch.zeros(0, 1, device="cuda"), shs=torch.zeros(0, 1, 3, device="cuda"),
                            scales=z, rotations=torch.zeros(0, 4, device="cuda"))
    assert color.shape == (3, 48, 64) and (color == 0).all() and radii.numel() == 0 and depth.shape == (1, 48, 64)
    with pytest.raises(RuntimeError, match="num_points, 3"):
        r(means3D=torch.zeros(5, 4, device="cuda"), means2D=z, opacities=torch.zeros(5, 1, device="cuda"),
          shs=torch.zeros(5, 1, 3, device="cuda"), scales=torch.zeros(5, 3, device="cuda"), rotations=torch.zeros(5, 4, device="cuda"))
    # prefiltered=True with a culled point: the reference traps; here a clean error
    rs2 = rs._replace(prefiltered=True)
    pts = torch.tensor([[0.0, 0.0, -1.0]], device="cuda")
    with pytest.raises(RuntimeError, match="prefiltered"):
        GaussianRasterizer(raster_settings=rs2)(means3D=pts, means2D=pts, opacities=torch.ones(1, 1, device="cuda"),
                                                shs=torch.zeros(1, 1, 3, device="cuda"), scales=torch.ones(1, 3, device="cuda"),
                                                rotations=torch.tensor([[1.0, 0, 0, 0]], device="cuda"))
    vis = r.markVisible(torch.tensor([[0.0, 0.0, -1.0], [0.0, 0.0, 2.0]], device="cuda"))
    assert vis.tolist() == [False, True]


def test_runs_on_non_default_stream():
    sc = small_scene(3000, 96, 64, 1, 47, 6.0)
    pc = StubGaussians(sc, 1)
    cam = sc["camera"].to("cuda")
    with torch.no_grad():
        ref = render(cam, pc, Pipe(), torch.zeros(3, device="cuda"))["render"].clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        got = render(cam, pc, Pipe(), torch.zeros(3, device="cuda"))["render"]
    s.synchronize()
    assert torch.equal(ref, got)


def test_batched_render_views_matches_per_view_render(oracle):
    """cuda_views_render (the render.py:32-39 / render_depth.py:31-39 / gen_seq.py:36-58 loops as
    ONE batched call: two streams, per-view workspaces, asynchronous N) gives bit-identical colour / depth / radii
    to one render() per camera, its `sink` sees every view on the view's stream, and the oracle agrees."""
    from multiview_inpaint_b200 import _C, multiview as mv
    from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings
    sc = small_scene(6000, 160, 96, 1, 51, 6.0)
    pc = StubGaussians(sc, 1)
    cams = S.orbit_cameras(5, 160, 96, max_deg=20.0)
    bg = torch.zeros(3, device="cuda")
    with torch.no_grad():
        want = [render(c.to("cuda"), pc, Pipe(), bg) for c in cams]
        gauss = dict(means3D=pc.get_xyz, shs=pc.get_features, opacities=pc.get_opacity, scales=pc.get_scaling,
                     rotations=pc.get_rotation)
    cd = [c.to("cuda") for c in cams]
    rs = [GaussianRasterizationSettings(image_height=96, image_width=160, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                                        scale_modifier=1.0, viewmatrix=c.world_view_transform, projmatrix=c.full_proj_transform,
                                        sh_degree=1, campos=c.camera_center, prefiltered=False) for c in cd]
    # synchronous form (exact sizes, no workspaces)
    plain = mv.cuda_views_render(gauss, rs)
    for w, g in zip(want, plain):
        assert torch.equal(w["render"], g.color) and torch.equal(w["depth"], g.depth) and torch.equal(w["radii"], g.radii)
        assert g.num_rendered > 0
    # asynchronous pipelined form: sinks copy to pinned host memory on the view's stream
    av = mv.AsyncViews(len(rs))
    for v, g in enumerate(plain):
        av.learn(v, g.num_rendered)
    pipe = mv.ViewPipeline(torch.device("cuda"), depth=2)
    ws = [_C.Workspace(torch.device("cuda")) for _ in rs]
    host_c = [torch.empty(3, 96, 160).pin_memory() for _ in rs]
    host_d = [torch.empty(1, 96, 160).pin_memory() for _ in rs]
    seen = []

    def sink(k, color, depth, radii):
        seen.append(k)
        host_c[k].copy_(color, non_blocking=True)
        host_d[k].copy_(depth, non_blocking=True)
    for _ in range(2):   # second pass reuses the workspaces
        seen.clear()
        res = mv.cuda_views_render(gauss, rs, capacities=[av.capacity(v) for v in range(len(rs))],
                                   async_results=[av.slot(v) for v in range(len(rs))], pipeline=pipe, workspaces=ws, sink=sink)
        torch.cuda.synchronize()
        assert not av.check(range(len(rs))) and seen == list(range(len(rs)))
        for v, (w, g) in enumerate(zip(want, res)):
            assert g.num_rendered == -1 and int(av.slot(v)[0]) == plain[v].num_rendered
            assert torch.equal(w["render"].cpu(), host_c[v]) and torch.equal(w["depth"].cpu(), host_d[v])
            assert torch.equal(w["radii"], g.radii)
    # and against the oracle for one rotated camera, through the same activations
    sc2 = dict(sc)
    sc2["rotations"], sc2["scales"], sc2["opacities"] = gauss["rotations"].cpu(), gauss["scales"].cpu(), gauss["opacities"].cpu()
    f = oracle_forward(oracle, sc2, cam=cams[3])
    assert np.abs(host_c[3].numpy() - f.color).max() < 1e-5
    assert np.array_equal(want[3]["radii"].cpu().numpy(), f.radii)
