"""GPU parity of the image-output stage of the inference loops: gsr_quantize_rgb8 is bit-exact
against the numpy oracle of torchvision.utils.save_image's quantisation, and AsyncImageWriter puts exactly those
pixels (PNG) / exactly those values (NPY) on disk while the caller keeps queueing views."""
import os

import numpy as np
import pytest
import torch

from oracle import trainstep_oracle as T

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def C():
    from multiview_inpaint_b200 import _C
    return _C


@pytest.mark.parametrize("shape", [(3, 1, 1), (3, 5, 7), (1, 9, 13), (3, 64, 48), (3, 1080, 1920), (1, 576, 1024), (3, 33, 3)])
def test_quantize_bit_exact(C, shape):
    torch.manual_seed(sum(shape))
    x = torch.rand(*shape) * 1.5 - 0.25                    # values below 0 and above 1 are clamped
    flat = x.view(-1)
    edge = torch.tensor([0.0, 1.0, 0.5 / 255, 1.5 / 255, 254.5 / 255, 255.5 / 255, -0.0, 2.0, -3.0, float("nan"),
                         float("inf"), -float("inf"), 127.5 / 255, 0.49999997 / 255])
    n = min(edge.numel(), flat.numel())
    flat[:n] = edge[:n]
    got = C.quantize_rgb8(x.to(DEV)).cpu().numpy()
    ref = T.save_image_u8(x.numpy())
    assert got.shape == ref.shape == (shape[1], shape[2], 3) and got.dtype == np.uint8
    assert np.array_equal(got, ref)
    # and against the torch expression itself on the device (finite values)
    xf = torch.nan_to_num(x, nan=0.0).to(DEV)
    t = xf.expand(3, -1, -1).clone().mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).cpu().numpy()
    assert np.array_equal(got, t)


def test_quantize_affine_normalize_0_to_1(C):
    """render_depth.py:36-39: disparity = 1 / clamp_min(depth, 0.001); save_image(normalize_0_to_1(disparity))"""
    torch.manual_seed(3)
    depth = (torch.rand(1, 40, 56) * 14 + 0.5).to(DEV)
    disp = 1.0 / torch.clamp_min(depth, 0.001)
    lo, hi = disp.min(), disp.max()
    aff = torch.stack([lo, 1.0 / (hi - lo)])              # stays on the device: no host round trip
    got = C.quantize_rgb8(disp.contiguous(), affine=aff).cpu().numpy()
    ref = T.save_image_u8(disp.cpu().numpy(), affine=aff.cpu().numpy())
    assert np.array_equal(got, ref)
    assert got.min() == 0 and got.max() == 255


def test_quantize_rejects_bad_arguments(C):
    with pytest.raises(RuntimeError):
        C.quantize_rgb8(torch.rand(2, 4, 4, device=DEV))
    with pytest.raises(RuntimeError, match="CUDA"):
        C.quantize_rgb8(torch.rand(3, 4, 4))


@pytest.mark.parametrize("use_pil", [True, False])
def test_async_writer_png_and_npy(C, tmp_path, use_pil):
    PIL = pytest.importorskip("PIL.Image")
    from multiview_inpaint_b200.imagewriter import AsyncImageWriter
    torch.manual_seed(7)
    imgs = [torch.rand(3, 120, 200, device=DEV) for _ in range(10)] + [torch.rand(1, 77, 33, device=DEV)]
    depth = torch.rand(1, 120, 200, device=DEV) * 15
    side = torch.cuda.Stream()
    with AsyncImageWriter(DEV, slots=3, workers=2, use_pil=use_pil) as w:     # fewer slots than images: back-pressure path
        side.wait_stream(torch.cuda.current_stream())          # the images were produced on the current stream
        for k, im in enumerate(imgs):
            if k % 2:
                with torch.cuda.stream(side):
                    w.submit_png(str(tmp_path / f"{k:05d}.png"), im)
            else:
                w.submit_png(str(tmp_path / f"{k:05d}.png"), im)
        w.submit_npy(str(tmp_path / "depth.npy"), depth)
        w.flush()
        assert w.bytes_d2h == sum(3 * i.shape[1] * i.shape[2] for i in imgs) + depth.numel() * 4
    for k, im in enumerate(imgs):
        on_disk = np.asarray(PIL.open(tmp_path / f"{k:05d}.png").convert("RGB"))
        assert np.array_equal(on_disk, T.save_image_u8(im.cpu().numpy())), k
    assert np.array_equal(np.load(tmp_path / "depth.npy"), depth.cpu().numpy())


def test_async_writer_reports_failures(tmp_path):
    from multiview_inpaint_b200.imagewriter import AsyncImageWriter
    w = AsyncImageWriter(DEV, slots=2, workers=1)
    w.submit_png(str(tmp_path / "no_such_dir" / "x.png"), torch.rand(3, 8, 8, device=DEV))
    with pytest.raises(RuntimeError, match="write"):
        w.flush()
    w.close()


def test_render_views_with_writer_sink(C, tmp_path):
    """render.py:32-39 as one batched call: every view's PNG equals quantising that view's colour."""
    PIL = pytest.importorskip("PIL.Image")
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    from multiview_inpaint_b200 import multiview as mv, scenes as S
    from multiview_inpaint_b200.imagewriter import AsyncImageWriter
    from tests.util import small_scene
    sc = small_scene(P=3000, W=96, H=80, deg=1, seed=21)
    g = {k: sc[k].to(DEV) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    cams = [c.to(DEV) for c in S.orbit_cameras(5, 96, 80, max_deg=10.0)]
    bg = torch.ones(3, device=DEV)
    settings = [GaussianRasterizationSettings(image_height=80, image_width=96, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                                              scale_modifier=1.0, viewmatrix=c.world_view_transform, projmatrix=c.full_proj_transform,
                                              sh_degree=1, campos=c.camera_center, prefiltered=False) for c in cams]
    pipe = mv.ViewPipeline(DEV, depth=2)
    keep = {}
    with AsyncImageWriter(DEV, slots=4, workers=2) as w:
        def sink(k, color, depth, radii):
            keep[k] = color.clone()
            w.submit_png(str(tmp_path / f"{k:05d}.png"), color)
        mv.cuda_views_render(g, settings, pipeline=pipe, sink=sink)
    torch.cuda.synchronize()
    for k in range(5):
        on_disk = np.asarray(PIL.open(tmp_path / f"{k:05d}.png").convert("RGB"))
        assert np.array_equal(on_disk, T.save_image_u8(keep[k].cpu().numpy()))
