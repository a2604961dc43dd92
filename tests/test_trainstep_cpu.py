"""Pins oracle/trainstep_oracle.py against tests/golden/trainstep.npz -- outputs of the REFERENCE's own
utils/loss_utils.py (l1_loss, ssim, autograd), torch's activation callables that gaussian_model.py:33-41 installs,
and torch.optim.Adam over the reference's six parameter groups (tests/golden/make_trainstep_golden.py).
This oracle is therefore parity-PINNED (the rasterizer's oracle is not: its source is outside the tree)."""
import os

import numpy as np
import pytest
import torch

from oracle import trainstep_oracle as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "trainstep.npz"))


def test_window_matches_reference_create_window(gold):
    w = T.gaussian_window()
    assert w.dtype == np.float32 and w.shape == (11,)
    assert abs(float(w.sum()) - 1.0) < 1e-6 and np.all(w == w[::-1]) and w.argmax() == 5


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_loss_forward_against_reference(gold, case):
    img, gt = gold[f"loss_{case}_img"], gold[f"loss_{case}_gt"]
    lam = float(gold["loss_lambda"])
    ref = gold[f"loss_{case}_out"]
    l1, ss, loss = T.loss_forward(img, gt, lam)
    # the reference computes in float32 (conv2d accumulation order unknown): 2e-6 absolute on O(0.1..1) scalars
    assert abs(l1 - ref[0]) < 2e-6 and abs(ss - ref[1]) < 2e-6 and abs(loss - ref[2]) < 2e-6, (l1, ss, loss, ref)


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_loss_backward_against_reference_autograd(gold, case):
    img, gt = gold[f"loss_{case}_img"], gold[f"loss_{case}_gt"]
    g = T.loss_backward(img, gt, float(gold["loss_lambda"]))
    ref = gold[f"loss_{case}_grad"].astype(np.float64)
    # exactly-equal pixels: |x - y| has subgradient sign(0) = 0 in torch
    assert np.abs(g - ref).max() <= 2e-5 * np.abs(ref).max(), np.abs(g - ref).max() / np.abs(ref).max()


def test_loss_backward_is_the_gradient_of_forward(gold):
    """finite differences of the oracle's own forward, in float64"""
    img, gt = gold["loss_c_img"].astype(np.float64), gold["loss_c_gt"].astype(np.float64)
    g = T.loss_backward(img, gt, 0.35)
    rng = np.random.default_rng(3)
    for _ in range(12):
        c, y, x = rng.integers(0, 3), rng.integers(0, 7), rng.integers(0, 9)
        if img[c, y, x] == gt[c, y, x]:
            continue
        h = 1e-6
        a, b = img.copy(), img.copy()
        a[c, y, x] += h
        b[c, y, x] -= h
        fd = (T.loss_forward(a, gt, 0.35)[2] - T.loss_forward(b, gt, 0.35)[2]) / (2 * h)
        assert abs(fd - g[c, y, x]) < 1e-6 * max(1.0, abs(fd)) + 1e-9, (fd, g[c, y, x])


def test_activations_against_torch(gold):
    s, q, o = T.activate_forward(gold["act_raw_s"], gold["act_raw_q"], gold["act_raw_o"])
    np.testing.assert_allclose(s, gold["act_s"], rtol=3e-7)
    np.testing.assert_allclose(q, gold["act_q"], rtol=3e-7, atol=1e-9)
    np.testing.assert_allclose(o, gold["act_o"], rtol=3e-7)
    assert np.all(q[3] == 0.0)      # zero quaternion stays zero (norm clamped at eps)


def test_activation_backward_against_torch_autograd(gold):
    ds, dq, do = T.activate_backward(gold["act_raw_s"], gold["act_raw_q"], gold["act_raw_o"], gold["act_gs"],
                                     gold["act_gq"], gold["act_go"])
    np.testing.assert_allclose(ds, gold["act_d_raw_s"], rtol=1e-6)
    # the reference forms (1 - y) in fp32, which cancels for y -> 1 (|raw| up to ~8): relative 1e-4 there
    np.testing.assert_allclose(do, gold["act_d_raw_o"], rtol=1e-4, atol=1e-9)
    ref = gold["act_d_raw_q"].astype(np.float64)
    ok = np.ones(len(ref), bool)
    ok[3] = False                   # norm == 0 < eps: gradient is g / eps, huge; compared relatively below
    np.testing.assert_allclose(dq[ok], ref[ok], rtol=1e-4, atol=2e-6 * np.abs(ref[ok]).max())
    np.testing.assert_allclose(dq[3], ref[3], rtol=1e-6)


GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def test_adam_against_torch_optim(gold):
    steps = int(gold["adam_steps"])
    for k in GROUPS:
        p, lr = gold[f"adam_p0_{k}"], float(gold[f"adam_lr_{k}"])
        m, v = np.zeros_like(p), np.zeros_like(p)
        for t in range(steps):
            p, m, v = T.adam_step(p, gold[f"adam_g{t}_{k}"], m, v, t + 1, lr)
            # fp32 rounding order only (lerp as fma or not; gradients jump decades between steps, so m cancels)
            mr, vr = gold[f"adam_m{t + 1}_{k}"], gold[f"adam_v{t + 1}_{k}"]
            np.testing.assert_allclose(m, mr, rtol=2e-6, atol=2e-7 * np.abs(mr).max())
            np.testing.assert_allclose(v, vr, rtol=2e-6, atol=2e-7 * np.abs(vr).max())
            # parameters are O(1) and move by ~lr per step: compare the position to 1 ulp-ish of fp32
            np.testing.assert_allclose(p, gold[f"adam_p{t + 1}_{k}"], rtol=0, atol=3e-7)


def test_adam_per_element_lr_equals_two_groups(gold):
    """The fused kernel keeps f_dc and f_rest in ONE (P,M,3) tensor with a column-dependent learning rate; in the
    oracle that is adam_step with an lr array, and it must equal the reference's two separate groups."""
    M = int(gold["adam_M"])
    dc, rest = gold["adam_p0_f_dc"], gold["adam_p0_f_rest"]
    sh = np.concatenate([dc, rest], axis=1)
    lr = np.empty_like(sh)
    lr[:, :1], lr[:, 1:] = float(gold["adam_lr_f_dc"]), float(gold["adam_lr_f_rest"])
    assert sh.shape[1] == M
    m, v = np.zeros_like(sh), np.zeros_like(sh)
    for t in range(int(gold["adam_steps"])):
        g = np.concatenate([gold[f"adam_g{t}_f_dc"], gold[f"adam_g{t}_f_rest"]], axis=1)
        sh, m, v = T.adam_step(sh, g, m, v, t + 1, lr)
    np.testing.assert_allclose(sh[:, :1], gold["adam_p5_f_dc"], rtol=0, atol=3e-7)
    np.testing.assert_allclose(sh[:, 1:], gold["adam_p5_f_rest"], rtol=0, atol=3e-7)


@pytest.mark.parametrize("P,M", [(1, 1), (61, 4), (1001, 16), (7, 9)])
def test_param_arena_layout_equals_grad_arena(P, M):
    """Parameter, moment and gradient slices must sit at identical offsets (one Adam launch walks them together),
    every slice 16-byte aligned, and _features_dc / _features_rest must be views of the one (P,M,3) SH tensor."""
    import torch
    from multiview_inpaint_b200.multiview import GradArena
    from multiview_inpaint_b200.trainstep import GaussianParamArena
    pa, ga = GaussianParamArena(P, M, "cpu"), GradArena(P, M, "cpu")
    pairs = {"_xyz": "dL_dmeans3D", "_features": "dL_dsh", "_opacity": "dL_dopacity", "_scaling": "dL_dscales",
             "_rotation": "dL_drotations"}
    assert pa.n_flat == ga.flat.numel()
    for pn, gn in pairs.items():
        p, g = getattr(pa, pn), ga.views[gn]
        assert p.shape == g.shape
        assert (p.data_ptr() - pa.param.data_ptr()) == (g.data_ptr() - ga.flat.data_ptr())
        assert (p.data_ptr() - pa.param.data_ptr()) % 16 == 0
        m, v = pa.moments(pn)
        assert m.shape == p.shape and (m.data_ptr() - pa.exp_avg.data_ptr()) == (p.data_ptr() - pa.param.data_ptr())
    assert pa._features_dc.shape == (P, 1, 3) and pa._features_rest.shape == (P, M - 1, 3)
    pa._features_dc.fill_(2.0)
    assert torch.all(pa.get_features[:, 0] == 2.0)
    x = torch.arange(P * 3, dtype=torch.float32).view(P, 3)
    pb = GaussianParamArena.from_tensors(x, torch.ones(P, 1, 3), torch.zeros(P, M - 1, 3), torch.zeros(P, 1), x, torch.ones(P, 4))
    assert torch.equal(pb._xyz, x) and torch.equal(pb._scaling, x) and torch.all(pb._features[:, 0] == 1)


def test_trainstep_entry_points_refuse_cpu_tensors():
    """no CPU / torch fallback: the product path fails loudly without a CUDA device"""
    import torch
    from multiview_inpaint_b200 import _C
    x = torch.zeros(3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        _C.loss_l1_ssim_forward(x, x, 0.2)
    with pytest.raises(RuntimeError, match="CUDA"):
        _C.activate_forward(torch.zeros(4, 3), torch.zeros(4, 4), torch.zeros(4, 1))
    with pytest.raises(RuntimeError, match="CUDA"):
        _C.adam_step([dict(param=x, grad=x, exp_avg=x, exp_avg_sq=x, lr=0.1)], 1)
    assert _C.loss_temp_bytes(3, 1008, 1600) >= 3 * 3 * 1008 * 1600 * 4


def test_png_encoder_round_trips_through_pil(tmp_path):
    """The fallback PNG encoder of the async writer (used when PIL is absent) must decode to the same pixels."""
    PIL = pytest.importorskip("PIL.Image")
    from multiview_inpaint_b200.imagewriter import encode_png_rgb8
    rng = np.random.default_rng(0)
    for shape in ((1, 1, 3), (7, 13, 3), (64, 48, 3)):
        rgb = rng.integers(0, 256, shape, dtype=np.uint8)
        p = tmp_path / f"a{shape[0]}.png"
        p.write_bytes(encode_png_rgb8(rgb))
        assert np.array_equal(np.asarray(PIL.open(p).convert("RGB")), rgb)


def test_save_image_quantisation_oracle_matches_torch_expression():
    """oracle save_image_u8 == the torch expression torchvision.utils.save_image applies (mul(255).add_(0.5)
    .clamp_(0, 255).to(uint8)), evaluated with torch on the CPU, including values outside [0,1] and ties."""
    import torch
    torch.manual_seed(1)
    x = torch.rand(3, 33, 47) * 1.4 - 0.2
    x[0, 0, :6] = torch.tensor([0.0, 1.0, 0.5 / 255, 1.5 / 255, 254.5 / 255, 2.0])
    ref = x.clone().mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()
    assert np.array_equal(T.save_image_u8(x.numpy()), ref)
    g = torch.rand(1, 5, 9)
    ref1 = g.expand(3, -1, -1).clone().mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()
    assert np.array_equal(T.save_image_u8(g.numpy()), ref1)


# ------------------------------------------------------------------------------------------------ schedules (host side)
def test_position_lr_schedule_matches_reference_get_expon_lr_func():
    """tests/golden/lr_schedule.npz holds the REFERENCE's get_expon_lr_func (utils/general_utils.py:31-64) evaluated for
    its two optimisation presets and a delayed variant (make_lr_golden.py): identical doubles."""
    from multiview_inpaint_b200.trainstep import expon_lr_func, learning_rates
    g = np.load(os.path.join(ROOT, "tests", "golden", "lr_schedule.npz"))
    steps = g["steps"]
    for name in ("default_30k", "inpaint_300", "delayed", "disabled"):
        a = g[f"{name}_args"]
        f = expon_lr_func(lr_init=a[0], lr_final=a[1], lr_delay_steps=int(a[2]), lr_delay_mult=a[3], max_steps=int(a[4]))
        got = np.array([float(f(int(s))) for s in steps])
        np.testing.assert_array_equal(got, g[f"{name}_lr"], err_msg=name)
    f = expon_lr_func(0.00016, 0.0000016, 0, 0.01, 30_000)
    lrs = learning_rates(0, f)
    assert lrs == dict(xyz=0.00016, f_dc=0.0025, f_rest=0.0025 / 20.0, opacity=0.05, scaling=0.005, rotation=0.001)
    assert abs(learning_rates(30_000, f)["xyz"] - 0.0000016) < 1e-18


def test_sh_degree_schedule_on_the_arena():
    from multiview_inpaint_b200.trainstep import GaussianParamArena
    pa = GaussianParamArena(5, 16, "cpu")
    assert (pa.max_sh_degree, pa.active_sh_degree) == (3, 0)
    for want in (1, 2, 3, 3, 3):                       # gaussian_model.py:120-122: saturates at max_sh_degree
        pa.oneupSHdegree()
        assert pa.active_sh_degree == want
    assert GaussianParamArena(5, 1, "cpu").max_sh_degree == 0 and GaussianParamArena(5, 4, "cpu").max_sh_degree == 1


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_create_from_pcd_matches_reference(case):
    """tests/golden/init_from_pcd.npz holds what the REFERENCE's GaussianModel.create_from_pcd (gaussian_model.py:124-147)
    builds from a seeded point cloud (make_init_golden.py; its distCUDA2 call answered by the brute-force oracle the
    CUDA kernel is bit-identical to).  Given the same neighbour distances the arena must hold the same six tensors."""
    from multiview_inpaint_b200.trainstep import GaussianParamArena
    g = np.load(os.path.join(ROOT, "tests", "golden", "init_from_pcd.npz"))
    deg = int(g[f"{case}_deg"])
    pa = GaussianParamArena.create_from_pcd(g[f"{case}_points"], g[f"{case}_colors"], deg, "cpu",
                                            dist2=torch.from_numpy(g[f"{case}_dist2"].copy()))
    assert pa.M == (deg + 1) ** 2 and pa.active_sh_degree == 0 and pa.max_sh_degree == deg
    for name, got in (("_xyz", pa._xyz), ("_features_dc", pa._features_dc), ("_features_rest", pa._features_rest),
                      ("_scaling", pa._scaling), ("_rotation", pa._rotation), ("_opacity", pa._opacity)):
        want = torch.from_numpy(g[f"{case}{name}"].copy())
        assert got.shape == want.shape, name
        assert torch.equal(got.contiguous(), want), name
    assert torch.equal(pa._scaling[0], pa._scaling[1])      # the coincident pair shares its neighbourhood (one distance is 0)
    with pytest.raises(RuntimeError):
        GaussianParamArena.create_from_pcd(g[f"{case}_points"], g[f"{case}_colors"], deg, "cpu")   # no CPU neighbour search
