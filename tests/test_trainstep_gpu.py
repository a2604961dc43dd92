"""GPU parity of the training-step stages, through the C ABI (multiview_inpaint_b200._C):
fused L1+SSIM loss, one-kernel activations, one-launch Adam, and the fused step that chains them with the rasterizer.

Checked against (1) the golden vectors produced by the REFERENCE's own utils/loss_utils.py, torch autograd and
torch.optim.Adam (tests/golden/trainstep.npz), (2) the numpy oracle (oracle/trainstep_oracle.py) on fresh seeded
inputs including ragged shapes, and (3) at BASELINE.json's full image size, properties and a torch restatement.
Tolerances (floating point, stated per test): loss scalars 1e-5 absolute; image gradient 1e-4 of its max;
activations 4 ulp; Adam positions 3e-7 absolute per step."""
import os

import numpy as np
import pytest
import torch

from oracle import trainstep_oracle as T
from tests.util import rel_err, small_scene

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda"


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "trainstep.npz"))


@pytest.fixture(scope="module")
def C():
    from multiview_inpaint_b200 import _C
    return _C


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(DEV)


# ------------------------------------------------------------------------------------------------ loss
@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_loss_matches_reference_golden(C, gold, case):
    img, gt = cu(gold[f"loss_{case}_img"]), cu(gold[f"loss_{case}_gt"])
    lam = float(gold["loss_lambda"])
    out3, temp = C.loss_l1_ssim_forward(img, gt, lam)
    ref = gold[f"loss_{case}_out"]
    got = out3.cpu().numpy().astype(np.float64)
    assert np.abs(got - ref).max() < 1e-5, (got, ref)
    g = C.loss_l1_ssim_backward(img, gt, lam, temp).cpu().numpy()
    gref = gold[f"loss_{case}_grad"]
    assert np.abs(g - gref).max() <= 1e-4 * np.abs(gref).max(), rel_err(g, gref)
    # exactly-equal pixels carry no L1 gradient (sign(0) = 0): identical to the reference there as well
    eq = gold[f"loss_{case}_img"] == gold[f"loss_{case}_gt"]
    assert eq.any()
    assert np.abs(g[eq] - gref[eq]).max() <= 1e-4 * np.abs(gref).max()


@pytest.mark.parametrize("shape,lam", [((3, 1, 1), 0.2), ((3, 5, 200), 0.2), ((3, 130, 33), 0.0), ((3, 65, 97), 1.0),
                                       ((1, 16, 32), 0.5), ((4, 17, 31), 0.3)])
def test_loss_matches_oracle_ragged(C, shape, lam):
    rng = np.random.default_rng(sum(shape))
    gt = rng.random(shape, dtype=np.float32)
    img = np.clip(gt + 0.2 * rng.standard_normal(shape).astype(np.float32), 0, 1).astype(np.float32)
    out3, temp = C.loss_l1_ssim_forward(cu(img), cu(gt), lam)
    ref = T.loss_forward(img, gt, lam)
    assert np.abs(out3.cpu().numpy().astype(np.float64) - np.array(ref)).max() < 1e-5
    g = C.loss_l1_ssim_backward(cu(img), cu(gt), lam, temp).cpu().numpy()
    gref = T.loss_backward(img, gt, lam)
    assert np.abs(g - gref).max() <= 1e-4 * np.abs(gref).max() + 1e-12, rel_err(g, gref)


def test_loss_seed_scales_gradient_and_autograd_function(C):
    from multiview_inpaint_b200.trainstep import l1_ssim_loss
    torch.manual_seed(5)
    gt = torch.rand(3, 40, 56, device=DEV)
    img = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1).requires_grad_(True)
    loss, l1, ss = l1_ssim_loss(img, gt, 0.2)
    (2.5 * loss).backward()
    out3, temp = C.loss_l1_ssim_forward(img.detach(), gt, 0.2)
    g1 = C.loss_l1_ssim_backward(img.detach(), gt, 0.2, temp)
    assert torch.equal(torch.stack([l1, ss, loss]), out3)
    assert torch.allclose(img.grad, 2.5 * g1, rtol=1e-6, atol=1e-12)
    ref = T.loss_forward(img.detach().cpu().numpy(), gt.cpu().numpy(), 0.2)
    assert abs(loss.item() - ref[2]) < 1e-5 and abs(l1.item() - ref[0]) < 1e-5


def _torch_l1_ssim(img, gt, lam):
    """torch restatement used as the checker at sizes the numpy oracle is too slow for (depthwise 11x11 conv)."""
    import torch.nn.functional as F
    Cn = img.shape[0]
    w1 = torch.as_tensor(T.gaussian_window(), device=img.device)
    w2 = torch.outer(w1, w1).expand(Cn, 1, 11, 11).contiguous()
    blur = lambda t: F.conv2d(t[None], w2, padding=5, groups=Cn)[0]
    mx, my = blur(img), blur(gt)
    vx, vy, cxy = blur(img * img) - mx * mx, blur(gt * gt) - my * my, blur(img * gt) - mx * my
    s = ((2 * mx * my + T.C1) * (2 * cxy + T.C2)) / ((mx * mx + my * my + T.C1) * (vx + vy + T.C2))
    l1 = (img - gt).abs().mean()
    return (1 - lam) * l1 + lam * (1 - s.mean()), l1, s.mean()


def test_loss_full_size_properties_and_torch(C):
    """BASELINE.json headline image (3,1008,1600): identities + a torch conv2d restatement on the GPU."""
    torch.manual_seed(11)
    H, W = 1008, 1600
    gt = torch.rand(3, H, W, device=DEV)
    gt = torch.nn.functional.avg_pool2d(gt[None], 5, 1, 2)[0].contiguous()      # some spatial structure
    img = (gt + 0.05 * torch.randn_like(gt)).clamp(0, 1).contiguous()
    # identical images: Ll1 = 0, ssim = 1, loss = 0, gradient ~ 0 (optimum of both terms, sign(0) = 0)
    out3, temp = C.loss_l1_ssim_forward(gt, gt, 0.2)
    o = out3.cpu().numpy()
    assert o[0] == 0.0 and abs(o[1] - 1.0) < 1e-6 and abs(o[2]) < 1e-6
    g0 = C.loss_l1_ssim_backward(gt, gt, 0.2, temp)
    assert g0.abs().max().item() < 1e-9
    # symmetry of ssim and of |x - y|
    a, _ = C.loss_l1_ssim_forward(img, gt, 0.2)
    b, _ = C.loss_l1_ssim_forward(gt, img, 0.2)
    assert torch.allclose(a, b, rtol=0, atol=1e-6)
    # against torch
    x = img.clone().requires_grad_(True)
    loss, l1, ss = _torch_l1_ssim(x, gt, 0.2)
    loss.backward()
    out3, temp = C.loss_l1_ssim_forward(img, gt, 0.2)
    got = out3.cpu().numpy()
    assert abs(got[0] - l1.item()) < 1e-5 and abs(got[1] - ss.item()) < 1e-5 and abs(got[2] - loss.item()) < 1e-5
    g = C.loss_l1_ssim_backward(img, gt, 0.2, temp)
    assert (g - x.grad).abs().max().item() <= 1e-4 * x.grad.abs().max().item()
    # run-to-run determinism (fixed-order reductions, no atomics)
    out3b, tempb = C.loss_l1_ssim_forward(img, gt, 0.2)
    assert torch.equal(out3, out3b) and torch.equal(g, C.loss_l1_ssim_backward(img, gt, 0.2, tempb))


def test_loss_rejects_bad_arguments(C):
    x = torch.rand(3, 8, 8, device=DEV)
    with pytest.raises(RuntimeError):
        C.loss_l1_ssim_forward(x, torch.rand(3, 8, 9, device=DEV), 0.2)
    with pytest.raises(RuntimeError, match="CUDA"):
        C.loss_l1_ssim_forward(x.cpu(), x.cpu(), 0.2)
    with pytest.raises(RuntimeError):   # temp too small -> GSR_E_INVALID
        C.loss_l1_ssim_backward(x, x, 0.2, torch.empty(16, dtype=torch.uint8, device=DEV))


# ------------------------------------------------------------------------------------------------ activations
def test_activations_match_golden(C, gold):
    rs, rq, ro = cu(gold["act_raw_s"]), cu(gold["act_raw_q"]), cu(gold["act_raw_o"])
    s, q, o = C.activate_forward(rs, rq, ro)
    # same formulas as torch (exp, x / max(|x|, eps), 1 / (1 + exp(-x))); device expf vs host expf: <= 4 ulp
    np.testing.assert_allclose(s.cpu().numpy(), gold["act_s"], rtol=5e-7)
    np.testing.assert_allclose(q.cpu().numpy(), gold["act_q"], rtol=5e-7, atol=1e-9)
    np.testing.assert_allclose(o.cpu().numpy(), gold["act_o"], rtol=5e-7)
    assert torch.all(q[3] == 0)
    gs, gq, go = cu(gold["act_gs"]), cu(gold["act_gq"]), cu(gold["act_go"])
    C.activate_backward(rs, rq, ro, gs, gq, go)           # in place
    np.testing.assert_allclose(gs.cpu().numpy(), gold["act_d_raw_s"], rtol=2e-6)
    np.testing.assert_allclose(go.cpu().numpy(), gold["act_d_raw_o"], rtol=1e-4, atol=1e-9)   # (1 - y) cancels in fp32
    ref = gold["act_d_raw_q"]
    ok = np.ones(len(ref), bool)
    ok[3] = False
    np.testing.assert_allclose(gq.cpu().numpy()[ok], ref[ok], rtol=1e-4, atol=2e-6 * np.abs(ref[ok]).max())
    np.testing.assert_allclose(gq.cpu().numpy()[3], ref[3], rtol=1e-6)       # |q| < eps: g / eps


def test_activations_subset_and_large(C):
    torch.manual_seed(2)
    P = 1_000_003
    rs, rq, ro = torch.randn(P, 3, device=DEV), torch.randn(P, 4, device=DEV), torch.randn(P, 1, device=DEV) * 3
    s, q, o = C.activate_forward(rs, rq, ro)
    assert torch.allclose(s, torch.exp(rs), rtol=5e-7, atol=0)
    assert torch.allclose(q, torch.nn.functional.normalize(rq), rtol=5e-7, atol=1e-9)
    assert torch.allclose(o, torch.sigmoid(ro), rtol=5e-7, atol=0)
    s2, q2, o2 = C.activate_forward(rs, None, None)
    assert q2 is None and o2 is None and torch.equal(s2, s)
    # chain rule against autograd
    rs_, rq_, ro_ = (t.clone().requires_grad_(True) for t in (rs, rq, ro))
    gs, gq, go = torch.randn_like(rs), torch.randn_like(rq), torch.randn_like(ro)
    ((torch.exp(rs_) * gs).sum() + (torch.nn.functional.normalize(rq_) * gq).sum() + (torch.sigmoid(ro_) * go).sum()).backward()
    C.activate_backward(rs, rq, ro, gs, gq, go)
    assert rel_err(gs.cpu().numpy(), rs_.grad.cpu().numpy()) < 1e-6
    assert rel_err(gq.cpu().numpy(), rq_.grad.cpu().numpy()) < 1e-5
    assert rel_err(go.cpu().numpy(), ro_.grad.cpu().numpy()) < 1e-6


# ------------------------------------------------------------------------------------------------ Adam
GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def test_adam_matches_torch_optim_golden(C, gold):
    """Five steps of the reference's optimizer (six groups) against ONE launch per step over five segments,
    f_dc / f_rest living in one (P,M,3) tensor with a column-dependent learning rate."""
    from multiview_inpaint_b200.multiview import GradArena
    from multiview_inpaint_b200.trainstep import GaussianParamArena
    M = int(gold["adam_M"])
    pa = GaussianParamArena.from_tensors(cu(gold["adam_p0_xyz"]), cu(gold["adam_p0_f_dc"]), cu(gold["adam_p0_f_rest"]),
                                         cu(gold["adam_p0_opacity"]), cu(gold["adam_p0_scaling"]), cu(gold["adam_p0_rotation"]))
    P = pa.P
    assert pa.M == M
    arena = GradArena(P, M, DEV)
    lrs = {k: float(gold[f"adam_lr_{k}"]) for k in GROUPS}
    for t in range(int(gold["adam_steps"])):
        arena.views["dL_dmeans3D"].copy_(cu(gold[f"adam_g{t}_xyz"]))
        arena.views["dL_dsh"][:, :1].copy_(cu(gold[f"adam_g{t}_f_dc"]))
        arena.views["dL_dsh"][:, 1:].copy_(cu(gold[f"adam_g{t}_f_rest"]))
        arena.views["dL_dopacity"].copy_(cu(gold[f"adam_g{t}_opacity"]))
        arena.views["dL_dscales"].copy_(cu(gold[f"adam_g{t}_scaling"]))
        arena.views["dL_drotations"].copy_(cu(gold[f"adam_g{t}_rotation"]))
        # Adam only (the golden gradients are already with respect to the raw parameters)
        pa.step_count += 1
        segs = []
        for pname, gname, kw in (("_xyz", "dL_dmeans3D", dict(lr=lrs["xyz"])),
                                 ("_features", "dL_dsh", dict(lr=lrs["f_dc"], lr_rest=lrs["f_rest"], row_len=3 * M, row_split=3)),
                                 ("_opacity", "dL_dopacity", dict(lr=lrs["opacity"])),
                                 ("_scaling", "dL_dscales", dict(lr=lrs["scaling"])),
                                 ("_rotation", "dL_drotations", dict(lr=lrs["rotation"]))):
            m, v = pa.moments(pname)
            segs.append(dict(param=getattr(pa, pname), grad=arena.views[gname], exp_avg=m, exp_avg_sq=v, **kw))
        C.adam_step(segs, pa.step_count)
        got = {"xyz": pa._xyz, "f_dc": pa._features_dc, "f_rest": pa._features_rest, "opacity": pa._opacity,
               "scaling": pa._scaling, "rotation": pa._rotation}
        mom = {"xyz": "_xyz", "opacity": "_opacity", "scaling": "_scaling", "rotation": "_rotation"}
        for k in GROUPS:
            np.testing.assert_allclose(got[k].cpu().numpy(), gold[f"adam_p{t + 1}_{k}"], rtol=0, atol=3e-7, err_msg=f"{k} step {t + 1}")
            if k in mom:
                m, v = pa.moments(mom[k])
                mr, vr = gold[f"adam_m{t + 1}_{k}"], gold[f"adam_v{t + 1}_{k}"]
                np.testing.assert_allclose(m.cpu().numpy(), mr, rtol=3e-6, atol=3e-7 * np.abs(mr).max())
                np.testing.assert_allclose(v.cpu().numpy(), vr, rtol=3e-6, atol=3e-7 * np.abs(vr).max())
        m, v = pa.moments("_features")
        mr = np.concatenate([gold[f"adam_m{t + 1}_f_dc"], gold[f"adam_m{t + 1}_f_rest"]], axis=1)
        np.testing.assert_allclose(m.cpu().numpy(), mr, rtol=3e-6, atol=3e-7 * np.abs(mr).max())


@pytest.mark.parametrize("n,row_len,row_split", [(1, 0, 0), (4099, 0, 0), (3 * 1001, 3, 3), (48 * 777, 48, 3),
                                                 (27 * 333, 27, 3), (12 * 100_003, 12, 3)])
def test_adam_segments_against_torch_optim(C, n, row_len, row_split):
    """Flat and row-structured segments, vector and scalar paths (unaligned views), against torch.optim.Adam on
    the same device with two parameter groups selected by a column mask."""
    torch.manual_seed(n)
    base = torch.randn(n + 1, device=DEV)
    for off in (0, 1):                                    # off = 1: pointer not 16-byte aligned -> scalar path
        holder = base.clone()
        p = holder[off:off + n]
        m, v = torch.zeros_like(base)[off:off + n], torch.zeros_like(base)[off:off + n]
        p_ref = p.clone()
        lr_a, lr_b = 0.01, 0.0005
        if row_len:
            col = torch.arange(n, device=DEV) % row_len
            lr_t = torch.where(col < row_split, torch.tensor(lr_a, device=DEV), torch.tensor(lr_b, device=DEV))
        else:
            lr_t = torch.full((n,), lr_a, device=DEV)
        m_ref, v_ref = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
        for step in range(1, 4):
            g = torch.randn(n, device=DEV) * 10.0 ** float(step - 3)
            C.adam_step([dict(param=p, grad=g, exp_avg=m, exp_avg_sq=v, lr=lr_a, lr_rest=lr_b, row_len=row_len,
                              row_split=row_split)], step)
            # torch/optim/adam.py _single_tensor_adam, elementwise with a learning-rate tensor
            m_ref.lerp_(g, 1 - 0.9)
            v_ref.mul_(0.999).addcmul_(g, g, value=1 - 0.999)
            bc1, bc2 = 1 - 0.9 ** step, 1 - 0.999 ** step
            denom = (v_ref.sqrt() / (bc2 ** 0.5)).add_(1e-15)
            p_ref -= (lr_t / bc1) * (m_ref / denom)
            assert (p - p_ref).abs().max().item() <= 3e-7 * max(1.0, p_ref.abs().max().item()), (off, step)
        assert torch.allclose(m, m_ref, rtol=3e-6, atol=3e-7 * m_ref.abs().max().item())
        assert torch.allclose(v, v_ref, rtol=3e-6, atol=3e-7 * v_ref.abs().max().item())
        if off == 1:
            assert holder[0] == base[0]                   # nothing written in front of the segment


def test_adam_rejects_bad_arguments(C):
    x = torch.zeros(8, device=DEV)
    seg = dict(param=x, grad=x.clone(), exp_avg=x.clone(), exp_avg_sq=x.clone(), lr=0.1)
    with pytest.raises(RuntimeError):
        C.adam_step([seg], 0)                              # step is 1-based
    with pytest.raises(RuntimeError):
        C.adam_step([seg] * 9, 1)                          # at most 8 segments
    with pytest.raises(RuntimeError):
        C.adam_step([dict(seg, row_len=4, row_split=5)], 1)


# ------------------------------------------------------------------------------------------------ the fused step
@pytest.mark.parametrize("deg,n_views", [(1, 2), (3, 3), (0, 1)])
def test_fused_step_equals_reference_structure_step(C, deg, n_views):
    """One iteration of train.py:86-128 two ways on the same raw parameters:
      (reference structure) torch getters (exp / normalize / sigmoid / cat) -> diff_gaussian_rasterization drop-in
                            -> torch l1 + ssim -> autograd -> per-group gradients
      (fused)               GaussianParamArena.activate -> multi-view step with ViewLoss -> in-place chain rule.
    Raw-parameter gradients must agree to 1e-3 of their max (BASELINE.json's gradient tolerance), and the fused Adam
    applied to them must equal the oracle's Adam on the same gradients."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from multiview_inpaint_b200 import scenes as S
    from multiview_inpaint_b200.multiview import GradArena
    from multiview_inpaint_b200.trainstep import GROUPS as G, GaussianParamArena, ViewLoss
    sc = small_scene(P=2501, W=112, H=80, deg=deg, seed=40 + deg)
    M = sc["shs"].shape[1]
    raw = dict(xyz=sc["means3D"], f_dc=sc["shs"][:, :1].contiguous(), f_rest=sc["shs"][:, 1:].contiguous(),
               opacity=torch.logit(sc["opacities"].clamp(1e-4, 1 - 1e-4)).reshape(-1, 1), scaling=torch.log(sc["scales"]),
               rotation=sc["rotations"] * 1.7)
    raw = {k: v.to(DEV) for k, v in raw.items()}
    cams = [c.to(DEV) for c in S.orbit_cameras(max(n_views, 2), 112, 80, max_deg=6.0)][:n_views]
    bg = torch.zeros(3, device=DEV)
    torch.manual_seed(9)
    gts = [torch.rand(3, 80, 112, device=DEV) for _ in range(n_views)]
    settings = [GaussianRasterizationSettings(image_height=80, image_width=112, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                                              scale_modifier=1.0, viewmatrix=c.world_view_transform,
                                              projmatrix=c.full_proj_transform, sh_degree=deg, campos=c.camera_center,
                                              prefiltered=False) for c in cams]
    # ---- reference structure, torch autograd ----
    leaves = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
    total = 0.0
    for rs, gt in zip(settings, gts):
        shs = torch.cat((leaves["f_dc"], leaves["f_rest"]), dim=1)
        color, radii, depth = GaussianRasterizer(rs)(
            means3D=leaves["xyz"], means2D=torch.zeros_like(leaves["xyz"], requires_grad=True), opacities=torch.sigmoid(leaves["opacity"]),
            shs=shs, scales=torch.exp(leaves["scaling"]), rotations=torch.nn.functional.normalize(leaves["rotation"]))
        total = total + _torch_l1_ssim(color, gt, 0.2)[0]
    total.backward()
    # ---- fused ----
    pa = GaussianParamArena.from_tensors(raw["xyz"], raw["f_dc"], raw["f_rest"], raw["opacity"], raw["scaling"], raw["rotation"])
    arena = GradArena(pa.P, M, DEV)
    losses = [ViewLoss(gt, 0.2) for gt in gts]
    from multiview_inpaint_b200.multiview import cuda_views_fwd_bwd
    before = pa.param.clone()
    cuda_views_fwd_bwd(pa.activate(), settings, losses, arena)
    lsum = sum(l.out3[2].item() for l in losses)
    assert abs(lsum - total.item()) < 1e-5 * n_views
    lrs = dict(xyz=0.00016, f_dc=0.0025, f_rest=0.0025 / 20, opacity=0.05, scaling=0.005, rotation=0.001)
    pa.apply_gradients(arena, lrs)
    torch.cuda.synchronize()
    g = arena.views
    got = dict(xyz=g["dL_dmeans3D"], f_dc=g["dL_dsh"][:, :1], f_rest=g["dL_dsh"][:, 1:], opacity=g["dL_dopacity"],
               scaling=g["dL_dscales"], rotation=g["dL_drotations"])
    for k in G:
        if leaves[k].numel() == 0:
            continue
        ref = leaves[k].grad.cpu().numpy()
        assert rel_err(got[k].cpu().numpy(), ref) < 1e-3, (k, rel_err(got[k].cpu().numpy(), ref))
    # Adam: oracle on the fused step's own gradients, per-column learning rates for the SH tensor
    assert pa.step_count == 1
    for pname, gname, lr in (("_xyz", "dL_dmeans3D", lrs["xyz"]), ("_opacity", "dL_dopacity", lrs["opacity"]),
                             ("_scaling", "dL_dscales", lrs["scaling"]), ("_rotation", "dL_drotations", lrs["rotation"])):
        p0 = before[pa._offs[pname]:pa._offs[pname] + getattr(pa, pname).numel()].cpu().numpy().reshape(getattr(pa, pname).shape)
        p1, m1, v1 = T.adam_step(p0, g[gname].cpu().numpy().reshape(p0.shape), np.zeros_like(p0), np.zeros_like(p0), 1, lr)
        np.testing.assert_allclose(getattr(pa, pname).cpu().numpy(), p1, rtol=0, atol=2e-7 * max(1.0, np.abs(p1).max()))
    p0 = before[pa._offs["_features"]:pa._offs["_features"] + pa._features.numel()].cpu().numpy().reshape(pa.P, M, 3)
    lr = np.full_like(p0, lrs["f_rest"])
    lr[:, :1] = lrs["f_dc"]
    p1, _, _ = T.adam_step(p0, g["dL_dsh"].cpu().numpy(), np.zeros_like(p0), np.zeros_like(p0), 1, lr)
    np.testing.assert_allclose(pa._features.cpu().numpy(), p1, rtol=0, atol=2e-7 * max(1.0, np.abs(p1).max()))




@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_create_from_pcd_on_the_device_matches_reference(case):
    """The whole initialisation on the GPU -- neighbour distances from gsr_knn3_mean_dist2 (bit-identical to the
    brute-force oracle that answered the reference's distCUDA2 call when the golden was made) -- must give the
    reference's six tensors (gaussian_model.py:124-147); log / sqrt on the device vs the golden's CPU torch: 2e-6."""
    from multiview_inpaint_b200.trainstep import GaussianParamArena
    g = np.load(os.path.join(ROOT, "tests", "golden", "init_from_pcd.npz"))
    deg = int(g[f"{case}_deg"])
    pa = GaussianParamArena.create_from_pcd(g[f"{case}_points"], g[f"{case}_colors"], deg, DEV)
    for name, got in (("_xyz", pa._xyz), ("_features_rest", pa._features_rest), ("_rotation", pa._rotation)):
        assert torch.equal(got.contiguous().cpu(), torch.from_numpy(g[f"{case}{name}"].copy())), name
    # RGB2SH = (c - 0.5) / C0: torch's CUDA kernel multiplies by the reciprocal of a scalar divisor, the golden's CPU
    # kernel divides (first device run, r2c1_pytest.log: 1-ulp differences) -- same tolerance as log / sqrt below
    for name, got in (("_features_dc", pa._features_dc.contiguous()), ("_scaling", pa._scaling), ("_opacity", pa._opacity)):
        want = torch.from_numpy(g[f"{case}{name}"].copy())
        assert float((got.cpu() - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max())), name


def test_ply_round_trip_from_the_device(tmp_path):
    from multiview_inpaint_b200 import plyio
    from multiview_inpaint_b200.trainstep import GaussianParamArena
    pa = GaussianParamArena(1000, 16, DEV)
    pa.param.normal_()
    path = str(tmp_path / "point_cloud.ply")
    plyio.save_ply(path, pa)
    back = plyio.load_ply(path, DEV, sh_degree=3)
    for name in ("_xyz", "_features", "_opacity", "_scaling", "_rotation"):
        assert torch.equal(getattr(back, name), getattr(pa, name)), name
    g = back.activate()
    assert g["shs"].shape == (1000, 16, 3) and bool(torch.isfinite(g["scales"]).all())
