"""GPU parity of gsr_knn3_mean_dist2 (csrc/knn.cu) == simple_knn._C.distCUDA2 of the reference
(gs-simp/scene/gaussian_model.py:20,134,546,623) against the brute-force CPU oracle: BIT-EXACT
(exact nearest neighbours, same fp32 op order), through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda(pts):
    from simple_knn._C import distCUDA2
    out = distCUDA2(torch.from_numpy(pts).cuda())
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _cloud(P, seed, kind):
    rng = np.random.default_rng(seed)
    if kind == "normal":
        return rng.normal(size=(P, 3)).astype(np.float32)
    if kind == "uniform":
        return rng.uniform(-3, 7, size=(P, 3)).astype(np.float32)
    if kind == "clusters":   # SfM-like: dense blobs of very different scale + sparse background
        c = rng.uniform(-20, 20, size=(12, 3))
        s = 10.0 ** rng.uniform(-3, 0.5, size=12)
        k = rng.integers(0, 12, size=P)
        p = c[k] + rng.normal(size=(P, 3)) * s[k, None]
        p[: P // 20] = rng.uniform(-100, 100, size=(P // 20, 3))
        return p.astype(np.float32)
    if kind == "plane":      # zero extent on one axis, a line on another range
        p = rng.uniform(0, 1, size=(P, 3))
        p[:, 2] = 0.25
        return p.astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 31, 32, 33, 64, 65, 1023, 1025])
def test_small_and_ragged_sizes_bit_exact(oracle, P):
    pts = _cloud(P, 100 + P, "normal")
    assert np.array_equal(_cuda(pts).view(np.uint32), oracle.knn3_mean_dist2(pts).view(np.uint32))


@pytest.mark.parametrize("kind,P", [("normal", 20000), ("uniform", 40000), ("clusters", 30000), ("plane", 10000)])
def test_clouds_bit_exact(oracle, kind, P):
    pts = _cloud(P, 7, kind)
    got, want = _cuda(pts), oracle.knn3_mean_dist2(pts)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.abs(got - want).max()


def test_coincident_points_and_empty(oracle):
    from simple_knn._C import distCUDA2
    assert distCUDA2(torch.zeros(0, 3, device="cuda")).shape == (0,)
    pts = _cloud(5000, 1, "normal")
    pts[1000:1200] = pts[3]          # a 201-fold point: distances 0
    pts[2000:2002] = pts[4]          # a triple: two zeros + one real distance
    got, want = _cuda(pts), oracle.knn3_mean_dist2(pts)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert got[1000] == 0.0
    allsame = np.full((70000, 3), 1.5, np.float32)     # degenerate: must not visit every leaf (finishes fast)
    assert np.array_equal(_cuda(allsame), np.zeros(70000, np.float32))


def test_reference_call_pattern_and_strided_input(oracle):
    """gaussian_model.py:134: clamp_min(distCUDA2(points.float().cuda()), 1e-7) -> log(sqrt()) scales."""
    from simple_knn._C import distCUDA2
    pts = _cloud(3000, 5, "uniform")
    big = torch.from_numpy(np.concatenate([pts, pts], axis=1)).cuda()
    view = big[:, 3:]                                  # non-contiguous (P,3) view
    dist2 = torch.clamp_min(distCUDA2(view), 0.0000001)
    scales = torch.log(torch.sqrt(dist2))[..., None].repeat(1, 3)
    assert scales.shape == (3000, 3) and torch.isfinite(scales).all()
    assert np.array_equal(distCUDA2(view).cpu().numpy(), oracle.knn3_mean_dist2(pts))


def test_full_size_sampled_bit_exact(oracle):
    """3 M points (the headline scene's size): 1500 sampled queries against the brute-force scan, plus the
    size-independent properties: every value positive/finite, invariant under a permutation of the input."""
    P = 3_000_000
    pts = _cloud(P, 11, "clusters")
    got = _cuda(pts)
    q = np.random.default_rng(5).integers(0, P, size=1500).astype(np.int32)
    want = oracle.knn3_mean_dist2(pts, q)
    assert np.array_equal(got[q].view(np.uint32), want.view(np.uint32))
    assert np.isfinite(got).all() and (got >= 0).all()
    perm = np.random.default_rng(6).permutation(P)
    got_p = _cuda(np.ascontiguousarray(pts[perm]))
    assert np.array_equal(got_p.view(np.uint32), got[perm].view(np.uint32))
