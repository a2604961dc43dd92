"""K2 / K4 through the C ABI: hand-written single-pass scan and onesweep radix sort, checked
against numpy (stable argsort) and against CUB (torch.sort(stable=True) / torch.cumsum run CUB
DeviceRadixSort / DeviceScan on CUDA) -- the libraries the reference calls."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C_():
    from multiview_inpaint_b200 import _C
    return _C


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 1023, 1024, 1025, 4096, 100_003, 3_000_000])
def test_inclusive_scan_matches_cumsum(C_, n):
    g = torch.Generator().manual_seed(n)
    x = torch.randint(0, 50, (n,), generator=g, dtype=torch.int32).cuda()
    out = C_.inclusive_scan(x)
    torch.testing.assert_close(out.long(), torch.cumsum(x.long(), 0), rtol=0, atol=0)


def test_inclusive_scan_with_gather_and_large_values(C_):
    n = 200_001
    g = torch.Generator().manual_seed(7)
    x = torch.randint(0, 20000, (n,), generator=g, dtype=torch.int32).cuda()   # total ~2e9 < 2^32 but > 2^31
    perm = torch.randperm(n, generator=g).int().cuda()
    out = C_.inclusive_scan(x, perm)
    ref = torch.cumsum(x[perm.long()].long(), 0)
    torch.testing.assert_close(out.long() & 0xFFFFFFFF, ref & 0xFFFFFFFF, rtol=0, atol=0)


def _ref_sort(keys_np, vals_np, end_bit):
    masked = keys_np & np.uint64((1 << end_bit) - 1) if end_bit < 64 else keys_np
    order = np.argsort(masked, kind="stable")
    return keys_np[order], vals_np[order]


@pytest.mark.parametrize("n,end_bit", [(1, 45), (255, 41), (3072, 45), (3073, 45), (50_000, 47), (1_000_003, 45), (777_777, 64), (12_345, 8), (12_345, 9)])
def test_sort_pairs_u64_stable_and_bit_exact(C_, n, end_bit):
    rng = np.random.default_rng(n + end_bit)
    tiles = rng.integers(0, 6300, n, dtype=np.uint64)
    depth = rng.integers(0, 64, n, dtype=np.uint64) << np.uint64(20)       # many ties -> stability matters
    junk = rng.integers(0, 1 << 16, n, dtype=np.uint64) << np.uint64(48)   # bits above end_bit must be ignored
    low = rng.integers(0, 1 << 12, n, dtype=np.uint64)                       # exercise the low digits too
    keys = (tiles << np.uint64(32)) | depth | low | (junk if end_bit < 48 else np.uint64(0))
    vals = rng.integers(0, 1 << 31, n, dtype=np.uint64).astype(np.uint32)
    k = torch.from_numpy(keys.view(np.int64)).cuda()
    v = torch.from_numpy(vals.view(np.int32)).cuda()
    ko, vo = C_.sort_pairs(k, v, end_bit)
    rk, rv = _ref_sort(keys, vals, end_bit)
    np.testing.assert_array_equal(ko.cpu().numpy().view(np.uint64), rk)
    np.testing.assert_array_equal(vo.cpu().numpy().view(np.uint32), rv)
    # inputs preserved
    np.testing.assert_array_equal(k.cpu().numpy().view(np.uint64), keys)


@pytest.mark.parametrize("n,end_bit", [(1, 32), (4096, 32), (4097, 13), (100_000, 13), (2_000_001, 32), (300_000, 15), (999, 5)])
def test_sort_pairs_u32_stable_and_bit_exact(C_, n, end_bit):
    rng = np.random.default_rng(n * 3 + end_bit)
    hi = 1 << min(end_bit, 20)
    keys = rng.integers(0, hi, n, dtype=np.uint64).astype(np.uint32)
    if end_bit == 32:
        keys = rng.random(n, dtype=np.float32).view(np.uint32)             # positive-float depth bits
        keys[rng.integers(0, n, n // 5)] = 0xFFFFFFFF                      # culled sentinel
    vals = np.arange(n, dtype=np.uint32)
    k = torch.from_numpy(keys.view(np.int32)).cuda()
    v = torch.from_numpy(vals.view(np.int32)).cuda()
    ko, vo = C_.sort_pairs(k, v, end_bit)
    rk, rv = _ref_sort(keys.astype(np.uint64), vals, end_bit)
    np.testing.assert_array_equal(ko.cpu().numpy().view(np.uint32), rk.astype(np.uint32))
    np.testing.assert_array_equal(vo.cpu().numpy().view(np.uint32), rv)


def test_sort_agrees_with_cub(C_):
    """torch.sort(stable=True) on CUDA int64 is cub::DeviceRadixSort -- the reference's sorter."""
    n = 2_500_000
    g = torch.Generator().manual_seed(99)
    tiles = torch.randint(0, 6300, (n,), generator=g, dtype=torch.int64)
    depth = torch.rand(n, generator=g).view(torch.int32).long() & 0xFFFFFFFF
    keys = ((tiles << 32) | depth).cuda()
    vals = torch.arange(n, dtype=torch.int32).cuda()
    ko, vo = C_.sort_pairs(keys, vals, 45)
    ck, ci = torch.sort(keys, stable=True)
    assert torch.equal(ko, ck) and torch.equal(vo.long(), ci)
    cs = C_.inclusive_scan(torch.ones(n, dtype=torch.int32).cuda())
    assert torch.equal(cs.long(), torch.arange(1, n + 1).cuda())


def test_sort_pairs_u32_beyond_2pow30(C_):
    """More than 2^30 pairs: the onesweep switches to 64-bit look-back words (a 30-bit prefix would wrap).
    Size-independent properties at a size no CPU reference sorts in seconds: the output is ordered on the masked
    key, equal keys keep their input order (values are the input positions), and the values are a permutation."""
    n = (1 << 30) + 123_457
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * (1 << 30):
        pytest.skip("needs ~40 GB of device memory")
    g = torch.Generator(device="cuda").manual_seed(3)
    # 15-bit tile ids (two 8-bit passes), skewed so that one digit alone exceeds 2^30 / 4 elements, + junk above end_bit
    keys = torch.randint(0, 32400, (n,), generator=g, dtype=torch.int32, device="cuda")
    keys[: n // 3] = 77
    keys |= (torch.randint(0, 4, (n,), generator=g, dtype=torch.int32, device="cuda") << 20)
    vals = torch.arange(n, dtype=torch.int32, device="cuda")          # wraps negative above 2^31: compared as uint32 below
    ko, vo = C_.sort_pairs(keys, vals, 15)
    del keys, vals
    torch.cuda.synchronize()
    bad = 0
    s = 0
    step = 1 << 27
    for a in range(0, n, step):
        b = min(n, a + step + 1)
        k = (ko[a:b] & 0x7FFF).long()
        v = vo[a:b].long() & 0xFFFFFFFF
        bad += int(((k[1:] < k[:-1]) | ((k[1:] == k[:-1]) & (v[1:] <= v[:-1]))).sum().item())
        s += int(v[: min(step, b - a)].sum().item())
    assert bad == 0
    assert s == n * (n - 1) // 2
    assert int((ko[: n // 3 + 100] & 0x7FFF).min().item()) == 0     # sorted front starts at tile 0
