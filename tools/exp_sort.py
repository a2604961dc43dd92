"""Throughput of the hand-written onesweep (gsr_sort_pairs_u32) against torch.sort (CUB) at the sizes the front end sorts:
V = 1.8 M depth keys (32 bits, 4 passes), N = 8.2 M tile keys (13 bits, 2 passes), and the 4-view batches of both.
Prints one JSON line: ms per sort, pairs/s, GB/s on 16 B per pair per pass, fraction of the measured HBM peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C  # noqa: E402

dev = torch.device("cuda")
peak = 6544.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = []
g = torch.Generator(device="cuda").manual_seed(1)
for n, bits in ((1_800_000, 32), (7_200_000, 32), (8_200_000, 13), (32_800_000, 13)):
    hi = (1 << bits) - 1 if bits < 32 else (1 << 31) - 1
    keys = torch.randint(0, hi, (n,), dtype=torch.int32, device=dev, generator=g)
    vals = torch.arange(n, dtype=torch.int32, device=dev)
    ms_t = timed(lambda: torch.sort(keys, stable=True))
    want_k, want_i = torch.sort(keys, stable=True)
    passes = (bits + 7) // 8
    row = {"n": n, "bits": bits, "passes": passes, "torch_sort_ms": round(ms_t, 4)}
    for mode, name in ((0, "match_any"), (1, "ballot"), (2, "atomic_or")):
        _C.debug_set(0, mode)
        k2, v2 = _C.sort_pairs(keys, vals, bits)
        ok = bool(torch.equal(k2, want_k) and torch.equal(v2.long(), want_i))
        ms = timed(lambda: _C.sort_pairs(keys, vals, bits))
        gbs = n * 16 * passes / (ms / 1e3) / 1e9
        row[name] = {"ms": round(ms, 4), "ok": ok, "gbps_16B_per_pair_pass": round(gbs, 1), "frac_hbm": round(gbs / peak, 3)}
    _C.debug_set(0, 1)
    out.append(row)
print(json.dumps(out))
