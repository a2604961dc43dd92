"""(Lives under tests/ because it checks against and times the brute-force oracle, which only test infrastructure may use.)
distCUDA2 (csrc/knn.cu) timing on the GPU box: points/s at 100k / 1M / 3M / 10M points, against the
brute-force oracle on a bounded query sample (CPU, all host threads).  python tests/bench_knn.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multiview_inpaint_b200 import _C
from oracle import oracle as O

O.build()
res = []
for P in (100_000, 1_000_000, 3_000_000, 10_000_000):
    rng = np.random.default_rng(P)
    c = rng.uniform(-20, 20, size=(12, 3)); s = 10.0 ** rng.uniform(-2, 0.5, size=12); k = rng.integers(0, 12, size=P)
    pts = (c[k] + rng.normal(size=(P, 3)) * s[k, None]).astype(np.float32)
    x = torch.from_numpy(pts).cuda()
    for _ in range(2):
        out = _C.dist_cuda2(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _C.kernel_launches()
    e0.record()
    for _ in range(5):
        out = _C.dist_cuda2(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nq = max(64, min(2000, int(2e9 / P)))
    q = rng.integers(0, P, size=nq).astype(np.int32)
    t0 = time.perf_counter(); want = O.knn3_mean_dist2(pts, q); cpu_s = time.perf_counter() - t0
    ok = bool(np.array_equal(out.cpu().numpy()[q].view(np.uint32), want.view(np.uint32)))
    res.append(dict(P=P, ms=round(ms, 3), mpoints_per_s=round(P / ms / 1e3, 1), launches=(_C.kernel_launches() - l0) // 5,
                    cpu_bruteforce_points_per_s=round(nq / cpu_s, 1), cpu_threads=O.num_threads(), sample_bit_exact=ok))
    print(json.dumps(res[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/knn_bench.json", "w"), indent=1)
