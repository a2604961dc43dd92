#!/bin/bash
# round 2, session 3, call 1 (one B200): tight binning (GSR_FLAG_TIGHT_BINNING) -- its tests first, the full GPU suite, then
# the bench with the literal lists (--flags 0) and the tight ones (--flags 32, the default), and the K1 prefetch variant.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tight_binning_gpu.py -m gpu -q -x --tb=short > gpurun_out/s3c1_tight.log 2>&1; echo "tight rc=$?"; tail -30 gpurun_out/s3c1_tight.log | cut -c1-400
timeout 420 python -m pytest tests -m gpu -q --maxfail=8 --tb=short > gpurun_out/s3c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3c1_pytest.log
tail -12 gpurun_out/s3c1_pytest.log | cut -c1-300
B="--steps 60 --warmup 4 --no-train-step --no-cpu-baseline --no-reference-structure --no-dropin"
for f in 0 32; do
  timeout 200 python bench.py $B --flags $f > gpurun_out/s3c1_bench_f$f.json 2> gpurun_out/s3c1_bench_f$f.err || tail -5 gpurun_out/s3c1_bench_f$f.err
done
GSR_LIB_VARIANT=k1pf timeout 200 python bench.py $B > gpurun_out/s3c1_bench_k1pf.json 2> gpurun_out/s3c1_bench_k1pf.err || tail -5 gpurun_out/s3c1_bench_k1pf.err
timeout 200 python bench.py $B --workload svd_orbit --total-views 25 --orbit-deg 30 --steps 20 --flags 0 > gpurun_out/s3c1_bench_svd_f0.json 2> gpurun_out/s3c1_bench_svd_f0.err || tail -5 gpurun_out/s3c1_bench_svd_f0.err
timeout 200 python bench.py $B --workload svd_orbit --total-views 25 --orbit-deg 30 --steps 20 > gpurun_out/s3c1_bench_svd_f32.json 2> gpurun_out/s3c1_bench_svd_f32.err || tail -5 gpurun_out/s3c1_bench_svd_f32.err
python - <<'PY'
import json
for tag in ("f0", "f32", "k1pf", "svd_f0", "svd_f32"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/s3c1_bench_{tag}.json") if l.startswith("{")][-1])
        print(tag, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "N", d["config"].get("N"), {k: v["ms_per_view"] for k, v in d["stages"].items()})
    except Exception as ex:
        print(tag, "no bench line:", ex)
PY
