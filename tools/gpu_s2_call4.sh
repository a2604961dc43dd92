#!/bin/bash
# round 2, session 2, call 4 (one B200): full GPU suite, default bench line, BASELINE configs 2-5 at N = 1 (the same command
# lines the multi-GPU call used), compute-sanitizer memcheck + racecheck over the parity / binning / batched tests.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --maxfail=8 --tb=short > gpurun_out/s2c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2c4_pytest.log
tail -4 gpurun_out/s2c4_pytest.log
timeout 420 python bench.py > gpurun_out/s2c4_bench.json 2> gpurun_out/s2c4_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s2c4_bench.err
B="--no-cpu-baseline --no-reference-structure --no-train-step --no-dropin --warmup 3"
timeout 600 python tools/exp_configs_multi.py "--workload mip360 --steps 50 $B" "--workload svd_orbit --total-views 25 --orbit-deg 30 --steps 20 $B" "--mode infer --workload inference --views-per-rank 25 --orbit-deg 30 --steps 5 $B" "--workload stress --views-per-rank 1 --steps 3 $B" > gpurun_out/s2c4_configs_n1.jsonl 2> gpurun_out/s2c4_configs_n1.err; echo "configs1 rc=$?"; tail -3 gpurun_out/s2c4_configs_n1.err | cut -c1-300
python - <<'PY'
import json
def show(path, tag):
    try:
        lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
        for l in lines:
            d = json.loads(l); c = d["config"]
            print(tag, c["workload"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), "N", c.get("N"), "V", c.get("V"))
            if "stages" in d: print("   ", {k: v["ms_per_view"] for k, v in d["stages"].items()})
            for k in ("dropin", "train_step", "parity_headline"):
                if k in d: print("   ", k, str(d[k])[:400])
    except Exception as ex:
        print(tag, "no bench line:", ex); print(open(path).read()[-1500:])
show("gpurun_out/s2c4_bench.json", "default")
show("gpurun_out/s2c4_configs_n1.jsonl", "config")
PY
SEL="config1 or tiny_capacity or dense_opaque or tensor_core or overflow or depth_ties or different_sizes or small_scenes"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py tests/test_batched_gpu.py tests/test_scan_sort_gpu.py -m gpu -q -k "$SEL or stable or cumsum" > gpurun_out/s2c4_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/s2c4_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py tests/test_batched_gpu.py -m gpu -q -k "config1 or tensor_core or dense_opaque or different_sizes" > gpurun_out/s2c4_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/s2c4_racecheck.log
