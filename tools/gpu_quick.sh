#!/bin/bash
# quick one-GPU check after a kernel change: the parity / batched / api tests and a short bench without the extra legs
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_batched_gpu.py tests/test_api_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -4
timeout 200 python bench.py --steps 60 --warmup 4 --no-train-step --no-cpu-baseline --no-reference-structure ${GSR_BENCH_ARGS} > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/quick_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/quick_bench.json") if l.startswith("{")][-1])
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "dropin", d.get("dropin", {}).get("value"))
    print({k: v["ms_per_view"] for k, v in d["stages"].items()})
except Exception as ex:
    print("no bench line:", ex)
PY
