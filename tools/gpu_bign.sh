#!/bin/bash
# instance counts beyond 2^30: sort test + the stress config at the survey's 40 px median
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_scan_sort_gpu.py -m gpu -q -x > gpurun_out/pytest_sort.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_sort.log; tail -15 gpurun_out/pytest_sort.log | cut -c1-300
timeout 500 python bench.py --workload stress --views-per-rank 2 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-structure --no-train-step > gpurun_out/cfg_stress.json 2> gpurun_out/cfg_stress.err; tail -4 gpurun_out/cfg_stress.err | cut -c1-300
python - <<'PY'
import json
for l in open("gpurun_out/cfg_stress.json"):
    if l.startswith("{"):
        d = json.loads(l); c = d["config"]
        print("stress value", round(d["value"], 2), "ms/view", round(d["ms_per_step"] / 2, 2), "N", c["N"], "V", c["V"], "reserved GB", d["alloc"]["reserved_gb"])
        print({k: v["ms_per_view"] for k, v in d["stages"].items()})
PY
