#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/exp_sort.py > gpurun_out/r2c3_sort.json 2> gpurun_out/r2c3_sort.err; echo "sort rc=$?"; cat gpurun_out/r2c3_sort.json; tail -3 gpurun_out/r2c3_sort.err
timeout 200 python bench.py --steps 50 --warmup 5 --no-train-step --no-cpu-baseline --no-reference-structure > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c3_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c3_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
    print({k: v["ms_per_view"] for k, v in d["stages"].items()})
except Exception as ex:
    print("no bench line:", ex)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/r2c3_launches.csv python bench.py --steps 3 --warmup 3 --no-train-step --no-cpu-baseline --no-reference-structure > gpurun_out/r2c3_ncu.log 2>&1; echo "ncu rc=$?"
