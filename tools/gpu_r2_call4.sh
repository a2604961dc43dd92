#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/exp_sort.py > gpurun_out/r2c4_sort.json 2> gpurun_out/r2c4_sort.err; echo "sort rc=$?"; cat gpurun_out/r2c4_sort.json; tail -3 gpurun_out/r2c4_sort.err
timeout 200 python -m pytest tests/test_scan_sort_gpu.py tests/test_batched_gpu.py tests/test_parity_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -5
for knob in 1 0; do
GSR_EXP_COUNT=$knob timeout 200 python - <<'PY' > gpurun_out/r2c4_bench_count$knob.txt 2>&1
import os, sys, json, subprocess
sys.argv = ["bench.py", "--steps", "30", "--warmup", "4", "--no-train-step", "--no-cpu-baseline", "--no-reference-structure"]
from multiview_inpaint_b200 import _C
_C.debug_set(1, int(os.environ["GSR_EXP_COUNT"]))
import bench
try:
    bench.run_ours(bench.parse())
except AssertionError as ex:
    print("assert", ex)
PY
python - "$knob" <<'PY'
import json, sys
try:
    lines = [l for l in open(f"gpurun_out/r2c4_bench_count{sys.argv[1]}.txt").read().strip().splitlines() if l.startswith("{")]
    d = json.loads(lines[-1])
    print("count_atomics", sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
    print({k: v["ms_per_view"] for k, v in d["stages"].items()})
except Exception as ex:
    print("no bench line:", ex); print(open(f"gpurun_out/r2c4_bench_count{sys.argv[1]}.txt").read()[-1500:])
PY
done
