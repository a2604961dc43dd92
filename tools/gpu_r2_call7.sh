#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/exp_sort.py > gpurun_out/r2c7_sort.json 2> gpurun_out/r2c7_sort.err; echo "sort rc=$?"; cat gpurun_out/r2c7_sort.json; tail -3 gpurun_out/r2c7_sort.err
timeout 200 python -m pytest tests/test_scan_sort_gpu.py tests/test_batched_gpu.py tests/test_parity_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -5
for pre in 4 5 6; do
GSR_RANK=2 GSR_PRE=$pre timeout 200 python - <<'PY' > gpurun_out/r2c7_bench_p$pre.txt 2>&1
import os, sys
sys.argv = ["bench.py", "--steps", "40", "--warmup", "4", "--no-train-step", "--no-cpu-baseline", "--no-reference-structure", "--streams", "1"]
from multiview_inpaint_b200 import _C
_C.debug_set(0, int(os.environ["GSR_RANK"]))
_C.debug_set(2, int(os.environ["GSR_PRE"]))
import bench
bench.run_ours(bench.parse())
PY
python - "$pre" <<'PY'
import json, sys
try:
    lines = [l for l in open(f"gpurun_out/r2c7_bench_p{sys.argv[1]}.txt").read().strip().splitlines() if l.startswith("{")]
    d = json.loads(lines[-1])
    print("pre_min_blocks", sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
    print({k: v["ms_per_view"] for k, v in d["stages"].items()})
except Exception as ex:
    print("no bench line:", ex); print(open(f"gpurun_out/r2c7_bench_p{sys.argv[1]}.txt").read()[-1500:])
PY
done
