#!/bin/bash
mkdir -p gpurun_out
for st in 1 2 4; do
for rk in 2; do
GSR_RANK=$rk timeout 200 python - $st <<'PY' > gpurun_out/r2c6_bench_s$st.txt 2>&1
import os, sys
st = sys.argv[1]
sys.argv = ["bench.py", "--steps", "40", "--warmup", "4", "--no-train-step", "--no-cpu-baseline", "--no-reference-structure", "--streams", st]
from multiview_inpaint_b200 import _C
_C.debug_set(0, int(os.environ["GSR_RANK"]))
import bench
bench.run_ours(bench.parse())
PY
python - "$st" <<'PY'
import json, sys
try:
    lines = [l for l in open(f"gpurun_out/r2c6_bench_s{sys.argv[1]}.txt").read().strip().splitlines() if l.startswith("{")]
    d = json.loads(lines[-1])
    print("streams", sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
    print({k: v["ms_per_view"] for k, v in d["stages"].items()})
except Exception as ex:
    print("no bench line:", ex); print(open(f"gpurun_out/r2c6_bench_s{sys.argv[1]}.txt").read()[-1500:])
PY
done
done
