#!/bin/bash
# round 2, call 1: the previously gated GPU tests (depth ties, permutation invariance, create_from_pcd, PLY), the
# CUDA-graph experiment, and a baseline bench line of the round-1 kernels on this round's box.
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q --maxfail=10 --tb=short > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
tail -15 gpurun_out/r2c1_pytest.log
timeout 120 python tools/exp_graph.py 30 > gpurun_out/r2c1_graph.json 2> gpurun_out/r2c1_graph.err; echo "graph rc=$?"; tail -3 gpurun_out/r2c1_graph.err; cat gpurun_out/r2c1_graph.json
timeout 150 python bench.py --steps 50 --warmup 5 --no-train-step > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c1_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c1_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"], "clocks", d["clocks"], "launches", d["gpu_launches"])
    print({k: v["ms_per_view"] for k, v in d["stages"].items()})
except Exception as ex:
    print("no bench line:", ex)
PY
