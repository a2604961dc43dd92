#!/bin/bash
# round-1 v5 evidence: trainstep parity + timings, ncu launch list of the bench, ncu full capture of the new kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainstep_gpu.py -m gpu -q -x > gpurun_out/pytest_trainstep.log 2>&1; echo "pytest trainstep rc=$?" >> gpurun_out/pytest_trainstep.log
tail -12 gpurun_out/pytest_trainstep.log
timeout 300 python tools/exp_trainstep.py > gpurun_out/trainstep.json 2> gpurun_out/trainstep.err; tail -3 gpurun_out/trainstep.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/trainstep.json"))
for r in d["rows"]:
    print(r["kernel"][:60], r["ms"], r["frac"], r.get("torch_structure_ms"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-structure > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ssim_l1|loss_finalize|adam_kernel|activate_' -c 12 -f -o gpurun_out/prof_trainstep python tools/exp_trainstep_once.py > gpurun_out/ncu_trainstep.log 2>&1
tail -3 gpurun_out/ncu_trainstep.log
ls -la gpurun_out
