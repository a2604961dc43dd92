#!/bin/bash
# round-1 v6 evidence: ncu launch list of the bench command, ncu full capture of the densification gather, default bench line
mkdir -p gpurun_out
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-structure > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:gather_rows -c 2 -f -o gpurun_out/prof_densify python tools/exp_densify.py > gpurun_out/ncu_densify.log 2>&1
echo "ncu densify rc=$?"; tail -2 gpurun_out/ncu_densify.log | cut -c1-300
timeout 100 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "split", d["split"], "steps", d["steps"], "clocks", d["clocks"])
except Exception as ex:
    print("no bench line:", ex)
PY
ls -la gpurun_out | tail -8
