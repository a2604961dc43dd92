#!/bin/bash
# round-1 closing check on one B200: full GPU parity suite (incl. the densification re-pack), smoke, the default bench line,
# the densification timing.  Everything under its own timeout; logs in gpurun_out/.
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q --maxfail=6 --tb=short > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -25 gpurun_out/pytest_final.log
timeout 90 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"], "clocks", d["clocks"], "launches", d["gpu_launches"])
except Exception as ex:
    print("no bench line:", ex)
PY
timeout 60 python tools/exp_densify.py > gpurun_out/densify.json 2> gpurun_out/densify.err; echo "densify rc=$?"; tail -3 gpurun_out/densify.err; cat gpurun_out/densify.json
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
