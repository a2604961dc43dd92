#!/bin/bash
# full 1-GPU check: GPU test suite, smoke, the default bench line with every leg
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --maxfail=5 --tb=short > gpurun_out/r2c8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c8_pytest.log
tail -6 gpurun_out/r2c8_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c8_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c8_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"], "steps", d["steps"], "clocks", d["clocks"])
    print({k: (v["ms_per_view"], v["frac_hbm"]) for k, v in d["stages"].items()})
    print("roofline", {k: d["roofline"][k] for k in ("bound", "kernel", "achieved", "frac", "traffic", "issue_active_frac")})
    print("dropin", d.get("dropin")); print("refstruct", d.get("reference_structure")); print("train", d.get("train_step"))
    print("parity", d.get("parity_headline")); print("cpu", d.get("cpu_baseline"))
except Exception as ex:
    print("no bench line:", ex)
PY
