#!/bin/bash
# round 2, session 3 (one B200): K6 at six CTAs per SM (A/B), then the final check: full suite, smoke, default bench line, inference config
mkdir -p gpurun_out
bash tools/gpu_variants.sh base fwd6
unset GSR_LIB_VARIANT
timeout 420 python -m pytest tests -m gpu -q --maxfail=8 --tb=short > gpurun_out/s3l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3l_pytest.log
tail -5 gpurun_out/s3l_pytest.log | cut -c1-300
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
B="--no-cpu-baseline --no-reference-structure --no-train-step --no-dropin --warmup 3"
timeout 300 python bench.py --mode infer --workload inference --views-per-rank 25 --orbit-deg 30 --steps 5 $B > gpurun_out/s3l_infer.json 2> gpurun_out/s3l_infer.err; echo "infer rc=$?"; tail -2 gpurun_out/s3l_infer.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/s3l_infer.json") if l.startswith("{")][-1])
    print("inference value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["roofline"]["view"])
except Exception as ex:
    print("no line", ex)
PY
