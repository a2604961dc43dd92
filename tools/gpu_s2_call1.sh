#!/bin/bash
# round 2, session 2, call 1 (one B200): GPU parity suite incl. the tensor-core K7 cross-check, smoke, the default bench line,
# the same bench with the shuffle-butterfly K7 (A/B), `ncu --set full` of one batched step, the launch list, memcheck.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --maxfail=8 --tb=short > gpurun_out/s2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2c1_pytest.log
tail -15 gpurun_out/s2c1_pytest.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 420 python bench.py > gpurun_out/s2c1_bench.json 2> gpurun_out/s2c1_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s2c1_bench.err
GSR_MMA=0 timeout 200 python - <<'PY' > gpurun_out/s2c1_bench_shuffle.txt 2>&1
import sys
sys.argv = ["bench.py", "--steps", "40", "--warmup", "4", "--no-train-step", "--no-cpu-baseline", "--no-reference-structure", "--no-dropin"]
from multiview_inpaint_b200 import _C
_C.debug_set(3, 0)
import bench
bench.run_ours(bench.parse())
PY
python - <<'PY'
import json
def show(path, tag):
    try:
        lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
        d = json.loads(lines[-1])
        print(tag, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"], "steps", d["steps"], "clocks", d["clocks"])
        print({k: (v["ms_per_view"], v["frac_hbm"]) for k, v in d["stages"].items()})
        for k in ("dropin", "reference_structure", "train_step", "parity_headline", "cpu_baseline"):
            if k in d: print(k, d[k])
    except Exception as ex:
        print(tag, "no bench line:", ex); print(open(path).read()[-1500:])
show("gpurun_out/s2c1_bench.json", "mma")
show("gpurun_out/s2c1_bench_shuffle.txt", "shuffle")
PY
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_full python tools/ncu_step.py headline 4 > gpurun_out/s2c1_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/s2c1_ncu_full.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py headline 4 > gpurun_out/s2c1_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py tests/test_batched_gpu.py -m gpu -q -x -k "config1 or tiny_capacity or dense_opaque or tensor_core or overflow" > gpurun_out/s2c1_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/s2c1_memcheck.log
ls -la gpurun_out | tail -12
