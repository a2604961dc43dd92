#!/bin/bash
# round 2, call 2: batched / compacted front end -- parity suite, then the bench with and without batching
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
tail -25 gpurun_out/r2c2_pytest.log
for mode in "" "--no-batched"; do
  tag=$( [ -z "$mode" ] && echo batched || echo perview )
  timeout 200 python bench.py --steps 50 --warmup 5 --no-train-step --no-cpu-baseline --no-reference-structure $mode > gpurun_out/r2c2_bench_$tag.json 2> gpurun_out/r2c2_bench_$tag.err; echo "bench $tag rc=$?"; tail -3 gpurun_out/r2c2_bench_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2c2_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
    print({k: v["ms_per_view"] for k, v in d["stages"].items()})
except Exception as ex:
    print("no bench line:", ex)
PY
done
