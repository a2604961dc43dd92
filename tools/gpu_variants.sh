#!/bin/bash
# benches several library variants (tools/build_variant.py) in one GPU call: usage  bash tools/gpu_variants.sh name1 name2 ...
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "base" ]; then unset GSR_LIB_VARIANT; else export GSR_LIB_VARIANT=$v; fi
  timeout 200 python -m pytest tests/test_parity_gpu.py tests/test_batched_gpu.py -m gpu -q -x --tb=line 2>&1 | tail -1
  timeout 200 python bench.py --steps 60 --warmup 4 --no-train-step --no-cpu-baseline --no-reference-structure --no-dropin > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err || tail -3 gpurun_out/variant_$v.err
  python - "$v" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/variant_{sys.argv[1]}.json") if l.startswith("{")][-1])
    print(sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: v["ms_per_view"] for k, v in d["stages"].items()})
except Exception as ex:
    print(sys.argv[1], "no bench line:", ex)
PY
done
