#!/bin/bash
# benches the product library under several gsr_debug_set settings in one GPU call: usage  bash tools/gpu_knobs.sh "" "2=5" "2=4,4=1" ...
mkdir -p gpurun_out
for k in "$@"; do
  export GSR_DEBUG_KNOBS=$k; [ -z "$k" ] && unset GSR_DEBUG_KNOBS
  tag=$(echo "k$k" | tr '=,' '__')
  timeout 200 python bench.py --steps 60 --warmup 4 --no-train-step --no-cpu-baseline --no-reference-structure --no-dropin ${GSR_BENCH_ARGS} > gpurun_out/knob_$tag.json 2> gpurun_out/knob_$tag.err || tail -3 gpurun_out/knob_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/knob_{sys.argv[1]}.json") if l.startswith("{")][-1])
    print(sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: v["ms_per_view"] for k, v in d["stages"].items()})
except Exception as ex:
    print(sys.argv[1], "no bench line:", ex)
PY
done
