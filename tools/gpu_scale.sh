#!/bin/bash
# bench.py at N ranks of one box (launched the way the driver does); appends the JSON line to gpurun_out/scale.jsonl
N=${1:-2}
mkdir -p gpurun_out
B="--no-cpu-baseline --no-reference-structure --no-train-step --steps 10 --warmup 3"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N $B > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
grep '^{' gpurun_out/scale_n$N.json >> gpurun_out/scale.jsonl
tail -3 gpurun_out/scale_n$N.err | cut -c1-300
python - <<PY
import json
for l in open("gpurun_out/scale_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l); print("N", d["n_gpus"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), d["config"]["allreduce"])
PY
