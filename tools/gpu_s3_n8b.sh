#!/bin/bash
# round 2, session 3 (EIGHT B200s): headline at 8 ranks with the tapered exchange candidates in the step autotune
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--no-cpu-baseline --no-reference-structure --no-train-step --no-dropin --warmup 3"
timeout 300 $TR --nproc-per-node 8 --master-port 29663 bench.py --gpus 8 --steps 60 $B > gpurun_out/s3n8b_headline.json 2> gpurun_out/s3n8b_headline.err; echo "rc=$?"; tail -2 gpurun_out/s3n8b_headline.err | cut -c1-300
python - <<'PY'
import json
for l in open("gpurun_out/s3n8b_headline.json"):
    if l.startswith("{"):
        d = json.loads(l); c = d["config"]
        print(c["workload"], "N", d["n_gpus"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), c.get("allreduce"))
PY
