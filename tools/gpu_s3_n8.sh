#!/bin/bash
# round 2, session 3 (EIGHT B200s): the final kernels at 8 ranks -- headline (weak, 4 views per rank), SVD orbit (strong, 25 views), inference
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--no-cpu-baseline --no-reference-structure --no-train-step --no-dropin --warmup 3"
timeout 500 $TR --nproc-per-node 8 --master-port 29662 tools/exp_configs_multi.py "--workload headline --steps 50 $B" "--workload svd_orbit --total-views 25 --orbit-deg 30 --steps 20 $B" "--mode infer --workload inference --views-per-rank 25 --orbit-deg 30 --steps 5 $B" > gpurun_out/s3n8_configs.jsonl 2> gpurun_out/s3n8_configs.err; echo "configs8 rc=$?"; tail -3 gpurun_out/s3n8_configs.err | cut -c1-300
python - <<'PY'
import json
for l in open("gpurun_out/s3n8_configs.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); c = d["config"]
        print(c["workload"], "N", d["n_gpus"], d["scaling"], "views/step", c.get("views_per_step"), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
              "ms/step", round(d["ms_per_step"], 3), {k: v for k, v in (c.get("allreduce") or {}).items()}, c.get("view_balance"))
PY
