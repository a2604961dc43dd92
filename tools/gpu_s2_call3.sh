#!/bin/bash
# round 2, session 2, call 3 (EIGHT B200s): NVLS vs NCCL values at 8 ranks, the exchange sweep at N = 8, every BASELINE.json
# config at its stated GPU counts (3: 25-view SVD-XT orbit strong-scaled over 4 and 8; 4: 200-view 1080p inference over 8;
# 5: stress at 8) and the headline at 4 and 8 with the step-level exchange autotune.
N=${1:-8}
mkdir -p gpurun_out
GSR_TEST_WORLD=$N timeout 200 python -m pytest tests/test_nvls_gpu.py -m gpu -q --tb=short 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node $N --master-port 29631 tools/exp_scale8.py 30 > gpurun_out/s2c3_scale$N.jsonl 2> gpurun_out/s2c3_scale$N.err; echo "scale8 rc=$?"; tail -2 gpurun_out/s2c3_scale$N.err | cut -c1-300
B="--no-cpu-baseline --no-reference-structure --no-train-step --no-dropin --warmup 3"
timeout 600 $TR --nproc-per-node $N --master-port 29632 tools/exp_configs_multi.py "--workload headline --steps 50 $B" "--workload svd_orbit --total-views 25 --orbit-deg 30 --steps 20 $B" "--mode infer --workload inference --views-per-rank 25 --orbit-deg 30 --steps 5 $B" "--workload stress --views-per-rank 1 --steps 3 $B" > gpurun_out/s2c3_configs_n$N.jsonl 2> gpurun_out/s2c3_configs_n$N.err; echo "configs8 rc=$?"; tail -3 gpurun_out/s2c3_configs_n$N.err | cut -c1-300
[ "$N" -ge 8 ] && timeout 400 $TR --nproc-per-node 4 --master-port 29633 tools/exp_configs_multi.py "--workload headline --steps 50 $B" "--workload svd_orbit --total-views 25 --orbit-deg 30 --steps 20 $B" > gpurun_out/s2c3_configs_n4.jsonl 2> gpurun_out/s2c3_configs_n4.err; echo "configs4 rc=$?"; tail -3 gpurun_out/s2c3_configs_n4.err | cut -c1-300
python - <<'PY'
import json
import glob
for fn in sorted(glob.glob("gpurun_out/s2c3_configs_n*.jsonl")):
  if True:
    for l in open(fn):
        if l.startswith("{"):
            d = json.loads(l); c = d["config"]
            print(c["workload"], "N", d["n_gpus"], d["scaling"], "views/step", c.get("views_per_step"), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
                  "ms/step", round(d["ms_per_step"], 3), {k: v for k, v in (c.get("allreduce") or {}).items() if k != "step_trials"}, c.get("view_balance"))
for l in [x for f in glob.glob("gpurun_out/s2c3_scale*.jsonl") for x in open(f)]:
    if l.startswith('{"world"'):
        d = json.loads(l); print("no exchange", d["no_exchange_ms"], "best", d["best"])
PY
