#!/bin/bash
# round 2, session 2, call 2 (TWO B200s): the world-2 GPU tests (NVLS vs NCCL values, fused step over two ranks), the exchange
# sweep at N = 2 (a dry run of the 8-GPU script), BASELINE config 3 (25-view SVD-XT orbit, strong scaling) and the headline at N = 2.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nvls_gpu.py tests/test_trainstep_world2_gpu.py -m gpu -q --tb=short > gpurun_out/s2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2c2_pytest.log
tail -6 gpurun_out/s2c2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29621 tools/exp_scale8.py 20 > gpurun_out/s2c2_scale2.jsonl 2> gpurun_out/s2c2_scale2.err; echo "scale2 rc=$?"; tail -3 gpurun_out/s2c2_scale2.err | cut -c1-300; tail -1 gpurun_out/s2c2_scale2.jsonl | cut -c1-600
B="--no-cpu-baseline --no-reference-structure --no-train-step --no-dropin --warmup 3"
timeout 500 $TR --master-port 29622 tools/exp_configs_multi.py "--workload svd_orbit --total-views 25 --steps 20 $B" "--workload headline --steps 30 $B" > gpurun_out/s2c2_configs_n2.jsonl 2> gpurun_out/s2c2_configs_n2.err; echo "configs rc=$?"; tail -3 gpurun_out/s2c2_configs_n2.err | cut -c1-300
python - <<'PY'
import json
for l in open("gpurun_out/s2c2_configs_n2.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); c = d["config"]
        print(c["workload"], "N", d["n_gpus"], d["scaling"], "views/step", c["views_per_step"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
              "ms/step", round(d["ms_per_step"], 3), c.get("allreduce"), c.get("view_balance"))
PY
