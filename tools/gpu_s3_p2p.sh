#!/bin/bash
# round 2, session 3 (TWO B200s): the peer-to-peer two-shot exchange -- value test, stress under skew, headline at N = 2 with the
# p2p candidates in the step autotune
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_nvls_gpu.py -m gpu -q --tb=short 2>&1 | tail -4 | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29671 tools/exp_nvls_stress.py 100 p2p 2>&1 | grep "^P=" | cut -c1-200
B="--no-cpu-baseline --no-reference-structure --no-train-step --no-dropin --warmup 3"
timeout 300 $TR --master-port 29672 bench.py --gpus 2 --steps 60 $B > gpurun_out/s3p2p_headline.json 2> gpurun_out/s3p2p_headline.err; echo "rc=$?"; tail -2 gpurun_out/s3p2p_headline.err | cut -c1-300
python - <<'PY'
import json
for l in open("gpurun_out/s3p2p_headline.json"):
    if l.startswith("{"):
        d = json.loads(l); c = d["config"]
        print(c["workload"], "N", d["n_gpus"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), c.get("allreduce"))
PY
