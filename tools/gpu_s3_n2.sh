#!/bin/bash
# round 2, session 3 (TWO B200s): world-2 GPU tests on the final kernels, headline (weak) and SVD orbit (strong, 25 views) at N = 2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nvls_gpu.py tests/test_trainstep_world2_gpu.py -m gpu -q --tb=short > gpurun_out/s3n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3n2_pytest.log
tail -5 gpurun_out/s3n2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29641 bench.py --gpus 2 > gpurun_out/s3n2_bench_default.json 2> gpurun_out/s3n2_bench_default.err; echo "driver-style bench rc=$?"; tail -2 gpurun_out/s3n2_bench_default.err | cut -c1-300
timeout 200 $TR --master-port 29643 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/s3n2_bench_ref.json 2> gpurun_out/s3n2_bench_ref.err; echo "reference arm rc=$?"; tail -c 600 gpurun_out/s3n2_bench_ref.json
B="--no-cpu-baseline --no-reference-structure --no-train-step --no-dropin --warmup 3"
timeout 500 $TR --master-port 29642 tools/exp_configs_multi.py "--workload svd_orbit --total-views 25 --orbit-deg 30 --steps 20 $B" "--mode infer --workload inference --views-per-rank 25 --orbit-deg 30 --steps 5 $B" > gpurun_out/s3n2_configs.jsonl 2> gpurun_out/s3n2_configs.err; echo "configs rc=$?"; tail -3 gpurun_out/s3n2_configs.err | cut -c1-300
python - <<'PY'
import json
for fn in ("gpurun_out/s3n2_bench_default.json", "gpurun_out/s3n2_configs.jsonl"):
    for l in open(fn):
        if l.startswith("{"):
            d = json.loads(l); c = d["config"]
            print(c["workload"], "N", d["n_gpus"], d["scaling"], "views/step", c.get("views_per_step"), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
                  "ms/step", round(d["ms_per_step"], 3), {k: v for k, v in (c.get("allreduce") or {}).items()}, c.get("view_balance"))
PY
