#!/bin/bash
# round 2, session 3, call 7 (one B200): expansion counts digits per CTA (mode 2) instead of per-instance tile_count atomics (mode 1)
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --maxfail=8 --tb=short > gpurun_out/s3c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3c7_pytest.log
tail -8 gpurun_out/s3c7_pytest.log | cut -c1-300
GSR_DEBUG_KNOBS="1=1" timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_batched_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -3
bash tools/gpu_knobs.sh "" "1=1"
GSR_BENCH_ARGS="--workload svd_orbit --total-views 25 --orbit-deg 30 --steps 20" bash tools/gpu_knobs.sh "" "1=1"
GSR_BENCH_ARGS="--workload stress --views-per-rank 1 --steps 3 --warmup 3" bash tools/gpu_knobs.sh "" "1=1"
