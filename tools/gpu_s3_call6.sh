#!/bin/bash
# round 2, session 3, call 6 (one B200): accumulator clear moved into K1; what the tile_count atomics of the expansion cost
# (knob 1 = 0: WRONG results, timing only); two view groups on two streams
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --maxfail=8 --tb=short > gpurun_out/s3c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3c6_pytest.log
tail -8 gpurun_out/s3c6_pytest.log | cut -c1-300
bash tools/gpu_knobs.sh "" "1=0"
GSR_BENCH_ARGS="--streams 2" bash tools/gpu_knobs.sh ""
