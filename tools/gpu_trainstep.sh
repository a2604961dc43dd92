#!/bin/bash
# GPU check of the training-step stages: parity tests, smoke, kernel timings vs torch structure
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_trainstep_gpu.py -m gpu -q > gpurun_out/pytest_trainstep.log 2>&1; echo "pytest trainstep rc=$?" >> gpurun_out/pytest_trainstep.log
tail -40 gpurun_out/pytest_trainstep.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 300 python tools/exp_trainstep.py > gpurun_out/trainstep.json 2> gpurun_out/trainstep.err; tail -5 gpurun_out/trainstep.err; cat gpurun_out/trainstep.json
