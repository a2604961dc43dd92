#!/bin/bash
# trainstep parity + whole GPU suite + bench (with the train_step leg)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainstep_gpu.py -m gpu -q -x > gpurun_out/pytest_trainstep.log 2>&1; echo "pytest trainstep rc=$?" >> gpurun_out/pytest_trainstep.log
tail -15 gpurun_out/pytest_trainstep.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_trainstep_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
