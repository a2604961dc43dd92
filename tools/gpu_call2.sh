#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_knn_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tests/bench_knn.py 2>&1 | tail -4
timeout 300 python -m pytest tests/test_api_gpu.py -m gpu -x -q 2>&1 | tail -5
bash tools/gpu_configs.sh
