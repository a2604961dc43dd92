#!/bin/bash
# Full validation pass on one B200 (run under gpurun): gpu tests, smoke, both bench arms, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 240 python -m pytest tests/test_knn_gpu.py -m gpu -x -q > gpurun_out/pytest_knn.log 2>&1; echo "pytest knn rc=$?" >> gpurun_out/pytest_knn.log
tail -5 gpurun_out/pytest_knn.log
timeout 200 python tests/bench_knn.py > gpurun_out/knn_bench.log 2>&1; tail -4 gpurun_out/knn_bench.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_knn_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.json
timeout 300 python tools/exp_breakdown.py > gpurun_out/breakdown.log 2>&1; cat gpurun_out/breakdown.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'blend_|geom_backward|preprocess_kernel|rs_onesweep|bin_expand|tile_prepare|view_stats|rs_hist' -s 32 -c 32 -f -o gpurun_out/prof_full python tools/exp_view.py 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
