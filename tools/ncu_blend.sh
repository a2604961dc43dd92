#!/bin/bash
# one `ncu --set full` capture of the blend kernels + a launch list (run under gpurun)
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'blend_|geom_backward|preprocess_kernel|rs_onesweep' -s 12 -c 12 -f -o gpurun_out/prof_blend python tools/exp_view.py 3 > gpurun_out/ncu_blend.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches.csv python tools/exp_view.py 4 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_blend.log
