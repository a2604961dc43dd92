#!/bin/bash
mkdir -p gpurun_out
python -c "
import sys; sys.path.insert(0,'.')
from multiview_inpaint_b200 import _C
" 
GSR_RANK=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"preprocess|rs_onesweep|rs_hist|bin_expand_kernel" -s 45 -c 9 -f -o gpurun_out/r2c5_front python tools/exp_front.py 3 > gpurun_out/r2c5_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2c5_ncu.log
ls -la gpurun_out/r2c5_front.ncu-rep
