#!/bin/bash
# round 2, session 3, call 2 (one B200): blend kernels with the trimmed inner loops (BMSK bit clear, NaN-coordinate as the only
# `done` state, median depth by id, no warp vote in the tensor-core K7) -- full GPU suite, then the bench, plus K6 with the exact cull
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --maxfail=8 --tb=short > gpurun_out/s3c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3c2_pytest.log
tail -12 gpurun_out/s3c2_pytest.log | cut -c1-300
bash tools/gpu_variants.sh base fwdrefine
