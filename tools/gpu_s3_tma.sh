#!/bin/bash
# round 2, session 3 (one B200): gsr_gather_rows with cp.async.bulk for the wide rows -- densification tests (both paths, bit-exact
# against torch indexing), timing at the headline model size, memcheck over the gather tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_densify_gpu.py -m gpu -q -x --tb=short -s 2>&1 | grep -E "gather_rows|passed|failed|Error|error" | head -20
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_densify_gpu.py -m gpu -q -k "gather_rows_equals and not 250001" 2>&1 | tail -3
