#!/bin/bash
# quick GPU check: parity tests + stage breakdown of the headline view + kernel timeline of the step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/exp_breakdown.py 2>&1 | tail -12
timeout 300 python tools/exp_trace.py 2 2>&1 | tail -3
timeout 300 python tools/exp_trace.py 4 2>&1 | tail -3
