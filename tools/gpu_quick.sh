#!/bin/bash
# quick GPU check: parity tests + stage breakdown of the headline view
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/exp_breakdown.py 2>&1 | tail -12
