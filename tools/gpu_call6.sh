#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_imageio_gpu.py tests/test_trainstep_world2_gpu.py tests/test_nvls_gpu.py -m gpu -q > gpurun_out/pytest_call6.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_call6.log
tail -40 gpurun_out/pytest_call6.log
timeout 400 python bench.py --workload inference --mode infer --views-per-rank 8 --orbit-deg 30 --steps 4 --warmup 3 --files 8 > gpurun_out/infer_files.json 2> gpurun_out/infer_files.err
tail -3 gpurun_out/infer_files.err; python - <<'PY'
import json
for l in open("gpurun_out/infer_files.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["value"], d["e2e"], d.get("to_disk"))
PY
