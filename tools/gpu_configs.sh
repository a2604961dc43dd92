#!/bin/bash
# BASELINE.json configs 2-5 on one B200 (run under gpurun): one bench line each into gpurun_out/configs.jsonl
mkdir -p gpurun_out
: > gpurun_out/configs.jsonl
B="--no-cpu-baseline --no-reference-structure --no-train-step --steps 5 --warmup 3"
timeout 300 python bench.py --workload mip360 --views-per-rank 4 $B >> gpurun_out/configs.jsonl 2> gpurun_out/cfg_mip360.err
timeout 300 python bench.py --workload svd_orbit --views-per-rank 25 --orbit-deg 30 $B >> gpurun_out/configs.jsonl 2> gpurun_out/cfg_svd.err
timeout 400 python bench.py --workload inference --mode infer --views-per-rank 8 --orbit-deg 30 $B >> gpurun_out/configs.jsonl 2> gpurun_out/cfg_infer.err
timeout 300 python bench.py --workload headline --mode infer --views-per-rank 8 $B >> gpurun_out/configs.jsonl 2> gpurun_out/cfg_infer_headline.err
timeout 600 python bench.py --workload stress --views-per-rank 2 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-structure --no-train-step >> gpurun_out/configs.jsonl 2> gpurun_out/cfg_stress.err
wc -l gpurun_out/configs.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/configs.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); c = d["config"]
        print(c["workload"], c.get("mode", "train"), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "N", c["N"], "V", c["V"],
              "dom", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "view frac", round(d["roofline"]["view"]["frac"], 3))
PY
tail -3 gpurun_out/cfg_*.err
