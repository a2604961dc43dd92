#!/bin/bash
# round 2, session 3 (one B200): the evidence run of the round's final build -- full GPU suite, smoke, the default bench line
# (all legs), BASELINE configs 2-5 at N = 1, `ncu --set full` + launch list of one batched step, memcheck over the new paths.
T=${1:-s3f}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --maxfail=8 --tb=short > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 480 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
B="--no-cpu-baseline --no-reference-structure --no-train-step --no-dropin --warmup 3"
[ -z "$SKIP_CONFIGS" ] && timeout 600 python tools/exp_configs_multi.py "--workload mip360 --steps 50 $B" "--workload svd_orbit --total-views 25 --orbit-deg 30 --steps 20 $B" "--mode infer --workload inference --views-per-rank 25 --orbit-deg 30 --steps 5 $B" "--workload stress --views-per-rank 1 --steps 3 $B" > gpurun_out/${T}_configs_n1.jsonl 2> gpurun_out/${T}_configs_n1.err; echo "configs1 rc=$?"; tail -3 gpurun_out/${T}_configs_n1.err | cut -c1-300
python - "$T" <<'PY'
import json, sys
T = sys.argv[1]
def show(path, tag):
    try:
        lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
        for l in lines:
            d = json.loads(l); c = d["config"]
            print(tag, c["workload"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), "N", c.get("N"), "V", c.get("V"), "launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
            if "stages" in d: print("   ", {k: (v["ms_per_view"], v["frac_hbm"]) for k, v in d["stages"].items()})
            for k in ("roofline", "dropin", "train_step", "reference_structure", "parity_headline", "cpu_baseline"):
                if k in d: print("   ", k, str(d[k])[:600])
    except Exception as ex:
        print(tag, "no bench line:", ex); print(open(path).read()[-1500:])
show(f"gpurun_out/{T}_bench.json", "default")
show(f"gpurun_out/{T}_configs_n1.jsonl", "config")
PY
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_full python tools/ncu_step.py headline 4 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/${T}_ncu_full.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py headline 4 > gpurun_out/${T}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
if [ -n "$SKIP_SAN" ]; then ls -la gpurun_out | tail -4; exit 0; fi
SEL="config1 or tiny_capacity or dense_opaque or tensor_core or overflow or depth_ties or different_sizes or small_scenes or tight or clears"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py tests/test_batched_gpu.py tests/test_tight_binning_gpu.py -m gpu -q -k "$SEL" > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${T}_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py tests/test_batched_gpu.py tests/test_tight_binning_gpu.py -m gpu -q -k "config1 or tensor_core or different_sizes or tight_lists or clears" > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${T}_racecheck.log
ls -la gpurun_out | tail -8
