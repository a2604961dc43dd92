mkdir -p gpurun_out
B="--no-cpu-baseline --no-reference-structure --no-train-step --steps 10 --warmup 3"
for AR in nvls nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 $B --allreduce $AR > gpurun_out/n2_$AR.json 2> gpurun_out/n2_$AR.err
python - <<PY
import json
for l in open("gpurun_out/n2_$AR.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$AR", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), d["config"]["allreduce"])
PY
done
