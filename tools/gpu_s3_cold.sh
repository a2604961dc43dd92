#!/bin/bash
# cold-start repetitions of the world-2 fused step (one process pair per repetition)
for i in $(seq 1 ${1:-24}); do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + i)) tools/exp_w2_loop.py 1 2>&1 | grep "^reps" | cut -c1-200
done
