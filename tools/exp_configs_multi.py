"""Several bench.py configurations in ONE torchrun launch (one NCCL / python start-up for all of them): every argument is a
quoted bench.py command line; rank 0 prints one JSON line per configuration, exactly the line `bench.py <args>` prints.
    torchrun --nproc-per-node 8 tools/exp_configs_multi.py "--workload svd_orbit --total-views 25 --steps 20" "--mode infer ..."
Used for the BASELINE.json configs 3-5 at 2 / 4 / 8 GPUs (profiles/r02_configs_multigpu.json)."""
import gc
import os
import shlex
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

specs = sys.argv[1:]
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for spec in specs:
    sys.argv = ["bench.py", "--gpus", str(world)] + shlex.split(spec)
    args = bench.parse()
    try:
        (bench.run_infer if args.mode == "infer" else bench.run_ours)(args)
    except Exception:
        if int(os.environ.get("RANK", "0")) == 0:
            print("FAILED", spec, file=sys.stderr)
        traceback.print_exc()
    gc.collect()
    import torch
    torch.cuda.empty_cache()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
if world > 1:
    import torch.distributed as dist
    dist.destroy_process_group()
