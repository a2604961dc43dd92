"""N-rank check + timing of the in-switch arena all-reduce against NCCL (run under torchrun)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import multiview as mv
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
P, M = (int(sys.argv[1]) if len(sys.argv) > 1 else 3000000), 16
a = mv.GradArena(P, M, dev, symmetric=True)
b = mv.GradArena(P, M, dev, symmetric=False)
if rank == 0: print("nvls:", a.uses_nvls, getattr(a, "symmetric_error", ""), "arena MB", a.storage.numel() * 4 / 1e6, flush=True)
gs = torch.Generator(device=dev); gs.manual_seed(99)   # same on all ranks: 35 % of the Gaussians are seen by nobody
never = torch.rand(P, device=dev, generator=gs) < 0.35
def fill(x):
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    x.flat.copy_(torch.randn(x.flat.shape, device=dev, generator=g))
    x.grad_norm_accum.copy_(torch.rand(P, device=dev, generator=g))
    x.visible_count.copy_(torch.randint(0, 5, (P,), device=dev, generator=g, dtype=torch.int32))
    x.max_radii.copy_(torch.randint(0, 500, (P,), device=dev, generator=g, dtype=torch.int32))
    x.visible_count[never] = 0
    x.views["dL_dsh"][x.visible_count == 0] = 0   # the invariant the sparse path relies on: unseen => zero row
fill(a); fill(b)
a.all_reduce(); b.all_reduce(); torch.cuda.synchronize()
err = (a.flat - b.flat).abs().max().item(); ref = b.flat.abs().max().item()
ok = err <= 1e-5 * ref and torch.equal(a.visible_count, b.visible_count) and torch.equal(a.max_radii, b.max_radii) \
    and (a.grad_norm_accum - b.grad_norm_accum).abs().max().item() <= 1e-5
print(f"rank {rank}: max |nvls - nccl| = {err:.3e} (scale {ref:.2f}) stats equal -> {'OK' if ok else 'MISMATCH'}", flush=True)
def timeit(x, label):
    for _ in range(3): x.all_reduce()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): x.all_reduce()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        nbytes = x.storage.numel() * 4
        print(f"{label:28s} {ms.item():7.3f} ms   algbw {nbytes / ms.item() / 1e6:7.1f} GB/s  busbw {nbytes / ms.item() / 1e6 * 2 * (world - 1) / world:7.1f} GB/s", flush=True)
timeit(b, "NCCL (3 all_reduce calls)")
timeit(a, "NVLS sparse rows + 2 barriers")
a.sparse = False
timeit(a, "NVLS dense + 2 barriers")
a.sparse = True
from multiview_inpaint_b200 import _C
if a.uses_nvls:
    for sparse_rows, blocks in [(P, b) for b in (148, 296, 592, 1184)] + [(0, b) for b in (148, 296, 592, 1184)]:
        fill(a)
        def f():
            a._handle.barrier(); _C.nvls_all_reduce(a._mc, dev, 0, a._n_f32, a._off_cnt, P, a._off_max, P, rank, world, blocks, a._sh_first, sparse_rows, 48); a._handle.barrier()
        class X: storage = a.storage; all_reduce = staticmethod(f)
        timeit(X, f"NVLS {'sparse' if sparse_rows else 'dense '} blocks={blocks}")
dist.destroy_process_group()
