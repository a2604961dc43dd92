"""Where a multi-rank step spends its time (run under torchrun): per rank, CUDA-event times of
  views   = forward + blend backward of this rank's views (two streams)
  geom    = batched per-Gaussian backward
  reduce  = arena all-reduce (NVLS kernel or NCCL), sequential after geom
  chunked = geom + reduce pipelined over Gaussian-range chunks
and the max / min over ranks of each (imbalance).  usage: torchrun ... tools/exp_scale.py [chunks]"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_inpaint_b200 import _C, multiview as mv, scenes as S
from multiview_inpaint_b200.rasterizer import GaussianRasterizationSettings

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
name = "headline"
cfg = S.CONFIGS[name]; sc = S.make_config_scene(name)
P, W, H, M, D = sc["P"], sc["W"], sc["H"], sc["M"], sc["sh_degree"]
gauss = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
vpr = 4; n_views = vpr * world
cams = S.orbit_cameras(n_views, W, H, max_deg=5.0)
mine = mv.shard_views(n_views, rank, world)
bg = torch.zeros(3, device=dev)
cams_dev = {v: cams[v].to(dev) for v in mine}
wts = {v: S.loss_weights(W, H, cfg["seed"] + v).to(dev) for v in mine}
def settings(c):
    return GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg, scale_modifier=1.0,
                                         viewmatrix=c.world_view_transform, projmatrix=c.full_proj_transform, sh_degree=D,
                                         campos=c.camera_center, prefiltered=False)
arena = mv.GradArena(P, M, dev, symmetric=True)
av = mv.AsyncViews(n_views)
for v in mine:
    r = mv.cuda_view_fwd_bwd(gauss, settings(cams_dev[v]), lambda c, v=v: wts[v], arena, capacity=0)
    av.learn(v, r.num_rendered)
Ns = [int(av.slots[v, 0]) for v in mine]
pipe = mv.ViewPipeline(dev, 2)
wss = [_C.Workspace(dev) for _ in mine]

def views():
    states = []
    with pipe.step():
        for k, v in enumerate(mine):
            states.append(mv.cuda_view_fwd_blend(gauss, settings(cams_dev[v]), lambda c, v=v: wts[v], capacity=av.capacity(v),
                                                 async_result=av.slot(v), pipeline=pipe, workspace=wss[k]))
    return states

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def measure(method, chunked, iters=6):
    arena.method = method
    acc = {"views": 0.0, "geom": 0.0, "reduce": 0.0, "step": 0.0}
    for it in range(iters + 2):
        torch.cuda.synchronize(); dist.barrier()
        e0 = ev(); st = views(); e1 = ev()
        if chunked:
            mv.cuda_views_geom_backward_allreduce(gauss, st, arena, chunks=chunks); e2 = e3 = ev()
        else:
            mv.cuda_views_geom_backward(gauss, st, arena); e2 = ev()
            arena.all_reduce(); e3 = ev()
        torch.cuda.synchronize()
        if it >= 2:
            acc["views"] += e0.elapsed_time(e1); acc["geom"] += e1.elapsed_time(e2); acc["reduce"] += e2.elapsed_time(e3); acc["step"] += e0.elapsed_time(e3)
    t = torch.tensor([acc[k] / iters for k in ("views", "geom", "reduce", "step")], device=dev)
    mx, mn = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    if rank == 0:
        lab = f"{method:5s} {'chunked x%d' % chunks if chunked else 'sequential'}"
        print(f"{lab:22s} views {mn[0]:.2f}..{mx[0]:.2f}  geom(+reduce if chunked) {mn[1]:.2f}..{mx[1]:.2f}  reduce {mn[2]:.2f}..{mx[2]:.2f}  step {mn[3]:.2f}..{mx[3]:.2f} ms", flush=True)

def chunk_loop(st, do_geom, do_reduce, same_stream=False):
    views_ = [dict(radii=s.radii, geom=s.geom, scratch=s.scratch, viewmatrix=s.settings.viewmatrix, projmatrix=s.settings.projmatrix,
                   campos=s.settings.campos, tanfovx=s.settings.tanfovx, tanfovy=s.settings.tanfovy, width=W, height=H) for s in st]
    main, comm = torch.cuda.current_stream(dev), arena.comm_stream()
    if same_stream:
        comm = main
    step = ((P + chunks - 1) // chunks + 31) // 32 * 32
    for g0 in range(0, P, step):
        g1 = min(P, g0 + step)
        if do_geom:
            _C.backward_geom_multi(gauss["means3D"], gauss["shs"], gauss["scales"], gauss["rotations"], 1.0, D, views_, arena.views,
                                   stats=(arena.grad_norm_accum, arena.visible_count, arena.max_radii), g_range=(g0, g1))
        if do_reduce:
            e = torch.cuda.Event(); e.record(main); comm.wait_event(e)
            with torch.cuda.stream(comm):
                arena.all_reduce_range(g0, g1, post_barrier=False)
    if do_reduce:
        with torch.cuda.stream(comm):
            arena.barrier()
        main.wait_stream(comm)

def measure_loop(label, **kw):
    arena.method = "nvls"
    tot = 0.0
    for it in range(8):
        torch.cuda.synchronize(); dist.barrier()
        st = views(); e1 = ev(); chunk_loop(st, **kw); e2 = ev(); torch.cuda.synchronize()
        if it >= 2: tot += e1.elapsed_time(e2)
    t = torch.tensor([tot / 6], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"chunk loop x{chunks} {label:34s} {t.item():.3f} ms", flush=True)

if rank == 0:
    print(f"world {world}, nvls mapping: {bool(arena._mc)}, arena {arena.storage.numel() * 4 / 1e6:.0f} MB", flush=True)
allN = [None] * world
dist.all_gather_object(allN, Ns)
if rank == 0:
    print("N per rank (sum over its views):", [sum(x) for x in allN], flush=True)
for method in (["nvls"] if arena._mc else []) + ["nccl"]:
    measure(method, False)
    if method == "nvls":
        measure(method, True)
if arena._mc:
    measure_loop("geom only", do_geom=True, do_reduce=False)
    measure_loop("reduce only", do_geom=False, do_reduce=True)
    measure_loop("geom + reduce, two streams", do_geom=True, do_reduce=True)
    measure_loop("geom + reduce, one stream", do_geom=True, do_reduce=True, same_stream=True)
dist.destroy_process_group()
