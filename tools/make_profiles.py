"""Turns the scratch outputs of tools/gpu_round.sh (gpurun_out/) into the tracked evidence under profiles/:
   python tools/make_profiles.py <tag>      e.g. r01_v2
 - profiles/<tag>_launches.csv      the ncu launch list of `bench.py --steps 2 --warmup 1` (gpu__time_duration per launch)
 - profiles/<tag>_launch_shares.md  per-kernel share of the step from that list
 - profiles/<tag>_ncu_full.md       one row per kernel from the `ncu --set full` capture
 - profiles/traffic.json            dram__bytes_read.sum + dram__bytes_write.sum per launch, per stage (bench.py reads it)
 - profiles/<tag>_bench.json        the bench lines of that run (ours + reference arm)"""
import collections, csv, io, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
STAGE_OF = {"preprocess_kernel": "preprocess", "bin_expand_kernel": "duplicate", "bin_expand_big_kernel": "duplicate",
            "tile_prepare_kernel": "tile_ranges", "tile_ranges_views_kernel": "tile_ranges", "blend_forward_kernel": "blend_forward", "blend_backward_kernel": "blend_backward",
            "geom_backward_kernel": "geom_backward", "view_stats_kernel": "view_stats", "scan_kernel": "scan"}

def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("gsr::", "")

# ---- launch list ----
src = os.path.join(G, "launches.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(PR, f"{tag}_launches.csv"))
    lines = [l for l in open(src) if l.startswith('"')]
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for row in csv.DictReader(lines):
        n = short(row["Kernel Name"]); t = float(row["Metric Value"]) / 1000
        tot[n] = tot.get(n, 0) + t; cnt[n] += 1
    T = sum(tot.values())
    with open(os.path.join(PR, f"{tag}_launch_shares.md"), "w") as f:
        f.write(f"Source: `profiles/{tag}_launches.csv` (`ncu --metrics gpu__time_duration.sum --clock-control none` around "
                f"`python bench.py --steps 2 --warmup 1 --no-cpu-baseline`; cold-cache, serialised launches: compare SHARES)\n\n")
        f.write("| kernel | launches | total us | share % | us / launch |\n|---|---|---|---|---|\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"| `{k[:90]}` | {cnt[k]} | {v:.1f} | {100 * v / T:.1f} | {v / cnt[k]:.1f} |\n")
        f.write(f"\nTotal {T:.1f} us over {sum(cnt.values())} launches.\n")

# ---- full capture ----
rep = os.path.join(G, "prof_full.ncu-rep")
if os.path.exists(rep):
    md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(PR, f"{tag}_ncu_full.md"), "w").write(md.replace(G, "gpurun_out"))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = collections.defaultdict(list)
    for r in rows[2:]:
        n = short(r[ki]).split("<")[0]
        b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
        acc[n].append(b)
    # the capture is ONE batched step of `views` views (tools/ncu_step.py): a stage's traffic PER VIEW is the sum over its
    # launches divided by the views.  The onesweep launches split into the depth sort (the first 4: 32 key bits) and the
    # tile sort (the rest) by order.
    views = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    STAGE_OF.update({"geom_backward_multi_kernel": "geom_backward"})   # zero_regions_kernel: counters / look-back words only
                                                                        # (K1 clears the accumulators since r02_v4)
    traffic, seen_sweeps = {}, 0
    for r in rows[2:]:
        n = short(r[ki]).split("<")[0]
        b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
        if n == "rs_onesweep_kernel":
            st = "depth_sort" if seen_sweeps < 4 else "tile_sort"
            seen_sweeps += 1
        else:
            st = STAGE_OF.get(n)
        if st:
            traffic[st] = traffic.get(st, 0) + b / views
    tp = os.path.join(PR, "traffic.json")
    old = json.load(open(tp)) if os.path.exists(tp) else {}
    old["headline"] = {k: int(v) for k, v in traffic.items()}
    old["_source"] = (f"profiles/{tag}_ncu_full.md (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full of one batched "
                      f"{views}-view step, per view)")
    json.dump(old, open(tp, "w"), indent=1)
    # issue-slot utilisation and warp instructions of the issue-bound blend kernels (bench.py: roofline.issue_active_frac)
    def col(name):
        return hdr.index(name) if name in hdr else None
    ii, wi_, ti = col("smsp__issue_active.avg.pct_of_peak_sustained_active"), col("smsp__inst_executed.sum"), col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    ctr = {}
    for r in rows[2:]:
        n = short(r[ki]).split("<")[0]
        st = STAGE_OF.get(n)
        if st in ("blend_forward", "blend_backward") and ii is not None:
            ctr[st] = {"issue_active_frac": round(float(r[ii].replace(",", "")) / 100.0, 4),
                       "warp_inst": int(float(r[wi_].replace(",", ""))) if wi_ is not None else None,
                       "tensor_pipe_active_frac": (round(float(r[ti].replace(",", "")) / 100.0, 4) if ti is not None else None),
                       "views_per_launch": views, "source": f"profiles/{tag}_ncu_full.md"}
    if ctr:
        json.dump(ctr, open(os.path.join(PR, "roofline_counters.json"), "w"), indent=1)

# ---- bench lines ----
out = {}
for k, fn in (("ours", "bench_n1.json"), ("reference", "bench_ref.json")):
    p = os.path.join(G, fn)
    if os.path.exists(p):
        for l in open(p):
            if l.startswith("{"):
                out[k] = json.loads(l)
if out:
    json.dump(out, open(os.path.join(PR, f"{tag}_bench.json"), "w"), indent=1)
print("wrote", sorted(os.listdir(PR)))
