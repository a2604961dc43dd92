"""Summarise an ncu report (.ncu-rep) into a markdown table for profiles/ (run here, no GPU needed).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_xxx.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("sm__inst_executed_pipe_fma.sum", "fma inst"),
    ("sm__inst_executed_pipe_xu.sum", "xu inst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("lts__t_sectors_op_red.sum", "L2 red sectors"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(m), label) for m, label in METRICS if m in hdr]
    groups = {}
    for r in rows[2:]:
        groups.setdefault(r[ki].split("(")[0].replace("void ", "").replace("gsr::", ""), []).append(r)
    print(f"Source: `{path}` (ncu --set full --clock-control none; cold-cache, serialised launches)\n")
    print("| kernel | n | " + " | ".join(f"{label} [{units[i]}]" if units[i] else label for i, label in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for name, rs in groups.items():
        vals = []
        for i, _ in cols:
            xs = []
            for r in rs:
                try:
                    xs.append(float(r[i].replace(",", "")))
                except ValueError:
                    pass
            vals.append(f"{sum(xs) / len(xs):.4g}" if xs else "-")
        print(f"| `{name}` | {len(rs)} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
