"""Hot SASS lines of one kernel from an ncu report: python tools/ncu_source.py rep.ncu-rep <kernel regex> [min %]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hdr_i[0]; end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
hdr = rows[start]; data = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ie]) for r in data); stot = sum(int(r[isamp]) for r in data)
print("kernel", rows[start - 1][1][:80] if start else "", "| total warp inst", tot, "samples", stot, "SASS lines", len(data))
for k, r in enumerate(data):
    c, s = int(r[ie]), int(r[isamp])
    if c > tot * thr / 100 or s > stot * thr / 100:
        print(f"{k:4d} inst {c / tot * 100:5.2f}%  smp {s / max(stot,1) * 100:5.2f}%  {r[ia][:100]}")
