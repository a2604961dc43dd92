"""Instruction / stall-sample distribution of a kernel by 50-line SASS regions: python tools/ncu_regions.py rep regex"""
import csv, subprocess, io, collections, re, sys
out = subprocess.run(["ncu","-i",sys.argv[1],"--page","source","--csv","--kernel-name","regex:"+sys.argv[2]],capture_output=True,text=True)
rows=list(csv.reader(io.StringIO(out.stdout)))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"]
hdr=rows[hi[0]]; data=[r for r in rows[hi[0]+1:(hi[1]-1 if len(hi)>1 else len(rows))] if len(r)==len(hdr)]
ia,ie,isamp=hdr.index("Source"),hdr.index("Instructions Executed"),hdr.index("# Samples")
tot=sum(int(r[ie]) for r in data); stot=sum(int(r[isamp]) for r in data)
step=int(sys.argv[3]) if len(sys.argv)>3 else 50
for a in range(0,len(data),step):
    c=sum(int(r[ie]) for r in data[a:a+step]); s=sum(int(r[isamp]) for r in data[a:a+step])
    ops=collections.Counter(re.search(r'(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)',r[ia].strip()).group(1).split('.')[0] for r in data[a:a+step])
    print(a, f"inst {c/tot*100:5.1f}% smp {s/stot*100:5.1f}%  exec/line {c/step/1000:.0f}k", dict(ops.most_common(4)))
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={h:sum(int(r[hdr.index(h)]) for r in data) for h in stalls}
print(sorted(agg.items(), key=lambda x:-x[1])[:8])
