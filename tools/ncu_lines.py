#!/usr/bin/env python3
"""Per-source-line totals from `ncu -i rep --page source --csv --print-source sass,cuda` (one section per file; SASS rows
carry the counters).  usage: ncu_lines.py export.csv [top_n]   -> file:line, warp instructions, thread instructions, stall samples"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0])
src_text = {}
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; i += 1; continue
    if r and r[0] == "Function Name":
        i += 1; continue
    if r and r[0] == "Line No":
        hdr = r; i += 1; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        line = r[0]
        def num(k):
            try: return int(float(d.get(k, "0").replace(",", "") or 0))
            except ValueError: return 0
        if r[1].strip():
            src_text[(cur_file, line)] = r[1].strip()[:90]
        a = agg[(cur_file, line)]
        a[0] += num("Instructions Executed"); a[1] += num("Thread Instructions Executed"); a[2] += num("# Samples")
    i += 1
tot = [sum(a[k] for a in agg.values()) for k in range(3)]
print(f"total warp inst {tot[0]:.3e}  thread inst {tot[1]:.3e}  samples {tot[2]}")
byfile = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), a in agg.items():
    for k in range(3): byfile[f][k] += a[k]
for f, a in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print(f"  {f:28s} inst {100*a[0]/max(tot[0],1):5.1f}%  samples {100*a[2]/max(tot[2],1):5.1f}%")
for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{f}:{l:>5s}  inst {100*a[0]/max(tot[0],1):5.2f}%  lanes {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/max(tot[2],1):5.2f}%  {src_text.get((f,l),'')}")
