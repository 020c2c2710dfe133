#!/usr/bin/env python
"""Per-source-line instruction shares of an .ncu-rep:  python tools/ncu_lines.py rep.ncu-rep [top_n]
Prints the lines with most warp instructions (share, mean active lanes, stall-sample share) and per-file totals."""
import csv, io, subprocess, sys, collections
args = [a for a in sys.argv[1:] if not a.startswith("--")]; rep = args[0]; top = int(args[1]) if len(args) > 1 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
cols = rows[h]; ci = {}
for i, c in enumerate(cols): ci.setdefault(c, i)
cur, data = "", []
for r in rows[:h]:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]
for r in rows[h + 1:]:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < len(cols) or r[0] in ("", "Line No"): continue
    try: smp, ie, te = float(r[ci["# Samples"]]), float(r[ci["Instructions Executed"]]), float(r[ci["Thread Instructions Executed"]])
    except ValueError: continue
    data.append((cur, int(r[0]), ie, te, smp, r[1].strip()[:100]))
ti = sum(d[2] for d in data) or 1; ts = sum(d[4] for d in data) or 1
print(f"total warp instructions {ti/1e9:.2f} G, thread instructions {sum(d[3] for d in data)/1e9:.1f} G")
files = collections.defaultdict(lambda: [0, 0, 0])
for d in data:
    f = files[d[0]]; f[0] += d[2]; f[1] += d[3]; f[2] += d[4]
for k, f in sorted(files.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {f[0]/ti*100:5.1f}% inst  lanes={f[1]/max(f[0],1):5.1f}  {f[2]/ts*100:5.1f}% stall samples")
for d in sorted(data, key=lambda d: -d[2])[:top]:
    print(f"{d[2]/ti*100:5.1f}% inst lanes={d[3]/max(d[2],1):5.1f} {d[4]/ts*100:5.1f}% smp {d[0]}:{d[1]}  {d[5]}")
# optional: sums over named line ranges, e.g.  --ranges trt_isect.cuh:290-323=node,trt_isect.cuh:334-358=leafloop
for a in sys.argv:
    if a.startswith("--ranges="):
        print("ranges:")
        for spec in a[9:].split(","):
            loc, name = spec.split("="); f, r = loc.split(":"); lo, hi = (int(x) for x in r.split("-"))
            sel = [d for d in data if d[0] == f and lo <= d[1] <= hi]
            ie = sum(d[2] for d in sel); te = sum(d[3] for d in sel); sm = sum(d[4] for d in sel)
            print(f"  {name:24s} {ie/ti*100:5.1f}% inst ({ie/1e9:.2f} G) lanes={te/max(ie,1):5.1f}  {sm/ts*100:5.1f}% stall samples")
