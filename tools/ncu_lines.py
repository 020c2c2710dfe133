#!/usr/bin/env python
"""Summarises `ncu --page source --print-source cuda,sass --csv` per CUDA source line:
share of warp-stall samples, share of executed warp instructions, mean active lanes."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
cols = rows[hdr]
ci = {}
for i, c in enumerate(cols):
    ci.setdefault(c, i)
cur_file = ""
data = []
for r in rows[:hdr]:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
for r in rows[hdr + 1:]:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) < len(cols) or r[0] in ("", "Line No"):
        continue
    try:
        smp = float(r[ci["# Samples"]]); ie = float(r[ci["Instructions Executed"]]); te = float(r[ci["Thread Instructions Executed"]])
    except ValueError:
        continue
    data.append((smp, ie, te, cur_file, r[0], r[1].strip()[:100]))
ts, ti, tt = (sum(d[k] for d in data) for k in range(3))
print(f"total samples {ts:.0f}, warp instr {ti:.3g}, thread instr {tt:.3g}, mean lanes {tt/ti:.2f}")
for d in sorted(data, key=lambda d: -d[0])[:top]:
    print(f"{d[0]/ts*100:5.1f}% smp {d[1]/ti*100:5.1f}% inst lanes={d[2]/max(d[1],1):5.1f} {d[3]}:{d[4]}  {d[5]}")
