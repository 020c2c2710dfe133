#!/usr/bin/env python
"""Turns an .ncu-rep (brought back in gpurun_out/) into a small text summary under profiles/:
selected raw metrics + the hottest CUDA source lines.   python tools/export_profile.py rep.ncu-rep out.md [title]"""
import csv, io, subprocess, sys, os
rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__t_sectors_op_red.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# {title}", "", f"source: `{os.path.basename(rep)}` (ncu --set full --clock-control none --import-source on; cold-cache, serialised replay)", ""]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    lines += [f"## {name[:110]}", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| {k} | {r[i]} | {units[i]} |")
    # SURVEY §8d cross-check: FLOP/s from the executed FP32 instructions (FFMA = 2) next to the algorithmic F_ray figure of bench.py
    try:
        g = lambda k: float(r[hdr.index(k)])
        fl = 2 * g("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum") + g("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum") + g("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum")
        ms = g("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "msecond": 1.0, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6, "second": 1e3, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)
        lines.append(f"| FP32 FLOP from FFMA x2 + FADD + FMUL | {fl/1e9:.1f} GFLOP = {fl/ms/1e9:.2f} TFLOP/s = {fl/ms/1e9/74.45*100:.1f} % of 74.45 TFLOP/s | derived |")
    except Exception:
        pass
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
try:
    h = next(i for i, r in enumerate(srows) if r and r[0] == "Line No")
    cols = srows[h]; ci = {}
    for i, c in enumerate(cols): ci.setdefault(c, i)
    cur, data = "", []
    for r in srows[:h]:
        if r and r[0] == "File Path": cur = r[1].split("/")[-1]
    for r in srows[h + 1:]:
        if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if len(r) < len(cols) or r[0] in ("", "Line No"): continue
        try: smp, ie, te = float(r[ci["# Samples"]]), float(r[ci["Instructions Executed"]]), float(r[ci["Thread Instructions Executed"]])
        except ValueError: continue
        data.append((smp, ie, te, cur, r[0], r[1].strip()[:95]))
    ts, ti, tt = (sum(d[k] for d in data) or 1 for k in range(3))
    lines += ["## hottest source lines (share of warp-stall samples, share of warp instructions, mean active lanes)", "", "```"]
    for d in sorted(data, key=lambda d: -d[0])[:25]:
        lines.append(f"{d[0]/ts*100:5.1f}% smp {d[1]/ti*100:5.1f}% inst lanes={d[2]/max(d[1],1):5.1f} {d[3]}:{d[4]}  {d[5]}")
    lines += ["```", ""]
except StopIteration:
    pass
open(out, "w").write("\n".join(lines))
print("wrote", out)
