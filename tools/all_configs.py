#!/usr/bin/env python
"""Device-resident throughput of the five BASELINE configs at full size (one GPU), as a markdown table."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S
r = trt.Renderer([0])
print("| cfg | scene | res x spp x bounces | paths | rays/path | trace ms | Mrays/s | Msamples/s | nominal Mrays/s |")
print("|---|---|---|---|---|---|---|---|---|")
for cfg in (sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5"]):
    sc = S.CONFIGS[cfg](); r.scene_set(sc)
    rp = S.params_for(sc)
    n = rp.width * rp.height
    acc = torch.zeros(9 * n, dtype=torch.int64, device="cuda")
    best = None
    for _ in range(2):
        acc.zero_(); torch.cuda.synchronize(); st = r.accumulate(rp, acc.data_ptr())
        if best is None or st["ms_trace"] < best["ms_trace"]: best = st
    ms = best["ms_trace"]
    print(f"| {cfg} | {sc.name} | {rp.width}x{rp.height} x {rp.spp} x {rp.bounces}{' AO' if rp.use_ao else ''} | {best['paths']/1e6:.0f} M | "
          f"{best['rays']/best['paths']:.3f} | {ms:.1f} | {best['rays']/ms/1e3:.0f} | {best['paths']/ms/1e3:.0f} | {best['paths']*rp.bounces/ms/1e3:.0f} |", flush=True)
