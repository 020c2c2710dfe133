#!/usr/bin/env python
"""Sweeps the kernel's tuning knobs (env TRT_TUNE_*) on one config; prints ms per launch."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = S.CONFIGS[cfg]()
r = trt.Renderer([0]); r.scene_set(sc)
rp = S.params_for(sc, spp=spp or sc.spp)
n = rp.width * rp.height
acc = torch.zeros(9 * n, dtype=torch.int64, device="cuda")
grid = {"TRT_TUNE_MIN_STEPS": sys.argv[3].split(",") if len(sys.argv) > 3 else ["8"],
        "TRT_TUNE_EXIT": sys.argv[4].split(",") if len(sys.argv) > 4 else ["20"],
        "TRT_TUNE_BLOCKS": sys.argv[5].split(",") if len(sys.argv) > 5 else ["2"],
        "TRT_TUNE_CHUNK": sys.argv[6].split(",") if len(sys.argv) > 6 else ["32"],
        "TRT_TUNE_VOTE": sys.argv[7].split(",") if len(sys.argv) > 7 else ["0"]}
keys = list(grid)
for combo in itertools.product(*[grid[k] for k in keys]):
    for k, v in zip(keys, combo):
        os.environ[k] = v
    best = 1e9
    for it in range(3):
        acc.zero_(); torch.cuda.synchronize()
        st = r.accumulate(rp, acc.data_ptr())
        best = min(best, st["ms_trace"])
    print(" ".join(f"{k[9:]}={v}" for k, v in zip(keys, combo)), f"ms={best:.2f} Mrays/s={st['rays']/best/1e3:.0f} rays/path={st['rays']/st['paths']:.3f}", flush=True)
