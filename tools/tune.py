#!/usr/bin/env python
"""Sweeps tuning knobs on one config and prints ms per trace.
   python tools/tune.py C2 [spp] KNOB=v1,v2 ...   (KNOB -> env TRT_TUNE_<KNOB>, except ENGINE -> TRT_ENGINE)"""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
rest = sys.argv[2:]
spp = 0
if rest and "=" not in rest[0]:
    spp = int(rest[0]); rest = rest[1:]
grid = {}
for a in rest:
    k, v = a.split("=")
    grid["TRT_ENGINE" if k == "ENGINE" else "TRT_TUNE_" + k] = v.split(",")
sc = S.CONFIGS[cfg]() if cfg in S.CONFIGS else S.tree_mesh(grid=int(cfg[4:]))
r = trt.Renderer([0]); r.scene_set(sc)
rp = S.params_for(sc, spp=spp or sc.spp)
n = rp.width * rp.height
acc = torch.zeros(9 * n, dtype=torch.int64, device="cuda")
keys = list(grid)
for combo in itertools.product(*[grid[k] for k in keys]):
    for k, v in zip(keys, combo):
        os.environ[k] = v
    best = 1e9
    for it in range(3):
        acc.zero_(); torch.cuda.synchronize()
        st = r.accumulate(rp, acc.data_ptr())
        best = min(best, st["ms_trace"])
    tag = " ".join(f"{k.replace('TRT_TUNE_', '').replace('TRT_', '')}={v}" for k, v in zip(keys, combo))
    print(f"{cfg} {tag} ms={best:.2f} Mrays/s={st['rays']/best/1e3:.0f} rays/path={st['rays']/st['paths']:.3f} launches={st['kernel_launches']} checksum={int(acc.sum().item())}", flush=True)
