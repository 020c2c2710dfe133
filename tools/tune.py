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
sched = "--sched" in rest            # also run once with TRT_FLAG_COUNT_TESTS and print the schedule counters
rest = [a for a in rest if a != "--sched"]
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
    if sched:
        acc.zero_(); torch.cuda.synchronize()
        rpc = S.params_for(sc, spp=spp or sc.spp, flags=trt.FLAG_COUNT_TESTS)
        cs = r.accumulate(rpc, acc.data_ptr()); d = r.debug_schedule()
        rays = max(cs["rays"], 1); trips = max(d["trips"], 1)
        print(f"   sched: box/ray={cs['box_tests']/rays:.2f} tri/ray={cs['tri_tests']/rays:.2f} trips/ray-lane={32*d['trips']/rays:.2f} inside/trip={d['inside']/trips:.1f} "
              f"node/trip={d['node']/trips:.1f} leaf/trip={d['leaf']/trips:.1f} trips/session={d['trips']/max(d['sessions'],1):.1f} "
              f"shade passes={d['hk']} consume/pass={d['consume']/max(d['hk'],1):.1f} count_ms={cs['ms_trace']:.1f}", flush=True)
    print(f"{cfg} {tag} ms={best:.2f} Mrays/s={st['rays']/best/1e3:.0f} rays/path={st['rays']/st['paths']:.3f} launches={st['kernel_launches']} checksum={int(acc.sum().item())}", flush=True)
