#!/usr/bin/env python
"""Trace time and ray count of a config as a function of nbRebondMax (1 = primary rays only):
   the marginal cost of each bounce level.   python tools/depth_profile.py C2 [ORDER]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
sc = S.CONFIGS[cfg]()
r = trt.Renderer([0]); r.scene_set(sc)
n = sc.width * int(sc.width / sc.ratio)
acc = torch.zeros(9 * n, dtype=torch.int64, device="cuda")
for b in ([int(x) for x in os.environ['DEPTHS'].split(',')] if os.environ.get('DEPTHS') else range(1, sc.bounces + 1)):
    rp = S.params_for(sc, bounces=b)
    best = 1e9
    for it in range(3):
        acc.zero_(); torch.cuda.synchronize()
        st = r.accumulate(rp, acc.data_ptr())
        best = min(best, st["ms_trace"])
    acc.zero_(); torch.cuda.synchronize()
    c = r.accumulate(S.params_for(sc, bounces=b, spp=16, flags=trt.FLAG_COUNT_TESTS), acc.data_ptr())
    d = r.debug_schedule()
    if d["trips"]:
        print(f"  sched: trips/ray={d['trips']*32/c['rays']:.2f} (x32 lanes) inside/trip={d['inside']/d['trips']:.1f} node/trip={d['node']/d['trips']:.1f} "
              f"leaf/trip={d['leaf']/d['trips']:.1f} trips/session={d['trips']/max(d['sessions'],1):.1f} hk/session={d['hk']/max(d['sessions'],1):.2f} consume/hk={d['consume']/max(d['hk'],1):.1f}")
    print(f"  box/ray={c['box_tests']/c['rays']:.2f} tri/ray={c['tri_tests']/c['rays']:.2f} (box tests {c['box_tests']}, tri tests {c['tri_tests']}, rays {c['rays']} at 16 spp)")
    print(f"{cfg} nbRebondMax={b} ms={best:.2f} rays={st['rays']} rays/path={st['rays']/st['paths']:.3f} Mrays/s={st['rays']/best/1e3:.0f}", flush=True)
