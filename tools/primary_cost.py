#!/usr/bin/env python
"""What the primary rays cost: a config traced with nbRebondMax = 1 (camera rays only) against the full bounce count, with the
box / triangle tests per ray of each.   python tools/primary_cost.py C2 [spp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
sc = S.CONFIGS[cfg]()
spp = int(sys.argv[2]) if len(sys.argv) > 2 else sc.spp
r = trt.Renderer([0]); r.scene_set(sc)
n = sc.width * sc.height
acc = torch.zeros(9 * (S.params_for(sc).width * S.params_for(sc).height), dtype=torch.int64, device="cuda")
for bounces in (1, sc.bounces):
    rp = S.params_for(sc, spp=spp); rp.bounces = bounces
    best = 1e9
    for _ in range(3):
        acc.zero_(); torch.cuda.synchronize()
        st = r.accumulate(rp, acc.data_ptr()); best = min(best, st["ms_trace"])
    rpc = S.params_for(sc, spp=spp, flags=trt.FLAG_COUNT_TESTS); rpc.bounces = bounces
    acc.zero_(); torch.cuda.synchronize()
    cs = r.accumulate(rpc, acc.data_ptr())
    print(f"{cfg} nbRebondMax={bounces}: {best:.2f} ms, rays {st['rays']/1e6:.1f} M, box/ray {cs['box_tests']/cs['rays']:.2f}, tri/ray {cs['tri_tests']/cs['rays']:.2f}", flush=True)
