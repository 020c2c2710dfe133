#!/usr/bin/env python
"""Prints BVH facts + measured tests per ray + trace time for one config (env TRT_BVH=lbvh|ploc)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = S.CONFIGS[cfg]() if cfg in S.CONFIGS else S.tree_mesh(grid=int(cfg[4:]))
r = trt.Renderer([0]); r.scene_set(sc); r.scene_set(sc)
rp = S.params_for(sc, spp=spp or sc.spp)
acc = torch.zeros(9 * rp.width * rp.height, dtype=torch.int64, device="cuda")
best = 1e9
for _ in range(3):
    acc.zero_(); torch.cuda.synchronize(); st = r.accumulate(rp, acc.data_ptr()); best = min(best, st["ms_trace"])
rpc = S.params_for(sc, spp=min(rp.spp, 16), flags=trt.FLAG_COUNT_TESTS)
acc.zero_(); torch.cuda.synchronize(); c = r.accumulate(rpc, acc.data_ptr())
print(f"{cfg} BVH={os.environ.get('TRT_BVH','ploc')} {r.bvh_info()} ms={best:.2f} Mrays/s={st['rays']/best/1e3:.0f} "
      f"box/ray={c['box_tests']/c['rays']:.2f} tri/ray={c['tri_tests']/c['rays']:.2f} checksum={int(acc.sum().item())}")
