#!/usr/bin/env python
"""Small renders through every kernel path (sphere kernel, mesh kernel, slot-pool kernel, packet primaries, FP64 / FP32 first hit).
With a -DTRT_DEBUG_BOUNDS build of the library every BVH node / triangle / texel / stack / slot index is range-checked on the device
(compute-sanitizer is closed on the GPU pool); a violation makes the render call fail."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S
r = trt.Renderer([0])
def run(name, sc, **kw):
    r.scene_set(sc)
    rp = S.params_for(sc, **kw)
    img = r.render(rp)
    print(name, os.environ.get("TRT_ENGINE", "1"), int(img["canva"].sum()), img["stats"]["rays"], flush=True)
run("spheres+AO", S.cornell_spheres(), width=48, spp=3)
run("dielectrics", S.nested_dielectrics(), width=48, spp=2)
for eng in ("1", "3", "4"):
    os.environ["TRT_ENGINE"] = eng
    run("mesh", S.nature_mesh(), width=64, spp=3)
    sc = S.nature_mesh(); sc.use_ao = True
    run("mesh+AO", sc, width=48, spp=2)
os.environ.pop("TRT_ENGINE")
run("tree", S.tree_mesh(), width=48, spp=2)
sc = S.nature_mesh(); rp = S.params_for(sc, width=64, spp=1); r.scene_set(sc)
print("primary", int(r.primary(rp)["id"].sum()), int(r.primary_fp32(rp)["id"].sum()))
r.close()
