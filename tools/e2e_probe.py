import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S
sc = S.nature_mesh(); r = trt.Renderer([0])
rp = S.params_for(sc)
n = rp.width*rp.height
host = dict(canva=np.empty((n,3)), albedo=np.empty((n,3)), normal=np.empty((n,3)))
for it in range(3):
    t0=time.perf_counter(); r.scene_set(sc); t1=time.perf_counter(); res=r.render(rp, out=host); t2=time.perf_counter()
    print(f"scene_set {1e3*(t1-t0):.1f} ms  render {1e3*(t2-t1):.1f} ms  stats {res['stats']}")
