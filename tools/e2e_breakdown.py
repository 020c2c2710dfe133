#!/usr/bin/env python
"""Where the end-to-end frame goes: python tools/e2e_breakdown.py [C2]   (trt_scene_set + trt_render with host buffers)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
sc = S.CONFIGS[cfg](); rp = S.params_for(sc)
n = rp.width * rp.height
host = dict(canva=np.empty((n, 3)), albedo=np.empty((n, 3)), normal=np.empty((n, 3)))
r = trt.Renderer([0])
for k in range(3): r.scene_set(sc); r.render(rp, out=host)
rows = []
for k in range(10):
    t0 = time.perf_counter(); r.scene_set(sc); t1 = time.perf_counter(); st = r.render(rp, out=host)["stats"]; t2 = time.perf_counter()
    rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, st["ms_trace"], st["ms_reduce_resolve"], st["ms_d2h"], st["ms_total"]))
a = np.median(np.array(rows), axis=0)
print(f"{cfg}: scene_set {a[0]:.3f} ms | render call {a[1]:.3f} ms = trace {a[2]:.3f} + resolve {a[3]:.3f} + d2h/host {a[4]:.3f} (ms_total {a[5]:.3f}) | frame {a[0]+a[1]:.3f} ms, overhead over the trace {a[0]+a[1]-a[2]:.3f} ms")
for thr in ("1", "2", "4", "8", "16"):
    os.environ["TRT_TUNE_COPY_THREADS"] = thr
    ts = []
    for k in range(6): ts.append(r.render(rp, out=host)["stats"]["ms_d2h"])
    print(f"  copy threads {thr}: d2h/host {np.median(ts):.3f} ms")
