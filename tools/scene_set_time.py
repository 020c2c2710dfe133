import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S
r = trt.Renderer([0])
for cfg in ("C2", "C4", "C3b", "C1"):
    sc = S.CONFIGS[cfg]()
    for mode, pre in (("1", "0"), ("1", "4"), ("1", "6"), ("1", "8"), ("0", "0")):
        os.environ["TRT_TUNE_SMALL_BUILD"] = mode; os.environ["TRT_TUNE_SMALL_PRE"] = pre
        r.scene_set(sc); r.scene_set(sc)
        ts = []
        for _ in range(10):
            t0 = time.perf_counter(); r.scene_set(sc); ts.append((time.perf_counter() - t0) * 1e3)
        rp = S.params_for(sc, width=200, spp=4)
        img = r.render(rp)
        print(cfg, f"single-CTA build after {pre} multi-block passes" if mode == "1" else "multi-kernel build", f"scene_set min {min(ts):.3f} ms median {np.median(ts):.3f} ms", r.bvh_info(), "canva sum", img["canva"].sum(), "rays", img["stats"]["rays"], flush=True)
