#!/usr/bin/env python
"""Multi-GPU evidence in one gpurun --gpus N call:  python tools/multi_gpu_check.py N
  1. in-process trt_render over N devices (peer-load resolve), C5 at full size: ms per frame (second frame: buffers exist),
     bit-identical to one GPU, spp split and tile split;
  2. one PROCESS per GPU through the plain-C host program (apps/main_b200 --nranks N, NCCL id through a file), C5 at full size:
     wall time of the slowest rank's render call, image identical to the in-process one."""
import hashlib, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sha = lambda r: hashlib.sha256(b"".join(np.ascontiguousarray(r[k]).tobytes() for k in ("canva", "albedo", "normal"))).hexdigest()[:16]
sc = S.readme_benchmark()
rp = S.params_for(sc)
npx = rp.width * rp.height
host = dict(canva=np.empty((npx, 3)), albedo=np.empty((npx, 3)), normal=np.empty((npx, 3)))
one = trt.Renderer([0]); one.scene_set(sc); one.render(rp, out=host)
t0 = time.perf_counter(); one.scene_set(sc); a = one.render(rp, out=host); t1 = (time.perf_counter() - t0) * 1e3
h1 = sha(a); print(f"C5 1 GPU in-process: {t1:.2f} ms per frame (scene_set + trt_render), trace {a['stats']['ms_trace']:.2f} ms, sha {h1}", flush=True)
one.close()
many = trt.Renderer(list(range(n))); many.scene_set(sc); many.render(rp, out=host)
for flags, name in ((0, "spp split"), (trt.FLAG_TILE_SPLIT, "tile split")):
    rpf = S.params_for(sc, flags=flags)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); many.scene_set(sc); b = many.render(rpf, out=host); best = min(best, (time.perf_counter() - t0) * 1e3)
    st = b["stats"]
    print(f"C5 {n} GPUs in-process, {name}: {best:.2f} ms per frame (scene_set + trt_render), trace {st['ms_trace']:.2f} ms, reduce+resolve {st['ms_reduce_resolve']:.2f} ms, "
          f"d2h+host {st['ms_d2h']:.2f} ms, efficiency vs 1 GPU {t1 / (n * best):.3f}, bit-identical to 1 GPU: {sha(b) == h1}", flush=True)
pinned = trt.host_arrays(npx, many.lib)
for flags, name in ((0, "spp split, page-locked output arrays (trt_host_alloc)"),):
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter(); many.scene_set(sc); b = many.render(S.params_for(sc, flags=flags), out=pinned); best = min(best, (time.perf_counter() - t0) * 1e3)
    st = b["stats"]
    print(f"C5 {n} GPUs in-process, {name}: {best:.2f} ms per frame, trace {st['ms_trace']:.2f} ms, reduce+resolve {st['ms_reduce_resolve']:.2f} ms, "
          f"d2h {st['ms_d2h']:.2f} ms, efficiency vs 1 GPU {t1 / (n * best):.3f}, bit-identical to 1 GPU: {sha(b) == h1}", flush=True)
many.close()
exe = os.path.join(ROOT, "apps", "main_b200")
subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "apps"), "main_b200"], stderr=subprocess.DEVNULL)
args = ["--width", "1200", "--spp", "1000", "--bounces", "6"]
out1, outn, idf = "/tmp/c5_one.ppm", "/tmp/c5_many.ppm", "/tmp/c5_nccl.id"
subprocess.check_call([exe, *args, "--out", out1], stdout=subprocess.DEVNULL)
if os.path.exists(idf): os.remove(idf)
t0 = time.perf_counter()
procs = [subprocess.Popen([exe, *args, "--nranks", str(n), "--rank", str(r), "--id-file", idf, "--out", outn], stdout=subprocess.PIPE, text=True) for r in range(n - 1, -1, -1)]
outs = [p.communicate(timeout=300)[0] for p in procs]
wall = time.perf_counter() - t0
print(f"C5 {n} processes (apps/main_b200 --nranks {n}, ncclReduce in libtipe_rt): wall {wall:.2f} s incl. process start, CUDA and NCCL init; image identical to 1 process: {open(out1).read() == open(outn).read()}")
print(outs[-1].strip().splitlines()[-1])
