#!/usr/bin/env python
"""bench.py — throughput of the path-tracing hot path on N B200s, next to the reference's CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C2]

Workload (config.workload): BASELINE config 2 — the textured triangle mesh with alpha maps
(`nature` mesh standing in for the absent books mesh), 1000x750, 200 spp, nbRebondMax 4,
AO off, emissive sky sphere.  A "step" renders every sample of every pixel of that frame once.

  value   Mrays/s with the scene already resident in HBM: one launch of the persistent path-tracing
          kernel into a zeroed int64 accumulation buffer (N>1: + NCCL reduce to rank 0 + resolve).
          rays = closest_hit evaluations actually traced (primary + bounce + AO), counted on the GPU.
  e2e     the same metric through the reference-facing C ABI with HOST buffers: trt_scene_set (scene
          upload + GPU BVH build) + trt_render (trace, tonemap, 3 colour arrays copied back).
  N > 1   weak scaling: every rank renders `spp` samples of the full frame with a disjoint global sample
          range (rank r: [r*spp, (r+1)*spp)), the int64 sums are reduced over NVLink with ncclReduce issued by
          the C library itself (trt_reduce_nccl; torch.distributed only carries the rendezvous, the barriers and
          the max-over-ranks of the timings), rank 0 resolves the N*spp-sample image.
  also, on rank 0 (outside the timed region of `value`):
          all_configs            N = 1: the other BASELINE configs (C1, C3, C4, C5) and the synthetic 1 M-triangle
                                 C4big at FULL size: kernel ms, Mrays/s, roofline fraction;
          strong_c5              BASELINE config 5 as written: the 1200x900x1000 spp frame split N ways
                                 (trt_render_distributed: scene upload, trace, ncclReduce, resolve, D2H on rank 0),
                                 efficiency against rank 0 rendering the whole frame alone;
          bit_identical_vs_1gpu  rank 0 re-renders the N*spp frame alone and compares SHA-256 of canva / albedo /
                                 normal with the NCCL-reduced image (north_star test 3 across GPU counts).

--impl reference times the reference's own CPU implementation (oracle/_ref: /root/reference/main.c
compiled as a library; pthread fan-out of main.c:402-453) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mrays/s"
UNIT = "Mrays/s"


def workload(name: str):
    from tipe_raytracer_b200 import scene as S
    sc = S.CONFIGS[name]()
    desc = {
        "C1": "C1 README Cornell-style 10-sphere scene, 1000x750, 100 spp, 5 bounces, AO on",
        "C2": "C2 textured triangle mesh with alpha maps (nature mesh, 5812 tris, 31 textures; substitute for the absent book_tri.obj) + sky sphere, 1000x750, 200 spp, 4 bounces",
        "C3": "C3 nested dielectric spheres (refraction stack), 1000x750, 1000 spp, 19 bounces",
        "C4": "C4 1st_mc_world tree mesh (1792 tris, 1024^2 RGBA texture) + sky, 1920x1080, 500 spp, 6 bounces",
        "C5": "C5 README benchmark config 1200x900, 1000 spp, 6 bounces, AO on",
        "C4big": "C4big SYNTHETIC: the C4 tree instanced 24x24 = 1 032 192 triangles + sky, 1920x1080, 6 bounces",
    }[name]
    return sc, desc


def job_config(desc, sc, spp, world, seed) -> dict:
    """The `config` object of the JSON line: the SAME dict on the GPU arm and on the reference arm (the reference arm's bounded sample is
    described in its `cpu_baseline.sample`, not here)."""
    return {"workload": desc, "width": sc.width, "height": sc.height, "spp_per_gpu": spp, "spp_total": spp * world,
            "nbRebondMax": sc.bounces, "useAO": bool(sc.use_ao), "seed": seed,
            "l2": "512 MB buffer written between timed steps (L2 flush)",
            "parallelism": f"spp-split x{world}, int64 ncclReduce issued by libtipe_rt (trt_reduce_nccl)" if world > 1 else "single GPU"}


# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons WHILE the timed region runs (B200_PROFILING.md recipe).  NVML is polled in-process
    every 5 ms (a timed step is ~40 ms; forking nvidia-smi takes longer than that); nvidia-smi is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag = gpu_index, threading.Event()
        self.sm, self.sm_max, self.reasons, self.source = [], [], set(), "nvml"
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.source = "nvidia-smi"

    def _poll_nvml(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        self.sm_max.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, nm in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap")):
            if r & bit:
                self.reasons.add(nm)

    def _poll_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [c.strip() for c in out.split(",")]
        if len(r) > 8:
            self.sm.append(float(r[1])); self.sm_max.append(float(r[2]))
            for k, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                if r[5 + k].lower().startswith("active"):
                    self.reasons.add(nm)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.h is not None:
                    self._poll_nvml()
                else:
                    self._poll_smi()
            except Exception:
                if self.h is not None:
                    self.h, self.source = None, "nvidia-smi"
            self.stop_flag.wait(0.005 if self.h is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.sm_max) if self.sm_max else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(cfg: str, steps: int, warmup: int, budget_s: float = 20.0):
    """The reference's CPU path on a bounded sample of the workload; returns (Mrays/s, info)."""
    from oracle.oracle_py import Ref, Oracle, RNG_LIBC
    from tipe_raytracer_b200 import scene as S
    sc, _ = workload(cfg)
    have_ref = Ref.available()
    ncpu = os.cpu_count() or 1
    # sample: same scene / camera / bounces, reduced resolution and spp (rays/s does not depend on either)
    width, spp = (96, 2) if sc.triangles.shape[0] > 0 else (200, 4)
    rp = S.params_for(sc, width=width, spp=spp)
    times, rays = [], None

    def one(threads):
        if have_ref:
            r = Ref().render(sc, rp, threads=threads, seed=1, count_rays=True)
        else:
            import ctypes
            ctypes.CDLL(None).srand(1)
            r = Oracle().render(sc, rp, rng=RNG_LIBC, threads=1)
        return r["stats"]["seconds"], r["stats"]["rays"]

    # the reference's rand() is one locked global stream: more threads are SLOWER; use the better of {1, nproc}
    t1, rays = one(1)
    best_threads, best_t = 1, t1
    if have_ref and ncpu > 1 and t1 < budget_s:
        tn, _ = one(ncpu)
        if tn < t1:
            best_threads, best_t = ncpu, tn
    n_steps = max(1, min(steps, int(budget_s / max(best_t, 1e-3))))
    for k in range(min(warmup, 1) + n_steps):
        t, rays = one(best_threads)
        if k >= min(warmup, 1):
            times.append(t)
    sec = float(np.mean(times))
    info = {"kind": "reference" if have_ref else "port", "cores": best_threads, "host_cores": ncpu,
            "sample": f"{cfg} scene, {rp.width}x{rp.height}, {spp} spp, {rp.bounces} bounces: {rays} rays/step, "
                      f"{len(times)} step(s), {sec:.2f} s/step; threads tried 1 and {ncpu}, best {best_threads}"}
    return rays / sec / 1e6, sec, info, len(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, sec, info, n = cpu_reference_run(args.config, args.steps, args.warmup, budget_s=60.0)
    sc, desc = workload(args.config)
    from tipe_raytracer_b200 import scene as S
    info["value"] = value
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": min(args.warmup, 1),
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": job_config(desc, sc, args.spp or sc.spp, args.gpus, S.params_for(sc).seed),
        "cpu_baseline": {**info, "unit": UNIT},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------
def sha(res) -> str:
    import hashlib
    h = hashlib.sha256()
    for k in ("canva", "albedo", "normal"):
        h.update(np.ascontiguousarray(res[k]).tobytes())
    return h.hexdigest()


def fp32_peak_tflops(torch, dev):
    peaks, peak_src = measured_peaks()
    sm_clock = (peaks.get("sm_max_mhz") or 1965.0) * 1e6
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    return n_sm * 128 * 2 * sm_clock / 1e12, n_sm, sm_clock, peaks, peak_src


def config_on_device(torch, trt, S, r, name, dev, spp_override=None):
    """One BASELINE config at full size on this GPU, scene resident: kernel ms (CUDA events around the launch, inside
    the library), rays, and the FP32 intersection roofline with MEASURED tests per ray."""
    sc, desc = workload(name)
    spp = spp_override or sc.spp
    r.scene_set(sc)
    rp = S.params_for(sc, spp=spp)
    npix = rp.width * rp.height
    acc = torch.zeros(9 * npix, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    r.accumulate(S.params_for(sc, spp=max(1, spp // 10)), acc.data_ptr())          # warm-up: a tenth of the samples
    best, st = None, None
    for _ in range(2 if name in ("C1", "C2") else 1):
        acc.zero_(); torch.cuda.synchronize()
        st = r.accumulate(rp, acc.data_ptr())
        best = st["ms_trace"] if best is None else min(best, st["ms_trace"])
    acc.zero_(); torch.cuda.synchronize()
    cst = r.accumulate(S.params_for(sc, spp=min(spp, 16), flags=trt.FLAG_COUNT_TESTS), acc.data_ptr())
    n_box, n_tri, n_sph = (cst[k] / max(cst["rays"], 1) for k in ("box_tests", "tri_tests", "sphere_tests"))
    f_ray = n_box * 18 + n_tri * 55 + n_sph * 25 + 29
    peak = fp32_peak_tflops(torch, dev)[0]
    achieved = st["rays"] * f_ray / (best * 1e-3) / 1e12
    del acc
    return {"workload": desc, "width": rp.width, "height": rp.height, "spp": spp, "nbRebondMax": rp.bounces, "useAO": bool(sc.use_ao),
            "kernel_ms": best, "rays": st["rays"], "rays_per_path": st["rays"] / st["paths"], "mrays_s": st["rays"] / best / 1e3,
            "nominal_mrays_s": npix * spp * rp.bounces / best / 1e3, "msamples_s": npix * spp / best / 1e3,
            "flops_per_ray": f_ray, "box_tests_per_ray": n_box, "tri_tests_per_ray": n_tri, "sphere_tests_per_ray": n_sph,
            "roofline_frac": achieved / peak, "bvh": r.bvh_info() if sc.triangles.shape[0] else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="C2")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip all_configs / strong_c5 / bit_identical_vs_1gpu")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    # stdout carries exactly ONE line, the JSON: whatever libraries print on fd 1 (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")

    import torch
    import torch.distributed as dist
    import tipe_raytracer_b200 as trt
    from tipe_raytracer_b200 import scene as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local if world > 1 else 0)
    steps, warmup = args.steps, max(args.warmup, 3)

    sc, desc = workload(args.config)
    spp = args.spp or sc.spp
    r = trt.Renderer([dev.index])          # no CPU fallback: raises without the .so / a B200
    r.scene_set(sc)
    W, H = sc.width, sc.height
    npix = W * H
    from tipe_raytracer_b200 import distributed as D
    rp = D.rank_params(S.params_for(sc, spp=spp), rank, world, weak=True)      # rank r: global samples [r*spp, (r+1)*spp)
    ext = torch.cuda.ExternalStream(r.stream_ptr(), device=dev)
    accum = torch.zeros(9 * npix, dtype=torch.int64, device=dev)
    out_dev = torch.zeros(9 * npix, dtype=torch.float64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    # the data-path collective is the library's own ncclReduce (trt_reduce_nccl): its communicator is made from an id
    # rank 0 draws and torch.distributed hands round (plumbing)
    comm = 0
    if world > 1:
        box = [r.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = r.nccl_comm_init(world, rank, box[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        """value: scene resident in HBM.  One kernel of ours per step (+1 resolve on rank 0 when N>1)."""
        with torch.cuda.stream(ext):
            accum.zero_()
        st = r.accumulate(rp, accum.data_ptr())
        if world > 1:
            r.reduce_nccl(comm, accum.data_ptr(), 9 * npix, 0)          # int64 SUM over NVLink, on the library's stream
            if rank == 0:
                r.resolve_device(rp, accum.data_ptr(), out_dev.data_ptr(), out_dev.data_ptr() + 24 * npix, out_dev.data_ptr() + 48 * npix)
        return st

    def timed(fn, n):
        tot, stats = 0.0, []
        for _ in range(n):
            flush.zero_()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            stats.append(fn())
            e1.record(ext)
            barrier()
            tot += e0.elapsed_time(e1)
        return tot, stats

    for _ in range(warmup):
        step_device()
    sampler = ClockSampler(dev.index)
    sampler.start()
    t_ms, stats = timed(step_device, steps)
    clocks = sampler.summary()
    rays_local = float(sum(s["rays"] for s in stats))
    kernel_ms = float(np.mean([s["ms_trace"] for s in stats]))
    tt = torch.tensor([t_ms, rays_local], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_ms, rays_total = float(tmax[0]), float(tsum[1])
    else:
        rays_total = rays_local
    value = rays_total / (t_ms * 1e-3) / 1e6
    ms_per_step = t_ms / steps

    # ---- e2e: the C-ABI call a user makes, host buffers, H2D + D2H inside the timed region --------
    host = dict(canva=np.empty((npix, 3)), albedo=np.empty((npix, 3)), normal=np.empty((npix, 3)))
    h2d = int(sc.spheres.nbytes + sc.triangles.nbytes + sc.mat_list.nbytes + sc.quel_mat_pour_tri.nbytes)
    d2h = 3 * npix + 2 * npix * 24          # canva crosses PCIe as 3 bytes per pixel (widened to doubles on the host), the two AOV tables as doubles
    e2e_parts = {"scene_set_ms": 0.0}
    rp_job = S.params_for(sc, spp=spp * world)                # the N*spp-sample frame; every rank takes its N-th of the samples

    def step_e2e():
        t0 = time.perf_counter()
        r.scene_set(sc)                                   # H2D of the reference-layout arrays + BVH build
        e2e_parts["scene_set_ms"] += (time.perf_counter() - t0) * 1e3
        if world == 1:
            return r.render(rp_job, out=host)["stats"]      # trt_render: trace + resolve + D2H
        return r.render_distributed(rp_job, comm, rank, world, out=host if rank == 0 else None)["stats"]   # + ncclReduce

    step_e2e(); step_e2e()
    e2e_parts["scene_set_ms"] = 0.0
    e2e_steps = max(10, steps)
    t0 = time.perf_counter()
    te_ms, est = timed(step_e2e, e2e_steps)
    e2e_wall = time.perf_counter() - t0
    e_rays = float(sum(s["rays"] for s in est))
    te = torch.tensor([te_ms, e_rays], dtype=torch.float64, device=dev)
    if world > 1:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        te_ms, e_rays = float(a[0]), float(b[1])
    e2e_value = e_rays / (te_ms * 1e-3) / 1e6
    e2e_hash = sha(host) if rank == 0 else None
    # the same call with the three output arrays in page-locked memory (trt_host_alloc instead of main.c's malloc): the GPU
    # writes them directly, no staging and no host copy threads.  Reported beside e2e, which keeps main.c's plain (pageable) arrays.
    e2e_pinned = None
    if world == 1 and not args.no_extras:
        pinned = trt.host_arrays(npix, r.lib)
        host_keep, host = host, pinned
        step_e2e()
        tp_ms, pst = timed(step_e2e, e2e_steps)
        e2e_pinned = {"value": float(sum(s["rays"] for s in pst)) / (tp_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": tp_ms / e2e_steps, "steps": e2e_steps,
                      "same_image": bool(sha(pinned) == e2e_hash), "what": "as e2e, output arrays from trt_host_alloc (page-locked)"}
        host = host_keep

    # ---- extras (outside the timed regions above) -------------------------------------------------------
    extras = {}
    if not args.no_extras:
        # north_star test 3: the image of N GPUs (NCCL-reduced, above) against ONE GPU rendering all N*spp samples
        if world > 1:
            if rank == 0:
                alone = r.render(rp_job)
                extras["bit_identical_vs_1gpu"] = {"weak_" + args.config: bool(sha(alone) == e2e_hash), "spp_total": spp * world,
                                                   "sha256_n_gpus": e2e_hash, "sha256_1_gpu": sha(alone)}
            barrier()
        # BASELINE config 5 as written: strong scaling of the 1200x900x1000 spp frame
        sc5, desc5 = workload("C5")
        rp5 = S.params_for(sc5)
        n5 = rp5.width * rp5.height
        # output arrays of rank 0: page-locked (trt_host_alloc), like apps/main_b200 allocates them; the N = 1 reference frame uses the same arrays
        host5 = trt.host_arrays(n5, r.lib) if rank == 0 else None

        def step_c5():
            r.scene_set(sc5)
            if world == 1:
                return r.render(rp5, out=host5)["stats"]
            return r.render_distributed(rp5, comm, rank, world, out=host5)["stats"]

        step_c5()                                          # warm-up: buffers of this size, pinned staging
        c5_steps = 3
        t5_ms, st5 = timed(step_c5, c5_steps)
        t5 = torch.tensor([t5_ms, float(sum(s["rays"] for s in st5)), float(max(s["ms_trace"] for s in st5))], dtype=torch.float64, device=dev)
        if world > 1:
            a = t5.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = t5.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            t5_ms, rays5, trace5 = float(a[0]), float(b[1]), float(a[2])
        else:
            rays5, trace5 = float(t5[1]), float(t5[2])
        strong = {"workload": desc5, "n_gpus": world, "steps": c5_steps, "ms_per_frame": t5_ms / c5_steps, "trace_ms_max_rank": trace5,
                  "mrays_s": rays5 / (t5_ms * 1e-3) / 1e6, "msamples_s": n5 * rp5.spp * c5_steps / (t5_ms * 1e-3) / 1e6,
                  "what": "trt_scene_set + trt_render_distributed (trace of spp/N, ncclReduce int64, resolve, D2H on rank 0), page-locked host output arrays (trt_host_alloc), max over ranks"}
        if world > 1:
            if rank == 0:
                h_n = sha(host5)
                t0 = time.perf_counter()
                r.scene_set(sc5); alone5 = r.render(rp5, out=host5)        # the whole frame on rank 0 alone, same call sequence
                one_ms = (time.perf_counter() - t0) * 1e3
                strong["ms_per_frame_1gpu_same_box"] = one_ms
                strong["efficiency_vs_1gpu"] = one_ms / (world * t5_ms / c5_steps)
                extras.setdefault("bit_identical_vs_1gpu", {})["strong_C5"] = bool(sha(alone5) == h_n)
            barrier()
        extras["strong_c5"] = strong
        del host5
        # every other BASELINE config at full size, N = 1 (scene resident, kernel-timed)
        if world == 1:
            allc = {}
            for name, spp_o in (("C1", None), ("C3", None), ("C4", None), ("C5", None), ("C4big", 100)):
                try:
                    allc[name] = config_on_device(torch, trt, S, r, name, dev, spp_o)
                except Exception as e:           # a config that fails must not take the headline line with it
                    allc[name] = {"error": str(e)[:200]}
            extras["all_configs"] = allc
            r.scene_set(sc)                      # back to the headline workload for the roofline pass below

    if rank == 0:
        # ---- roofline of the dominant kernel (pathtrace_kernel): measured tests per ray x FLOPs per test
        rpc = S.params_for(sc, spp=min(spp, 16), flags=trt.FLAG_COUNT_TESTS)
        accum.zero_(); torch.cuda.synchronize()
        cst = r.accumulate(rpc, accum.data_ptr())
        n_box, n_tri, n_sph = (cst[k] / max(cst["rays"], 1) for k in ("box_tests", "tri_tests", "sphere_tests"))
        f_ray = n_box * 18 + n_tri * 55 + n_sph * 25 + 29
        rays_per_launch = rays_local / steps
        fp32_peak, n_sm, sm_clock, peaks, peak_src = fp32_peak_tflops(torch, dev)
        achieved = rays_per_launch * f_ray / (kernel_ms * 1e-3) / 1e12
        chunk = max(4, min(16, (spp + 10) // 20)) if sc.triangles.shape[0] > 0 else 16      # samples per item (api.cu)
        hbm_alg = 9 * 8 * npix * ((spp + chunk - 1) // chunk)       # int64 RED traffic of the item flushes, per launch
        traffic = None
        try:        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, from the committed ncu capture
            cand = [os.path.join(ROOT, "profiles", n) for n in ("r02_traffic_c2.json", "r01_traffic_c2.json")]
            with open(next(c for c in cand if os.path.exists(c))) as f:
                tj = json.load(f)
            if args.config == "C2" and spp == sc.spp:
                traffic = {"bytes_per_launch": tj["traffic_bytes_per_launch"], "source": tj["source"]}
        except Exception:
            pass
        roofline = {
            "bound": "fp32", "kernel": "pathtrace_kernel (persistent megakernel, 4-wide BVH)", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": achieved / fp32_peak, "traffic": traffic,
            "peak_source": f"{n_sm} SMs x 128 FP32 lanes x 2 x {sm_clock/1e6:.0f} MHz (sm_max_mhz from MEASURED_PEAKS.json, {peak_src})",
            "flops_per_ray": f_ray, "box_tests_per_ray": n_box, "tri_tests_per_ray": n_tri, "sphere_tests_per_ray": n_sph,
            "box_tests_note": "child boxes actually tested: empty slots of the 4-wide nodes are not counted",
            "rays_per_launch": rays_per_launch, "kernel_ms": kernel_ms,
            "hbm": {"algorithmic_bytes_per_launch": hbm_alg, "achieved_gbs": hbm_alg / (kernel_ms * 1e-3) / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs"), "note": "working set (BVH+triangles+texels) is L2-resident; HBM is not the limiter"},
        }
        cpu = None
        if world == 1 and not args.no_cpu:
            v, sec, info, _ = cpu_reference_run(args.config, 3, 0, budget_s=12.0)
            cpu = {"value": v, "unit": UNIT, **info}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": job_config(desc, sc, spp, world, rp.seed),
            "rays_per_step": rays_total / steps, "rays_per_path": rays_local / steps / (npix * spp),
            "nominal_mrays_s": npix * spp * world * sc.bounces / (ms_per_step * 1e-3) / 1e6,
            "msamples_s": npix * spp * world / (ms_per_step * 1e-3) / 1e6,
            "clocks": clocks, "gpu_launches": int(sum(s["kernel_launches"] for s in stats)) + (steps if world > 1 else 0),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": te_ms / e2e_steps, "steps": e2e_steps, "wall_s": e2e_wall,
                    "scene_set_ms_per_step": e2e_parts["scene_set_ms"] / e2e_steps,
                    "what": ("trt_scene_set (upload + GPU BVH build) + trt_render into host color arrays" if world == 1 else
                             "trt_scene_set + trt_render_distributed (N*spp samples split over the ranks, ncclReduce, resolve, D2H on rank 0)")},
            "roofline": roofline, "bvh": r.bvh_info(),
        }
        line.update(extras)
        if e2e_pinned:
            line["e2e_page_locked_outputs"] = e2e_pinned
        if cpu:
            line["cpu_baseline"] = cpu
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if comm:
        r.nccl_comm_destroy(comm)
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
