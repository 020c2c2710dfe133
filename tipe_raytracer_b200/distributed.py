"""Host logic of the multi-GPU path: one process per GPU, spp split, int64 NCCL reduce.

The path shards by SAMPLES (BASELINE config 5): rank r renders the global sample range
[r*spp_per_rank, (r+1)*spp_per_rank) of EVERY pixel into int64 fixed-point sums (value * 2^28), the sums
are added over NVLink with one `reduce(SUM)` to rank 0, and rank 0 resolves (tonemaps) the image.
Integer addition is associative and the Philox counters carry the GLOBAL sample index, so the reduced
image is bit-identical to a single-GPU render of the same samples, whatever the number of ranks.

Nothing here touches CUDA kernels: it is plumbing around Renderer.accumulate / Renderer.resolve and is
covered on CPU by world_size-2 gloo tests (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import dataclasses
from typing import Tuple

import numpy as np

from .scene import RenderParams

ACCUM_FRAC_BITS = 28
ACCUM_SCALE = float(1 << ACCUM_FRAC_BITS)
SAMPLE_CLAMP = float(1 << 20)


def sample_range(rank: int, world: int, spp_total: int) -> Tuple[int, int]:
    """Strong split of spp_total samples: (first global sample, count) of `rank`. Ranges tile [0, spp_total)."""
    if world < 1 or not (0 <= rank < world) or spp_total < 0:
        raise ValueError("bad rank/world/spp")
    lo = spp_total * rank // world
    hi = spp_total * (rank + 1) // world
    return lo, hi - lo


def rank_params(rp: RenderParams, rank: int, world: int, weak: bool = False, split: str = "spp") -> RenderParams:
    """RenderParams for one rank.  weak=False: rp.spp is the job total, split over ranks (BASELINE C5).
    weak=True: every rank renders rp.spp samples, the job has world*rp.spp (bench.py's scaling runs).
    split="tile" (SURVEY §8e partitioning B): every rank renders ALL samples of the interleaved 8x4 tiles t with
    t mod world == rank; the other pixels of its sum buffer stay zero, so the same int64 reduce assembles the image."""
    if split == "tile":
        return dataclasses.replace(rp, tile_part=(rank | (world << 16)) if world > 1 else 0, sample_total=rp.sample_total or rp.spp)
    if weak:
        return dataclasses.replace(rp, sample_offset=rp.sample_offset + rank * rp.spp, sample_total=world * rp.spp)
    lo, n = sample_range(rank, world, rp.spp)
    return dataclasses.replace(rp, spp=n, sample_offset=rp.sample_offset + lo, sample_total=rp.spp)


def to_fixed(values: np.ndarray) -> np.ndarray:
    """Per-sample float -> int64 fixed point exactly as the kernels do (trt_device.cuh to_fixed)."""
    v = np.nan_to_num(np.asarray(values, np.float32), nan=0.0)
    v = np.clip(v, -SAMPLE_CLAMP, SAMPLE_CLAMP)
    return np.rint(v.astype(np.float64) * ACCUM_SCALE).astype(np.int64)


def resolve_host(accum: np.ndarray, width: int, height: int, divisor: int):
    """Host statement of the resolve kernel (write_color_canva, rtutility.h:56-71, + main.c:275-279) on planar
    int64 sums [9][H*W] -> (canva, albedo, normal) float64 [H*W, 3].  Used to check the reduce plumbing."""
    npix = width * height
    a = np.asarray(accum, np.int64).reshape(9, npix).astype(np.float64) / ACCUM_SCALE
    mean = (1.0 / divisor) * a[0:3]
    r = np.sqrt(mean.astype(np.float32)).astype(np.float64)
    r = np.where(r < 0.0, 0.0, np.where(r > 0.999, 0.999, r))
    canva = np.trunc(256.0 * r).T.copy()
    albedo = (a[3:6] / divisor).T.copy()
    normal = (a[6:9] / divisor).T.copy()
    return canva, albedo, normal


def reduce_accum(accum, dst: int = 0):
    """Sum the int64 accumulation tensors of all ranks onto `dst` (NCCL over NVLink on GPUs, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
