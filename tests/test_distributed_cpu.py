"""world_size-2 gloo test of the multi-GPU plumbing (spp split -> int64 reduce -> resolve), on CPU.
The per-rank 'kernel' is the oracle in Philox mode restricted to the rank's global sample range."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tipe_raytracer_b200 import distributed as D
from tipe_raytracer_b200 import scene as S


def test_sample_ranges_tile_the_job():
    for world in (1, 2, 3, 4, 8):
        for spp in (1, 7, 100, 1000):
            got = [D.sample_range(r, world, spp) for r in range(world)]
            assert got[0][0] == 0 and sum(n for _, n in got) == spp
            for (a, n), (b, _) in zip(got, got[1:]):
                assert a + n == b
    with pytest.raises(ValueError):
        D.sample_range(2, 2, 10)
    sc = S.cornell_spheres()
    rp = S.params_for(sc, width=16, spp=10)
    parts = [D.rank_params(rp, r, 4) for r in range(4)]
    assert [p.spp for p in parts] == [2, 3, 2, 3] and [p.sample_offset for p in parts] == [0, 2, 5, 7]
    assert all(p.sample_total == 10 for p in parts)
    weak = [D.rank_params(rp, r, 4, weak=True) for r in range(4)]
    assert [p.sample_offset for p in weak] == [0, 10, 20, 30] and all(p.sample_total == 40 and p.spp == 10 for p in weak)


def test_tile_split_rank_params():
    sc = S.cornell_spheres()
    rp = S.params_for(sc, width=64, spp=10)
    parts = [D.rank_params(rp, r, 4, split="tile") for r in range(4)]
    assert [p.tile_part & 0xffff for p in parts] == [0, 1, 2, 3] and all((p.tile_part >> 16) == 4 for p in parts)
    assert all(p.spp == 10 and p.sample_offset == 0 and p.sample_total == 10 for p in parts)      # every rank: all samples, divisor = spp
    assert D.rank_params(rp, 0, 1, split="tile").tile_part == 0


def test_resolve_host_matches_oracle_tonemap(oracle):
    rng = np.random.default_rng(5)
    sums = rng.uniform(0, 40, (3, 64))
    acc = np.zeros((9, 64), np.int64)
    acc[0:3] = np.rint(sums * D.ACCUM_SCALE)
    canva, _, _ = D.resolve_host(acc, 8, 8, 16)
    for p in range(64):
        want = oracle.tonemap(acc[0:3, p] / D.ACCUM_SCALE, 16)
        assert np.array_equal(canva[p], want)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle.oracle_py import Oracle, RNG_PHILOX
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = S.cornell_spheres()
    rp = S.params_for(sc, width=40, spp=9)
    mine = D.rank_params(rp, rank, world)
    # per-SAMPLE fixed-point conversion, like the kernels: render one sample at a time
    import dataclasses
    npix = mine.width * mine.height
    acc = np.zeros((9, npix), np.int64)
    o = Oracle()
    for s in range(mine.spp):
        one = dataclasses.replace(mine, spp=1, sample_offset=mine.sample_offset + s)
        sums = o.render(sc, one, rng=RNG_PHILOX, threads=1, want_sums=True)["sums"]      # [npix, 9]
        acc += D.to_fixed(sums.T)
    t = torch.from_numpy(acc.reshape(-1).copy())
    D.reduce_accum(t, dst=0)
    if rank == 0:
        q.put(t.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_spp_split_reduces_to_the_single_rank_image(oracle):
    import dataclasses
    import torch.multiprocessing as mp
    from oracle.oracle_py import RNG_PHILOX
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    reduced = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single rank, same 9 samples
    sc = S.cornell_spheres()
    rp = S.params_for(sc, width=40, spp=9)
    npix = rp.width * rp.height
    acc = np.zeros((9, npix), np.int64)
    for s in range(9):
        one = dataclasses.replace(rp, spp=1, sample_offset=s)
        acc += D.to_fixed(oracle.render(sc, one, rng=RNG_PHILOX, threads=1, want_sums=True)["sums"].T)
    assert np.array_equal(reduced.reshape(9, npix), acc)                     # integer sums: bit-identical
    canva, albedo, normal = D.resolve_host(reduced, rp.width, rp.height, 9)
    full = oracle.render(sc, rp, rng=RNG_PHILOX, threads=1)
    assert np.abs(canva - full["canva"]).max() <= 1                           # 2^-28 quantisation vs FP64 sums
    assert (canva != full["canva"]).mean() < 0.01
    assert np.abs(albedo - full["albedo"]).max() < 1e-6
