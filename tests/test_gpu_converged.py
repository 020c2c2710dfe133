"""Converged-radiance parity (north_star test 2) and the FP32 first hit (test 1, second half).

  * signed bias at HIGH spp: the GPU (FP32, BVH) and the FP64 oracle are driven by the SAME Philox stream, so at
    4096 spp the sampling noise cancels and what is left is the effect of FP32 decisions that fall on the other side
    (a texel edge, a silhouette, tri_tmin standing in for 1e-7, MUFU approximations).  Asserted on the LINEAR sums
    (the sqrt tonemap would hide a bias): per-channel mean within 0.1 % (0.5 % asked; 0.016 % observed), and the per-pixel differences must not lean
    to one side;
  * noise floor: canva against the reference's own rand()-stream render (oracle/_ref = main.c compiled unmodified,
    seeds 1 and 2): RMSE(GPU, ref1) <= 1.10 x RMSE(ref1, ref2) (1.25 asked), per-channel means within 1 %;
  * the ids the path tracer's own FP32 / BVH arithmetic finds for the pixel-centre rays against the FP64 pass
    (which is bit-exact against the CPU): flip rate at full resolution.
Observed values are printed (pytest -s / -rP) so that the thresholds can be read against them.
"""
import numpy as np
import pytest
import torch

from oracle.oracle_py import RNG_PHILOX
from tipe_raytracer_b200 import distributed as D
from tipe_raytracer_b200 import scene as S

pytestmark = pytest.mark.gpu


def _gpu_sums(rt, sc, rp):
    n = rp.width * rp.height
    acc = torch.zeros(9 * n, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    st = rt.accumulate(rp, acc.data_ptr())
    torch.cuda.synchronize()
    return acc.view(9, n).double().cpu().numpy().T / D.ACCUM_SCALE, st          # [npix, 9] linear sums


# spp >= 4096 on the sphere scenes, >= 256 on the meshes (the FP64 oracle scans every triangle per ray)
@pytest.mark.parametrize("cfg,spp", [("C1", 4096), ("C3", 4096), ("C2", 256), ("C4", 256), ("C3b", 256)])
def test_high_spp_signed_bias_vs_fp64_oracle_same_stream(rt, oracle, cfg, spp):
    sc = S.CONFIGS[cfg]()
    rp = S.params_for(sc, width=64, spp=spp)                       # 64 x 48 (C4: 64 x 36, C3b: 64 x 40)
    rt.scene_set(sc)
    g, st = _gpu_sums(rt, sc, rp)
    o = oracle.render(sc, rp, rng=RNG_PHILOX, want_sums=True)       # all host threads: Philox mode is thread-invariant
    gs, os_ = g / spp, o["sums"] / spp                              # per-pixel means: radiance 0:3, albedo 3:6, normal 6:9
    rad_g, rad_o = gs[:, :3].mean(0), os_[:, :3].mean(0)
    rel = (rad_g - rad_o) / rad_o
    # per-pixel signed differences of the radiance: mean against its own standard error (a lean to one side shows up here
    # long before it moves the image mean)
    diff = (gs[:, :3] - os_[:, :3])
    lean = np.abs(diff.mean(0)) / (diff.std(0) / np.sqrt(diff.shape[0]) + 1e-300)
    alb = np.abs(gs[:, 3:6].mean(0) - os_[:, 3:6].mean(0)).max()
    nrm = np.abs(gs[:, 6:9].mean(0) - os_[:, 6:9].mean(0)).max()
    rays = st["rays"] / o["stats"]["rays"] - 1
    print(f"{cfg} {rp.width}x{rp.height}x{spp}: radiance mean GPU {rad_g} CPU {rad_o} signed rel {rel}, lean (sigmas) {lean}, "
          f"albedo mean diff {alb:.2e}, normal mean diff {nrm:.2e}, rays {rays:+.2e}, pixel |diff| max {np.abs(diff).max():.3e}")
    # observed on B200 (max over the five scenes): |rel| 1.6e-4 (C3b, 20 bounces through glass and water; 8e-5 elsewhere), lean 2.0 sigmas,
    # AOV mean differences 7e-6, ray count 1.5e-5.  The bar asked for is 5e-3; the assertion keeps a 5x margin over what is seen.
    assert np.abs(rel).max() <= 1e-3, rel
    assert lean.max() <= 4.0, lean
    assert alb <= 5e-5 and nrm <= 5e-5
    assert abs(rays) <= 1e-4


@pytest.mark.parametrize("cfg,spp", [("C1", 512), ("C3", 256), ("C2", 16), ("C4", 32), ("C3b", 64)])
def test_noise_floor_vs_reference_rand_stream(rt, ref, cfg, spp):
    sc = S.CONFIGS[cfg]()
    rp = S.params_for(sc, width=64, spp=spp)
    ref1 = ref.render(sc, rp, threads=1, seed=1)["canva"]            # main.c itself, deterministic at one thread
    ref2 = ref.render(sc, rp, threads=1, seed=2)["canva"]
    rt.scene_set(sc)
    gpu = rt.render(rp)["canva"]
    floor = np.sqrt(((ref1 - ref2) ** 2).mean())
    rmse = np.sqrt(((gpu - ref1) ** 2).mean())
    m_ref = 0.5 * (ref1.mean(0) + ref2.mean(0))
    rel = (gpu.mean(0) - m_ref) / np.maximum(m_ref, 1.0)
    spread = np.abs(ref1.mean(0) - ref2.mean(0)) / np.maximum(m_ref, 1.0)      # what two reference runs differ by
    print(f"{cfg} {rp.width}x{rp.height}x{spp}: RMSE(GPU,ref1) {rmse:.3f} vs floor RMSE(ref1,ref2) {floor:.3f} (ratio {rmse / floor:.3f}); "
          f"canva means GPU {gpu.mean(0)} ref {m_ref}, signed rel {rel}, ref1-ref2 spread {spread}")
    assert rmse <= 1.10 * floor, (rmse, floor)          # observed 0.93 .. 1.02; the bar asked for is 1.25
    assert np.all(np.abs(rel) <= np.maximum(0.01, 2.5 * spread)), (rel, spread)


@pytest.mark.parametrize("cfg", ["C1", "C2", "C4"])
def test_fp32_first_hit_ids_vs_fp64_pass_full_resolution(rt, cfg):
    """SURVEY §7 predicted 5-8 flipped ids per 750 k pixels for FP32 on C1 (wall seams of the r = 500 spheres)."""
    sc = S.CONFIGS[cfg]()
    rp = S.params_for(sc, spp=1)
    rt.scene_set(sc)
    exact = rt.primary(rp)                  # FP64, bit-exact against the CPU path (test_gpu_parity)
    fast = rt.primary_fp32(rp)              # what the path tracer's arithmetic finds
    n = exact["id"].shape[0]
    flips = np.nonzero(exact["id"] != fast["id"])[0]
    hit = (exact["id"] >= 0) & (fast["id"] == exact["id"])
    rel_t = np.abs(fast["t"][hit] - exact["t"][hit]) / np.abs(exact["t"][hit])
    print(f"{cfg} {rp.width}x{rp.height}: {len(flips)} of {n} ids differ ({len(flips) / n:.2e}); t rel err max {rel_t.max():.2e} median {np.median(rel_t):.2e}")
    # observed on B200: C1 6 / 750 000 (8.0e-6), C2 4 / 750 000 (5.3e-6), C4 4 / 2 073 600 (1.9e-6)
    assert len(flips) / n <= 1.5e-5, (len(flips), n)
    assert rel_t.max() <= 1e-4
    # a flipped pixel sits on a silhouette: the FP64 ids around it are not all the same
    W = rp.width
    ids2d = exact["id"].reshape(-1, W)
    for p in flips:
        j, i = divmod(int(p), W)
        nb = {int(ids2d[jj, ii]) for jj in (j - 1, j, j + 1) for ii in (i - 1, i, i + 1) if 0 <= jj < ids2d.shape[0] and 0 <= ii < W}
        assert len(nb) > 1, (p, fast["id"][p], nb)


def test_many_spheres_need_more_than_48k_shared_memory(rt, oracle):
    """600 spheres: the sphere table + per-lane sums pass the 48 KB a kernel gets without opting in."""
    rng = np.random.default_rng(5)
    rows = [S.sky_sphere()]
    rows = [((float(x), float(y), float(z)), 0.08, S.material((0.8, 0.6, 0.4), S.BLACK, 0.0, 0.2)) for x, y, z in rng.uniform(-2, 2, (599, 3))] + rows
    sc = S.Scene(name="many_spheres", spheres=S.spheres(rows), triangles=np.zeros(0, S.TRIANGLE), mat_list=np.zeros(0, S.MATERIAL),
                 cam_origin=(0.0, 0.0, 7.0), cam_target=(0.0, 0.0, 0.0), vfov=40.0, ratio=4.0 / 3.0, width=64, spp=4, bounces=3)
    rt.scene_set(sc)
    rp = S.params_for(sc)
    g = rt.render(rp)
    o = oracle.render(sc, rp, rng=RNG_PHILOX)
    w1 = (np.abs(g["canva"] - o["canva"]).max(axis=1) <= 1).mean()
    print(f"600 spheres: within 1 LSB {w1:.4f}, rays {g['stats']['rays'] / o['stats']['rays'] - 1:+.2e}")
    assert w1 >= 0.98                 # observed 0.9909 (599 small glossy spheres: many silhouettes per pixel)
    assert abs(g["stats"]["rays"] / o["stats"]["rays"] - 1) < 5e-3
