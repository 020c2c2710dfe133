"""Pins the CPU oracle (oracle/tipe_oracle.c) to the reference.

Golden vectors under tests/golden/ were produced by tools/make_fixtures.py from oracle/_ref, i.e.
from /root/reference/main.c itself.  In libc-RNG mode the restatement must reproduce them BIT FOR
BIT (FP64, same operation order, same rand() call order).  Where oracle/_ref is present the same
comparisons are repeated live on fresh seeds.
"""
import ctypes

import numpy as np
import pytest

from oracle.oracle_py import RNG_LIBC, RNG_PHILOX
from tipe_raytracer_b200 import scene as S

libc = ctypes.CDLL(None)


def _scene(name):
    base = name.split("_")[0]          # "C1_noao" -> C1
    sc = S.CONFIGS[base]()
    if name.endswith("_noao"):
        sc.use_ao = False
    return sc


# ---- leaves ---------------------------------------------------------------------------------
def test_struct_layout_matches_reference(ref):
    # vec3, ray, material, HitInfo, sphere, UV, triangle, camera, col_alb_norm, off(uvA), off(sphere.mat), off(alpha)
    assert ref.sizes() == [24, 48, 80, 144, 112, 16, 200, 96, 72, 152, 32, 64]
    assert S.MATERIAL.itemsize == 80 and S.SPHERE.itemsize == 112 and S.TRIANGLE.itemsize == 200
    assert S.TRIANGLE.fields["uvA"][1] == 152 and S.SPHERE.fields["mat"][1] == 32


def test_hit_sphere_golden(oracle, golden):
    g = golden["leaves"]
    n = g["sph_hit"].shape[0]
    assert 0.2 < g["sph_hit"].mean() < 0.9
    for k in range(n):
        hit, dst, pt, nr = oracle.hit_sphere(g["sph_center"][k], g["sph_radius"][k], g["sph_o"][k], g["sph_d"][k])
        assert hit == bool(g["sph_hit"][k])
        if hit:
            assert dst == g["sph_dst"][k] and np.array_equal(pt, g["sph_point"][k]) and np.array_equal(nr, g["sph_normal"][k])


def test_hit_triangle_golden(oracle, golden):
    g = golden["leaves"]
    n = g["tri_hit"].shape[0]
    assert 0.1 < g["tri_hit"].mean() < 0.9
    tri = np.zeros(1, S.TRIANGLE)
    for k in range(n):
        tri["A"], tri["B"], tri["C"] = g["tri_A"][k], g["tri_B"][k], g["tri_C"][k]
        hit, dst, pt, nr = oracle.hit_triangle(tri, g["tri_o"][k], g["tri_d"][k])
        assert hit == bool(g["tri_hit"][k]) and dst == g["tri_dst"][k]
        assert np.array_equal(pt, g["tri_point"][k]) and np.array_equal(nr, g["tri_normal"][k])


def test_reflect_refract_golden(oracle, golden):
    g = golden["leaves"]
    for k in range(g["rf_v"].shape[0]):
        assert np.array_equal(oracle.reflect(g["rf_v"][k], g["rf_n"][k]), g["rf_reflect"][k])
        assert np.array_equal(oracle.refract(g["rf_v"][k], g["rf_n"][k], g["rf_n1"][k], g["rf_n2"][k]), g["rf_refract"][k])


def test_tonemap_hsl_pile_camera_golden(oracle, golden):
    g = golden["leaves"]
    spp = int(g["tm_spp"])
    for s, want in zip(g["tm_sum"], g["tm_out"]):
        assert np.array_equal(oracle.tonemap(s, spp), want)
    for c, want in zip(g["hsl_in"], g["hsl_out"]):
        assert np.array_equal(oracle.hsl_roundtrip(c), want)
        assert np.abs(want - c).max() < 1e-14              # the round trip is the identity
    assert np.array_equal(oracle.pile(g["pile_kind"], g["pile_value"]), g["pile_out"])
    cam = g["cam"]
    for k in range(g["cam_uv"].shape[0]):
        o, d = oracle.get_ray(g["cam_uv"][k, 0], g["cam_uv"][k, 1], cam, 3.0, g["cam_dxy"][k, 0], g["cam_dxy"][k, 1])
        assert np.array_equal(o, g["cam_o"][k]) and np.array_equal(d, g["cam_d"][k])


def test_pile_is_a_scalar(golden):
    """the linked-list stack is equivalent to one scalar `cur` — the form the GPU uses."""
    g = golden["leaves"]
    cur = 1.0
    for kind, m, (n1, n2) in zip(g["pile_kind"], g["pile_value"], g["pile_out"]):
        if kind == 0:                       # entering: (n1, n2) = (cur, m), cur <- m
            assert (n1, n2) == (cur, m)
            cur = m
        else:                               # leaving: (n1, n2) = (m, cur), cur unchanged
            assert (n1, n2) == (m, cur)


def test_tri_uvmapping_golden(oracle, golden):
    g = golden["leaves"]
    nat = S.nature_mesh()
    for k in range(g["uv_tri"].shape[0]):
        m = oracle.tri_uvmapping(nat, int(g["uv_tri"][k]), g["uv_point"][k], g["uv_normal"][k])
        assert np.array_equal(m["diffuseColor"], g["uv_diffuse"][k]) and m["alpha"] == g["uv_alpha"][k]
        assert m["emissionStrength"] == g["uv_emissionStrength"][k] and m["materialIndex"] == g["uv_materialIndex"][k]
    assert set(np.unique(g["uv_alpha"])) >= {0.0, 1.0}     # cut-outs and opaque texels both present


def test_sphere_uvmapping_golden(oracle, golden):
    """Sky lookup: the texel index of sphere_uvmapping (texture.h:92-112), pinned against the
    function compiled from the reference; its call site (main.c:64-71) is commented out upstream, so this row is
    pinned at function level and the image-level checks run against the restatement."""
    g = golden["sky"]
    w, h = (int(x) for x in g["size"])
    for p, want in zip(g["points"], g["index"]):
        assert oracle.sphere_uvmapping(g["center"], float(g["radius"]), p, w, h) == int(want)
    assert len(np.unique(g["index"])) > 400                  # the vectors really spread over the image


def test_sphere_uvmapping_live(oracle, ref):
    rng = np.random.default_rng(5)
    for _ in range(200):
        c = rng.uniform(-3, 3, 3); r = float(rng.uniform(0.5, 1e5))
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        p = c + r * d
        assert oracle.sphere_uvmapping(c, r, p, 512, 256) == ref.sphere_uvmapping(c, r, p, 512, 256)


def test_sky_changes_only_the_last_sphere(oracle):
    """main.c:64-71: emissionColor of the LAST sphere comes from the sky texel, everything else is untouched;
    sky_mat_list = None is main.c as shipped (bit-identical to the golden renders, test_render_libc_golden)."""
    sc = S.tree_mesh_sky()
    rp = S.params_for(sc, width=48, spp=2)
    with_sky = oracle.render(sc, rp, rng=RNG_PHILOX, want_sums=True)
    sc2 = S.tree_mesh()
    without = oracle.render(sc2, rp, rng=RNG_PHILOX, want_sums=True)
    a, b = with_sky["sums"], without["sums"]
    assert np.array_equal(a[:, 6:], b[:, 6:])                # geometry (normal AOV) is unchanged
    assert not np.array_equal(a[:, :3], b[:, :3])            # radiance and, for directly seen sky (main.c:154-160), the albedo AOV change
    assert not np.array_equal(a[:, 3:6], b[:, 3:6])
    # a uniform sky of the sphere's own emission colour is the same image as no sky at all
    sc3 = S.tree_mesh_sky()
    sc3.sky_mat_list = sc3.sky_mat_list.copy()
    sc3.sky_mat_list["diffuseColor"] = sc3.spheres[-1]["mat"]["emissionColor"]
    same = oracle.render(sc3, rp, rng=RNG_PHILOX, want_sums=True)
    assert np.array_equal(same["sums"], b)


def test_init_camera_matches_survey_vectors(oracle):
    sc = S.cornell_spheres()
    cam = np.ascontiguousarray(sc.camera()).view(np.float64).reshape(12)
    assert np.array_equal(cam, oracle.init_camera(sc.cam_origin, sc.cam_target, sc.cam_up, sc.vfov, sc.ratio))
    # SURVEY.md §4 known-answer vectors (printed with %.17g from the reference)
    assert np.allclose(cam[3:6], [1.8584717400487587, 0, -0.18053725474759372], rtol=0, atol=1e-15)
    assert np.allclose(cam[6:9], [-0.030036746433818585, 1.3655234989045113, -0.30920180152460308], rtol=0, atol=1e-15)
    assert np.allclose(cam[9:12], [-0.66849622616424687, -0.60459405382114229, -0.22564680347778088], rtol=0, atol=1e-15)


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert [int(x) for x in oracle.philox(ctr, key)] == want


# ---- primary-hit buffers ------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", ["C1", "C2", "C3", "C4", "C3b"])
def test_primary_golden(oracle, golden, cfg):
    g = golden["primary"]
    sc = S.CONFIGS[cfg]()
    rp = S.params_for(sc, width=int(g[f"{cfg}_width"]), spp=1)
    r = oracle.primary(sc, rp)
    assert np.array_equal(r["id"], g[f"{cfg}_id"])
    assert np.array_equal(r["t"], g[f"{cfg}_t"])
    assert np.array_equal(r["normal"], g[f"{cfg}_normal"]) and np.array_equal(r["albedo"], g[f"{cfg}_albedo"])


def test_primary_survey_vectors(golden):
    """The six full-resolution vectors of SURVEY.md §4 are what the fixture generator captured."""
    g = golden["primary"]
    assert g["C1_full_id"].tolist() == [9, 0, 2, 1, 0, 5]
    assert np.allclose(g["C1_full_t"], [3.60001851597, 2.0436055891, 1.23834747912, 3.78845628151, 1.99725590006, 1.83713044576], rtol=1e-11)
    assert np.allclose(g["C1_full_normal"][0], [-0.792256895, 0.009363881, 0.610115833], atol=1e-9)
    assert np.allclose(g["C1_full_albedo"][0], [0.784, 0.965, 1.0])


# ---- full renders on the rand() stream ---------------------------------------------------------
@pytest.mark.parametrize("cfg", ["C1", "C1_noao", "C3", "C2", "C4", "C3b"])
def test_render_libc_golden(oracle, golden, cfg):
    g = golden["render_libc"]
    width, spp, bounces, use_ao = (int(x) for x in g[f"{cfg}_cfg"])
    sc = _scene(cfg)
    assert sc.use_ao == bool(use_ao)
    rp = S.params_for(sc, width=width, spp=spp, bounces=bounces)
    libc.srand(1)
    r = oracle.render(sc, rp, rng=RNG_LIBC, threads=1)
    assert np.array_equal(r["canva"], g[f"{cfg}_canva"].astype(np.float64))
    assert np.array_equal(r["albedo"], g[f"{cfg}_albedo"]) and np.array_equal(r["normal"], g[f"{cfg}_normal"])
    assert r["stats"]["rays"] == int(g[f"{cfg}_rays"])


def test_render_live_against_reference(oracle, ref):
    """Fresh seeds, live: reference binary vs restatement, bit for bit, incl. the ray count."""
    for cfg, width, spp, seed in (("C1", 64, 5, 12345), ("C3", 48, 4, 777)):
        sc = S.CONFIGS[cfg]()
        rp = S.params_for(sc, width=width, spp=spp)
        a = ref.render(sc, rp, threads=1, seed=seed, count_rays=True)
        libc.srand(seed)
        b = oracle.render(sc, rp, rng=RNG_LIBC, threads=1)
        for k in ("canva", "albedo", "normal"):
            assert np.array_equal(a[k], b[k]), (cfg, k)
        assert a["stats"]["rays"] == b["stats"]["rays"]


def test_survey_radiometric_pin(oracle):
    """C1 at 200x150x20spp, 1 thread, default rand() stream -> mean RGB (71.3, 60.5, 36.9) with AO."""
    sc = S.cornell_spheres()
    rp = S.params_for(sc, width=200, spp=20)
    libc.srand(1)
    r = oracle.render(sc, rp, rng=RNG_LIBC, threads=1)
    assert np.allclose(r["canva"].mean(0), [71.3, 60.5, 36.9], atol=0.06)
    assert abs(r["stats"]["rays"] / r["stats"]["paths"] - 9.534) < 0.01


# ---- Philox mode: statistically the same renderer --------------------------------------------------
def test_philox_mode_is_thread_invariant_and_matches_libc_statistics(oracle):
    sc = S.cornell_spheres()
    rp = S.params_for(sc, width=96, spp=24)
    a = oracle.render(sc, rp, rng=RNG_PHILOX, threads=1)
    b = oracle.render(sc, rp, rng=RNG_PHILOX, threads=5)
    assert np.array_equal(a["canva"], b["canva"]) and a["stats"]["rays"] == b["stats"]["rays"]
    libc.srand(3)
    c = oracle.render(sc, rp, rng=RNG_LIBC, threads=1)
    assert np.abs(a["canva"].mean(0) - c["canva"].mean(0)).max() < 1.5
    assert abs(a["stats"]["rays"] / c["stats"]["rays"] - 1) < 0.01
    # sample-range split: two half renders add up to the full one (what the spp split relies on)
    rp1 = S.params_for(sc, width=96, spp=12, sample_total=24)
    rp2 = S.params_for(sc, width=96, spp=12, sample_offset=12, sample_total=24)
    s1 = oracle.render(sc, rp1, rng=RNG_PHILOX, want_sums=True)["sums"]
    s2 = oracle.render(sc, rp2, rng=RNG_PHILOX, want_sums=True)["sums"]
    full = oracle.render(sc, rp, rng=RNG_PHILOX, want_sums=True)["sums"]
    assert np.allclose(s1 + s2, full, rtol=1e-12, atol=1e-12)
