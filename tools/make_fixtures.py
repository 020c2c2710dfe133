#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref = /root/reference/main.c
compiled as a library).  Run in the build container, where /root/reference exists:

    python tools/make_fixtures.py            # everything
    python tools/make_fixtures.py scenes     # only the scene fixtures

Two kinds of file are written:

  scene_*.npz    geometry + textures of a reference asset, captured from the reference's own loaders
                 (list_of_mesh, mesh.h:96-218; create_mat_list_mtl, texture.h:175-354) so that tests
                 and bench.py can rebuild the exact `triangle[]` / `material[]` arrays main.c would
                 hold, on a machine where /root/reference does not exist.  Textures are stored as the
                 PPM's integer samples (u8 RGBA, A from the _alpha.ppm); mesh_from_fixture() divides
                 by maxVal like the loader does, and this script asserts the rebuilt array is
                 bit-identical to the loader's.
  golden_*.npz   outputs of the reference's functions on seeded inputs: primary-hit buffers,
                 1-thread rand()-stream renders (deterministic: main.c never calls srand), leaf
                 known-answer vectors.

PPMs stripped from the reference mount (.MISSING_LARGE_BLOBS) are regenerated from the PNGs that are
present, into a scratch directory (never into /root/reference), per README.md:90-94.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.oracle_py import Ref, Oracle, RNG_LIBC  # noqa: E402
from tipe_raytracer_b200 import scene as S  # noqa: E402

REFERENCE = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def write_p3(path, rgb):
    h, w, _ = rgb.shape
    with open(path, "w") as f:
        f.write(f"P3\n{w} {h}\n255\n")
        flat = rgb.reshape(-1, 3)
        f.write("\n".join(" ".join(map(str, px)) for px in flat.tolist()))
        f.write("\n")


def capture_mesh(ref: Ref, obj: str, mtl: str, out_name: str):
    tris, quel, mats, tw, th, nb, names = ref.load_mesh(obj, mtl)
    assert mats is not None, f"textures missing for {obj}: {names}"
    m = mats.reshape(nb, th, tw)
    rgba = np.zeros((nb, th, tw, 4), np.uint8)
    rgb = np.rint(m["diffuseColor"] * 255.0)
    a = np.rint(m["alpha"] * 255.0)
    rgba[..., :3] = rgb.astype(np.uint8)
    rgba[..., 3] = a.astype(np.uint8)
    path = os.path.join(GOLDEN, out_name)
    np.savez_compressed(path, A=tris["A"], B=tris["B"], C=tris["C"], uvA=tris["uvA"], uvB=tris["uvB"], uvC=tris["uvC"],
                        quel_mat=quel.astype(np.int32), tex_rgba=rgba, maxval=np.float64(255.0),
                        source=np.array(os.path.relpath(obj, REFERENCE)))
    # the rebuilt arrays must be what the reference loader produced, bit for bit
    t2, m2, tw2, th2, nb2, q2 = S.mesh_from_fixture(path)
    for f in ("A", "B", "C", "uvA", "uvB", "uvC"):
        assert np.array_equal(t2[f], tris[f]), f
    assert np.array_equal(q2, quel) and (tw2, th2, nb2) == (tw, th, nb)
    assert np.array_equal(m2["diffuseColor"], mats["diffuseColor"]) and np.array_equal(m2["alpha"], mats["alpha"])
    assert np.all(mats["emissionStrength"] == 0.0)
    print(f"{out_name}: {tris.shape[0]} triangles, {nb} materials {tw}x{th}, {os.path.getsize(path)/1024:.0f} KiB")


def make_scenes(ref: Ref):
    capture_mesh(ref, f"{REFERENCE}/model3D/RTX_MAP/nature/mineways_doubleface_tri.obj",
                 f"{REFERENCE}/model3D/RTX_MAP/nature/mineways_doubleface_tri.mtl", "scene_nature.npz")
    capture_mesh(ref, f"{REFERENCE}/model3D/water/mineways_tri.obj",
                 f"{REFERENCE}/model3D/water/mineways_tri.mtl", "scene_water.npz")
    # tree: PPMs are stripped; regenerate from tex.png (RGBA) into a scratch copy laid out as the MTL expects
    from PIL import Image
    scratch = tempfile.mkdtemp(prefix="trt_tree_")
    try:
        src = f"{REFERENCE}/model3D/1st_mc_world/tree"
        for f in ("tree_tri.obj", "tree_tri.mtl"):
            shutil.copy(os.path.join(src, f), os.path.join(scratch, f))
        os.makedirs(os.path.join(scratch, "tex"))
        im = np.array(Image.open(os.path.join(src, "tex.png")).convert("RGBA"))
        write_p3(os.path.join(scratch, "tex", "tree-RGBA.ppm"), im[..., :3])
        write_p3(os.path.join(scratch, "tex", "tree-RGBA_alpha.ppm"), np.repeat(im[..., 3:4], 3, axis=2))
        capture_mesh(ref, os.path.join(scratch, "tree_tri.obj"), os.path.join(scratch, "tree_tri.mtl"), "scene_tree.npz")
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    # pyramide_eau — the scene main.c itself is configured for (main.c:293-322): 34 triangles whose five materials are
    # exactly the ordinals texture.h:71-87 overrides (1 light, 3 glass, 4 water); coordinates reach +-1800.
    # Its five 1500x1227 PPMs are stripped; they are regenerated from the PNGs at 1/4 resolution (nearest) to keep the
    # fixture small — a REDUCED-TEXTURE variant of the author's scene, labelled as such.
    scratch = tempfile.mkdtemp(prefix="trt_pyr_")
    try:
        src = f"{REFERENCE}/model3D/pyramide_eau"
        for f in ("scene.obj", "scene.mtl"):
            shutil.copy(os.path.join(src, f), os.path.join(scratch, f))
        os.makedirs(os.path.join(scratch, "tex"))
        for name in ("murs", "lumiere", "plafond_sol", "vitre", "eau"):
            rgb = Image.open(os.path.join(src, "tex", name + ".png")).convert("RGB")
            alpha = Image.open(os.path.join(src, "tex", name + "_alpha.png")).convert("L")
            size = (rgb.size[0] // 4, rgb.size[1] // 4)
            write_p3(os.path.join(scratch, "tex", name + ".ppm"), np.array(rgb.resize(size, Image.NEAREST)))
            a = np.array(alpha.resize(size, Image.NEAREST))
            write_p3(os.path.join(scratch, "tex", name + "_alpha.ppm"), np.repeat(a[..., None], 3, axis=2))
        capture_mesh(ref, os.path.join(scratch, "scene.obj"), os.path.join(scratch, "scene.mtl"), "scene_pyramide_eau.npz")
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


# -------------------------------------------------------------------------------------------------
def golden_primary(ref: Ref):
    out = {}
    cases = {
        "C1": (S.cornell_spheres(), 200),
        "C2": (S.nature_mesh(), 160),
        "C3": (S.nested_dielectrics(), 200),
        "C4": (S.tree_mesh(), 160),
        "C3b": (S.pyramide_eau(), 160),
    }
    for name, (sc, width) in cases.items():
        rp = S.params_for(sc, width=width, spp=1)
        r = ref.primary(sc, rp)
        out[f"{name}_width"] = np.int32(width)
        out[f"{name}_t"] = r["t"]
        out[f"{name}_id"] = r["id"]
        out[f"{name}_normal"] = r["normal"]
        out[f"{name}_albedo"] = r["albedo"]
        print(f"primary {name}: {rp.width}x{rp.height}, {np.unique(r['id']).size} distinct ids, {(r['id'] < 0).sum()} misses")
    # the six full-resolution vectors quoted in SURVEY.md §4 (C1, 1000x750)
    sc = S.cornell_spheres()
    rp = S.params_for(sc, width=1000, spp=1)
    full = ref.primary(sc, rp)
    px = np.array([[500, 375], [0, 0], [999, 749], [620, 300], [100, 400], [560, 40]], np.int32)
    idx = px[:, 1] * 1000 + px[:, 0]
    out["C1_full_px"] = px
    out["C1_full_t"], out["C1_full_id"] = full["t"][idx], full["id"][idx]
    out["C1_full_normal"], out["C1_full_albedo"] = full["normal"][idx], full["albedo"][idx]
    out["C1_full_id_histogram"] = np.bincount(full["id"] + 1, minlength=12)
    np.savez_compressed(os.path.join(GOLDEN, "golden_primary.npz"), **out)


def golden_renders(ref: Ref):
    """1-thread renders on the default rand() stream (srand(1) == never calling srand)."""
    out = {}
    cases = {
        "C1": (S.cornell_spheres(), 100, 8, 5),
        "C1_noao": (S.cornell_spheres(), 100, 8, 5),
        "C3": (S.nested_dielectrics(), 80, 6, 19),
        "C2": (S.nature_mesh(), 48, 2, 4),
        "C4": (S.tree_mesh(), 48, 2, 6),
        "C3b": (S.pyramide_eau(), 64, 4, 20),
    }
    for name, (sc, width, spp, bounces) in cases.items():
        if name.endswith("_noao"):
            sc.use_ao = False
        rp = S.params_for(sc, width=width, spp=spp, bounces=bounces)
        r = ref.render(sc, rp, threads=1, seed=1, count_rays=True)
        out[f"{name}_cfg"] = np.array([width, spp, bounces, int(sc.use_ao)], np.int32)
        out[f"{name}_canva"] = r["canva"].astype(np.uint8)
        assert np.array_equal(out[f"{name}_canva"].astype(np.float64), r["canva"])
        out[f"{name}_albedo"], out[f"{name}_normal"] = r["albedo"], r["normal"]
        out[f"{name}_rays"] = np.int64(r["stats"]["rays"])
        print(f"render {name}: {rp.width}x{rp.height}x{spp}, rays/path {r['stats']['rays']/r['stats']['paths']:.3f}, "
              f"mean rgb {r['canva'].mean(0).round(2)}, {r['stats']['seconds']:.1f}s")
    np.savez_compressed(os.path.join(GOLDEN, "golden_render_libc.npz"), **out)


def golden_leaves(ref: Ref):
    rng = np.random.default_rng(20261018)
    out = {}
    n = 256
    # hit_sphere: mix of outside / inside / grazing / behind
    centers = rng.uniform(-2, 2, (n, 3)); radii = rng.uniform(0.1, 2.0, n)
    centers[:8] = [[-501, 0, 0], [0, -501, 0], [501, 0, 0], [0, 0, -504], [0, 501, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0]]
    radii[:8] = [500, 500, 500, 500, 500, 100000, 100000, 1]
    origins = rng.uniform(-3, 3, (n, 3))
    origins[32:64] = centers[32:64] + rng.uniform(-0.5, 0.5, (32, 3)) * radii[32:64, None]   # inside: far root
    aim = rng.normal(size=(n, 3)); aim /= np.linalg.norm(aim, axis=1, keepdims=True)
    dirs = centers + aim * radii[:, None] * rng.uniform(0, 1.4, (n, 1)) - origins            # ~70 % hit
    dirs[200:] = -dirs[200:]                                                                  # sphere behind
    dirs[::2] /= np.linalg.norm(dirs[::2], axis=1, keepdims=True)          # half unit, half un-normalised
    res = [ref.hit_sphere(centers[k], radii[k], origins[k], dirs[k]) for k in range(n)]
    out["sph_center"], out["sph_radius"], out["sph_o"], out["sph_d"] = centers, radii, origins, dirs
    out["sph_hit"] = np.array([r[0] for r in res]); out["sph_dst"] = np.array([r[1] for r in res])
    out["sph_point"] = np.array([r[2] for r in res]); out["sph_normal"] = np.array([r[3] for r in res])
    # hit_triangle: rays aimed at (and around) random triangles, both facings
    tris = np.zeros(n, S.TRIANGLE)
    tris["A"], tris["B"], tris["C"] = (rng.uniform(-1, 1, (n, 3)) for _ in range(3))
    bary = rng.uniform(-0.2, 1.2, (n, 2))
    target = tris["A"] + bary[:, :1] * (tris["B"] - tris["A"]) + bary[:, 1:] * (tris["C"] - tris["A"])
    o = rng.uniform(-3, 3, (n, 3)); d = target - o
    d[::2] /= np.linalg.norm(d[::2], axis=1, keepdims=True)
    res = [ref.hit_triangle(tris[k], o[k], d[k]) for k in range(n)]
    out["tri_A"], out["tri_B"], out["tri_C"], out["tri_o"], out["tri_d"] = tris["A"], tris["B"], tris["C"], o, d
    out["tri_hit"] = np.array([r[0] for r in res]); out["tri_dst"] = np.array([r[1] for r in res])
    out["tri_point"] = np.array([r[2] for r in res]); out["tri_normal"] = np.array([r[3] for r in res])
    # reflect / refract (incl. total internal reflection)
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    nn = rng.normal(size=(n, 3)); nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    n1 = rng.choice([1.0, 1.33, 1.5], n); n2 = rng.choice([1.0, 1.33, 1.5], n)
    out["rf_v"], out["rf_n"], out["rf_n1"], out["rf_n2"] = v, nn, n1, n2
    out["rf_reflect"] = np.array([ref.reflect(v[k], nn[k]) for k in range(n)])
    out["rf_refract"] = np.array([ref.refract(v[k], nn[k], n1[k], n2[k]) for k in range(n)])
    # tonemap around the truncation boundaries
    spp = 16
    lv = np.concatenate([np.linspace(0, 1.2, 200), ((np.arange(1, 57) / 256.0) ** 2)]) * spp
    sums = np.stack([lv, lv * 0.5, lv * 2.0], axis=1)
    out["tm_sum"], out["tm_spp"] = sums, np.int32(spp)
    out["tm_out"] = np.array([ref.tonemap(s, spp) for s in sums])
    # HSL round trip (identity to rounding)
    rgb = rng.uniform(0, 1, (64, 3)); rgb[:4] = [[1, 1, 1], [1, 0.6, 0.2], [0.784, 0.965, 1], [0, 0, 0]]
    out["hsl_in"] = rgb; out["hsl_out"] = np.array([ref.hsl_roundtrip(c) for c in rgb])
    # refraction stack protocol
    kinds = rng.integers(0, 2, 64).astype(np.int32); vals = rng.choice([1.0, 1.33, 1.5], 64)
    out["pile_kind"], out["pile_value"], out["pile_out"] = kinds, vals, ref.pile(kinds, vals)
    # camera rays with aperture
    sc = S.cornell_spheres(); cam = sc.camera()
    uv = rng.uniform(0, 1, (64, 2)); dxy = rng.uniform(-0.5, 0.5, (64, 2)) * 2.0
    rays = [ref.get_ray(uv[k, 0], uv[k, 1], cam, 3.0, dxy[k, 0], dxy[k, 1]) for k in range(64)]
    out["cam"] = np.ascontiguousarray(cam).view(np.float64).reshape(12)
    out["cam_uv"], out["cam_dxy"] = uv, dxy
    out["cam_o"] = np.array([r[0] for r in rays]); out["cam_d"] = np.array([r[1] for r in rays])
    # texture lookups on the nature mesh: reference hit point/normal -> material
    nat = S.nature_mesh()
    ids = rng.integers(0, nat.triangles.shape[0], 128)
    pts, nrm, mats = [], [], []
    for k in ids:
        t = nat.triangles[k]
        b = rng.dirichlet([1, 1, 1])
        p = b[0] * t["A"] + b[1] * t["B"] + b[2] * t["C"]
        nv = np.cross(t["B"] - t["A"], t["C"] - t["A"]); nv /= np.linalg.norm(nv)
        pts.append(p); nrm.append(nv); mats.append(ref.tri_uvmapping(nat, int(k), p, nv))
    out["uv_tri"], out["uv_point"], out["uv_normal"] = ids.astype(np.int32), np.array(pts), np.array(nrm)
    m = np.array(mats, S.MATERIAL)
    out["uv_diffuse"], out["uv_alpha"] = m["diffuseColor"], m["alpha"]
    out["uv_emissionStrength"], out["uv_materialIndex"] = m["emissionStrength"], m["materialIndex"]
    np.savez_compressed(os.path.join(GOLDEN, "golden_leaves.npz"), **out)
    print("leaves: written")


def golden_sky(ref):
    """sphere_uvmapping (texture.h:92-112): hit points on a sphere -> the sky texel the reference reads."""
    rng = np.random.default_rng(77)
    out = {}
    center = np.array([0.5, -2.0, 3.0]); radius = 100000.0; w, h = 256, 128
    d = rng.normal(size=(512, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    d[:6] = [[1, 0, 0], [-1, 1e-9, 1e-9], [0, 0, 1], [0, 0, -1], [0, 0.999999, 0.001], [0.001, -0.999, 0.001]]
    d /= np.linalg.norm(d, axis=1)[:, None]
    pts = center + radius * d
    out["center"], out["radius"], out["size"], out["points"] = center, np.float64(radius), np.array([w, h]), pts
    out["index"] = np.array([ref.sphere_uvmapping(center, radius, p, w, h) for p in pts], np.int64)
    np.savez_compressed(os.path.join(GOLDEN, "golden_sky.npz"), **out)
    print("sky: written", out["index"][:8])


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = Ref()
    what = sys.argv[1:] or ["scenes", "primary", "renders", "leaves", "sky"]
    if "scenes" in what:
        make_scenes(ref)
    if "primary" in what:
        golden_primary(ref)
    if "renders" in what:
        golden_renders(ref)
    if "leaves" in what:
        golden_leaves(ref)
    if "sky" in what:
        golden_sky(ref)


if __name__ == "__main__":
    main()
