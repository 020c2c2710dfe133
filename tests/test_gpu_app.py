"""GPU tests of the reference-facing entry points: the plain-C host program (apps/main_b200.c), the
fill_canva(ThreadData) mirror, in-process multi-GPU, and the mini OBJ/MTL/PPM scene that exercises
texture overrides (light / glass / water ordinals of texture.h:71-87) and alpha cut-outs."""
import dataclasses
import os
import subprocess

import numpy as np
import pytest

import tipe_raytracer_b200 as trt
from oracle.oracle_py import RNG_PHILOX
from tipe_raytracer_b200 import assets
from tipe_raytracer_b200 import scene as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINI = os.path.join(ROOT, "tests", "golden", "mini_scene")


def mini_scene():
    tris, quel, paths = assets.load_mesh(os.path.join(MINI, "mini.obj"), os.path.join(MINI, "mini.mtl"))
    mats, tw, th = assets.load_textures(paths)
    return S.Scene(name="mini", spheres=S.spheres([S.sky_sphere()]), triangles=tris, mat_list=mats, tex_w=tw, tex_h=th,
                   nb_materials=len(paths), quel_mat_pour_tri=quel, cam_origin=(2.2, 1.6, 3.0), cam_target=(0.0, 0.5, 0.0),
                   vfov=45.0, ratio=4.0 / 3.0, width=160, spp=16, bounces=6, use_ao=False)


def read_ppm(path):
    tok = open(path).read().split()
    w, h = int(tok[1]), int(tok[2])
    img = np.array(tok[4:], np.float64).reshape(h, w, 3)
    return img[::-1].reshape(-1, 3), w, h          # file is written top row first; pixel_index has j=0 at the bottom


def test_mini_scene_matches_oracle_same_stream(rt, oracle):
    sc = mini_scene()
    rp = S.params_for(sc)
    rt.scene_set(sc)
    g = rt.render(rp)
    o = oracle.render(sc, rp, rng=RNG_PHILOX)
    d = np.abs(g["canva"] - o["canva"])
    print(f"mini scene vs oracle: within 1 LSB {(d.max(axis=1) <= 1).mean():.4f}, mean |diff| {d.mean():.4f}")
    assert (d.max(axis=1) <= 1).mean() > 0.995 and d.mean() < 0.05, ((d.max(axis=1) <= 1).mean(), d.mean())     # observed 1.0 / 0.0
    assert abs(g["stats"]["rays"] / o["stats"]["rays"] - 1) < 5e-3
    prim_g, prim_o = rt.primary(rp), oracle.primary(sc, rp)
    assert np.array_equal(prim_g["id"], prim_o["id"]) and np.array_equal(prim_g["t"], prim_o["t"])
    assert rt.bvh_info()["n_twins"] >= 4            # the double-faced glass pane and plant quads


def test_c_host_program_renders_the_same_image(rt, tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "apps"), "main_b200"], stderr=subprocess.DEVNULL)
    exe = os.path.join(ROOT, "apps", "main_b200")
    # README scene (spheres, AO on, main.c's int truncation of AO_intensity 2.5 -> 2)
    out = str(tmp_path / "readme.ppm")
    subprocess.check_call([exe, "--width", "120", "--spp", "8", "--out", out], stdout=subprocess.DEVNULL)
    img, w, h = read_ppm(out)
    sc = S.cornell_spheres()
    rt.scene_set(sc)
    want = rt.render(S.params_for(sc, width=120, spp=8))["canva"]
    assert (w, h) == (120, 90) and np.array_equal(img, want)
    # mesh from OBJ/MTL/PPM files through the C loaders
    out = str(tmp_path / "mini.ppm")
    sc = mini_scene()
    subprocess.check_call([exe, "--scene", "sky", "--obj", os.path.join(MINI, "mini.obj"), "--mtl", os.path.join(MINI, "mini.mtl"),
                           "--width", "160", "--spp", "16", "--bounces", "6", "--ao", "0", "--vfov", "45",
                           "--origin", "2.2,1.6,3.0", "--target", "0.0,0.5,0.0", "--out", out], stdout=subprocess.DEVNULL)
    img, w, h = read_ppm(out)
    rt.scene_set(sc)
    assert np.array_equal(img, rt.render(S.params_for(sc))["canva"])
    # sky image through the C loader (create_mat_list's layout) and trt_sky_set
    sky_sc = mini_scene()
    sky_sc.sky_w, sky_sc.sky_h = 64, 32
    sky_sc.sky_mat_list = S.procedural_sky(64, 32)
    rows = (sky_sc.sky_mat_list["diffuseColor"].reshape(32, 64, 3) * 255.0 + 0.5).astype(int)[::-1]      # file row 0 = top
    sky_ppm = tmp_path / "sky.ppm"
    sky_ppm.write_text("P3\n64 32\n255\n" + "\n".join(" ".join(str(v) for v in px) for row in rows for px in row) + "\n")
    out = str(tmp_path / "mini_sky.ppm")
    subprocess.check_call([exe, "--scene", "sky", "--obj", os.path.join(MINI, "mini.obj"), "--mtl", os.path.join(MINI, "mini.mtl"),
                           "--width", "160", "--spp", "16", "--bounces", "6", "--ao", "0", "--vfov", "45", "--sky-image", str(sky_ppm),
                           "--origin", "2.2,1.6,3.0", "--target", "0.0,0.5,0.0", "--out", out], stdout=subprocess.DEVNULL)
    img_sky, w, h = read_ppm(out)
    rt.scene_set(sky_sc)
    assert np.array_equal(img_sky, rt.render(S.params_for(sky_sc))["canva"]) and not np.array_equal(img_sky, img)
    # the overrides of texture.h:71-87 as data: a scene file that drops them and sets one of its own == the same table through the Python mirror
    cfg = tmp_path / "mini.cfg"
    cfg.write_text(f"scene = sky\nobj = {os.path.join(MINI, 'mini.obj')}\nmtl = {os.path.join(MINI, 'mini.mtl')}\nwidth = 160\nspp = 16\nbounces = 6\n"
                   "ao = 0\nvfov = 45\norigin = 2.2,1.6,3.0\ntarget = 0.0,0.5,0.0\nno-overrides\noverride = 3:alpha=0.3,index=1.4,reflection=0.5\n")
    out = str(tmp_path / "mini_ov.ppm")
    subprocess.check_call([exe, "--config", str(cfg), "--out", out], stdout=subprocess.DEVNULL)
    img_ov, w, h = read_ppm(out)
    ov_sc = mini_scene()
    ov_sc.overrides = np.zeros(1, S.MAT_OVERRIDE)
    ov_sc.overrides[0] = (3, S.OV_ALPHA | S.OV_MATERIAL_INDEX | S.OV_REFLECTION, S.BLACK, 0.0, 0.5, 0.3, 1.4)
    rt.scene_set(ov_sc)
    assert np.array_equal(img_ov, rt.render(S.params_for(ov_sc))["canva"]) and not np.array_equal(img_ov, img)
    # failure is reported, not a crash: missing texture
    bad = subprocess.run([exe, "--scene", "sky", "--obj", os.path.join(MINI, "mini.obj"), "--mtl", "/nonexistent.mtl"], capture_output=True, text=True)
    assert bad.returncode == 1 and "cannot open" in bad.stderr


def test_fill_canva_thread_data_mirror(rt):
    sc = S.cornell_spheres()
    rp = S.params_for(sc, width=80, spp=6)
    n = rp.width * rp.height
    full = {k: np.zeros((n, 3)) for k in ("canva", "albedo", "normal")}
    td = trt.ThreadData(start_row=rp.height - 1, end_row=0, canva=full["canva"], albedo_tab=full["albedo"], normal_tab=full["normal"],
                        mat_list=sc.mat_list, cam=rp.cam, largeur_image=rp.width, hauteur_image=rp.height, tex_width=0, tex_height=0,
                        quelMatPourTri=sc.quel_mat_pour_tri, nbRayonParPixel=6, nbRebondMax=5, sphere_list=sc.spheres,
                        triangle_list=sc.triangles, nbSpheres=10, nbTriangles=0, AO_intensity=2, useAO=True)
    trt.fill_canva(td, renderer=rt)
    rt.scene_set(sc)
    want = rt.render(rp)
    assert np.array_equal(full["canva"], want["canva"]) and np.array_equal(full["normal"], want["normal"])
    # a row band, like one pthread of main.c:407-449: rows 10..29 only
    band = {k: np.zeros((n, 3)) for k in ("canva", "albedo", "normal")}
    td.start_row, td.end_row = 29, 10
    td.canva, td.albedo_tab, td.normal_tab = band["canva"], band["albedo"], band["normal"]
    trt.fill_canva(td, renderer=rt)
    lo, hi = 10 * rp.width, 30 * rp.width
    assert np.array_equal(band["canva"][lo:hi], want["canva"][lo:hi]) and not band["canva"][:lo].any() and not band["canva"][hi:].any()


def test_slot_pool_kernel_is_bit_identical_to_the_register_kernel(rt):
    """pathtrace.cu (engine 1, a lane owns a path) and pathtrace_pool.cu (engine 3, a warp owns 64 path slots in shared memory): same image."""
    sc = S.nature_mesh()
    rt.scene_set(sc)
    rp = S.params_for(sc, width=200, spp=12)
    try:
        os.environ["TRT_ENGINE"] = "1"
        a = rt.render(rp)
        os.environ["TRT_ENGINE"] = "3"
        b = rt.render(rp)
        sc2 = S.nature_mesh(); sc2.use_ao = True                         # AO rays: the parked bounce direction of the slot record
        rt.scene_set(sc2)
        rp2 = S.params_for(sc2, width=97, spp=5)
        d = rt.render(rp2)
        os.environ["TRT_ENGINE"] = "1"
        c = rt.render(rp2)
    finally:
        os.environ.pop("TRT_ENGINE", None)
    for k in ("canva", "albedo", "normal"):
        assert a[k].tobytes() == b[k].tobytes() and c[k].tobytes() == d[k].tobytes()
    assert a["stats"]["rays"] == b["stats"]["rays"] and c["stats"]["rays"] == d["stats"]["rays"]


def test_packet_primaries_are_bit_identical_to_the_register_kernel(rt):
    """Engine 4: the primary rays of the frame traced as 32-ray packets with one warp-shared stack (primary_packet.cu), the
    path-tracing kernel starting every path from its stored hit.  Same image, same ray count; ragged size, AO on a mesh too."""
    for use_ao, width, spp in ((False, 200, 12), (True, 97, 5)):
        sc = S.nature_mesh(); sc.use_ao = use_ao
        rt.scene_set(sc)
        rp = S.params_for(sc, width=width, spp=spp)
        try:
            os.environ["TRT_ENGINE"] = "1"
            a = rt.render(rp)
            os.environ["TRT_ENGINE"] = "4"
            b = rt.render(rp)
        finally:
            os.environ.pop("TRT_ENGINE", None)
        for k in ("canva", "albedo", "normal"):
            assert a[k].tobytes() == b[k].tobytes()
        assert a["stats"]["rays"] == b["stats"]["rays"] and b["stats"]["kernel_launches"] == a["stats"]["kernel_launches"] + 1


def test_in_process_multi_gpu_is_bit_identical():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sc = S.cornell_spheres()
    rp = S.params_for(sc, width=160, spp=21)
    one = trt.Renderer([0]); one.scene_set(sc); a = one.render(rp); one.close()
    two = trt.Renderer([0, 1]); two.scene_set(sc); b = two.render(rp); two.close()
    for k in ("canva", "albedo", "normal"):
        assert a[k].tobytes() == b[k].tobytes()
    assert a["stats"]["rays"] == b["stats"]["rays"]
    # tile split (SURVEY §8e partitioning B): every device renders all samples of the tiles k mod N
    two = trt.Renderer([0, 1]); two.scene_set(sc)
    c = two.render(S.params_for(sc, width=160, spp=21, flags=trt.FLAG_TILE_SPLIT)); two.close()
    for k in ("canva", "albedo", "normal"):
        assert a[k].tobytes() == c[k].tobytes()
    assert a["stats"]["rays"] == c["stats"]["rays"]


def test_tile_parts_assemble_the_image(rt):
    """trt_params.tile_part: the interleaved-tile parts of an image, rendered by separate calls (separate processes / GPUs in the
    multi-process path) into one sum buffer, are the image of a single call, bit for bit; ragged size, mesh and sphere scenes."""
    import torch
    from tipe_raytracer_b200 import distributed as D
    for sc, width, spp in ((S.cornell_spheres(), 150, 9), (S.nature_mesh(), 97, 5)):
        rt.scene_set(sc)
        rp = S.params_for(sc, width=width, spp=spp)
        want = rt.render(rp)
        n = rp.width * rp.height
        acc = torch.zeros(9 * n, dtype=torch.int64, device="cuda")
        rays = 0
        for part in range(3):
            rays += rt.accumulate(D.rank_params(rp, part, 3, split="tile"), acc.data_ptr())["rays"]
        got = rt.resolve(rp, acc.data_ptr())
        for k in ("canva", "albedo", "normal"):
            assert got[k].tobytes() == want[k].tobytes()
        assert rays == want["stats"]["rays"]
        one_part = torch.zeros(9 * n, dtype=torch.int64, device="cuda")
        rt.accumulate(D.rank_params(rp, 1, 3, split="tile"), one_part.data_ptr())
        touched = (one_part.view(9, n) != 0).any(0).float().mean().item()
        assert 0.2 < touched < 0.5                                   # a third of the tiles (sky / black pixels stay zero anyway)
    with pytest.raises(trt.TrtError):
        rt.accumulate(dataclasses.replace(rp, tile_part=5 | (3 << 16)), acc.data_ptr())


def test_render_distributed_single_rank_equals_render(rt):
    """trt_render_distributed with one rank (no communicator) is trt_render; its stats describe the rank's share."""
    sc = S.cornell_spheres()
    rt.scene_set(sc)
    rp = S.params_for(sc, width=120, spp=9)
    a = rt.render(rp)
    b = rt.render_distributed(rp, 0, 0, 1)
    for k in ("canva", "albedo", "normal"):
        assert a[k].tobytes() == b[k].tobytes()
    assert a["stats"]["rays"] == b["stats"]["rays"] and b["stats"]["paths"] == rp.width * rp.height * rp.spp
    # the AO ray of a path's last bounce is counted like the reference counts it, but not traced (trt_stats.rays_skipped)
    assert 0 < a["stats"]["rays_skipped"] <= a["stats"]["paths"] and a["stats"]["rays_skipped"] < a["stats"]["rays"]
    with pytest.raises(trt.TrtError):
        rt.render_distributed(rp, 0, 1, 2)                      # two ranks need a communicator
    assert rt.lib.trt_nccl_version() >= 20000                    # NCCL is found at run time (libnccl.so.2)


def test_c_host_program_one_process_per_gpu(rt, tmp_path):
    """apps/main_b200 --nranks 2: two PROCESSES, one GPU each, the NCCL id handed over through a file, ncclReduce inside
    libtipe_rt.so (no Python, no torch.distributed on the data path).  Same image as one process on one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "apps"), "main_b200"], stderr=subprocess.DEVNULL)
    exe = os.path.join(ROOT, "apps", "main_b200")
    common = ["--width", "200", "--spp", "30", "--bounces", "5"]
    one = str(tmp_path / "one.ppm")
    subprocess.check_call([exe, *common, "--out", one], stdout=subprocess.DEVNULL)
    for extra, name in (([], "two.ppm"), (["--tile-split"], "two_tiles.ppm")):
        out = str(tmp_path / name)
        idf = str(tmp_path / (name + ".id"))
        procs = [subprocess.Popen([exe, *common, *extra, "--nranks", "2", "--rank", str(r), "--id-file", idf, "--out", out],
                                  stdout=subprocess.DEVNULL) for r in (1, 0)]
        assert [p.wait(timeout=120) for p in procs] == [0, 0]
        assert open(out).read() == open(one).read()


def test_page_locked_output_arrays_are_written_directly(rt):
    """trt_host_alloc'ed canva / albedo_tab / normal_tab: the GPU writes them over PCIe without the staging buffer; same values as into
    ordinary numpy arrays, on a frame large enough for the staged path (>= 8 MB) and on a small one."""
    sc = S.cornell_spheres()
    rt.scene_set(sc)
    for width, spp in ((640, 4), (64, 6)):
        rp = S.params_for(sc, width=width, spp=spp)
        a = rt.render(rp)
        pinned = trt.host_arrays(rp.width * rp.height, rt.lib)
        b = rt.render(rp, out=pinned)
        for k in ("canva", "albedo", "normal"):
            assert b[k] is pinned[k] and a[k].tobytes() == b[k].tobytes()
