"""CPU tests of the host side: the C ABI surface, the C asset loaders (apps/trt_assets.h) against the
reference's own loaders, and the Python mirror's marshalling.  No GPU, no compute calls."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import tipe_raytracer_b200 as trt
from tipe_raytracer_b200 import scene as S

MINI = os.path.join(ROOT, "tests", "golden", "mini_scene")
REFERENCE = "/root/reference"


def test_abi_library_loads_and_exports_every_declared_symbol():
    if not os.path.exists(trt.LIB_PATH):
        trt.build_extension()
    header = open(os.path.join(ROOT, "include", "tipe_rt.h")).read()
    declared = sorted(set(re.findall(r"TRT_API\s+[\w\s\*]+?\b(trt_\w+)\s*\(", header)))
    assert declared == sorted(trt.ABI_SYMBOLS), (declared, sorted(trt.ABI_SYMBOLS))
    lib = trt.load_library()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.trt_version() >= 100
    # struct sizes the header static_asserts, seen from Python
    assert ctypes.sizeof(trt.TrtParams) == 5 * 4 + 4 + 4 * 8 + 96 + 8 + 4 * 4
    ov = np.zeros(3, S.MAT_OVERRIDE)
    assert lib.trt_default_overrides(ov.ctypes.data_as(ctypes.c_void_p), 3) == 3
    assert np.array_equal(ov, S.default_overrides())


def test_no_device_fails_loudly_instead_of_falling_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(trt.TrtError, match="no CUDA device|CPU fallback"):
        trt.Renderer()


def _dump(obj, mtl, tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "apps"), "assets_dump"])
    out = str(tmp_path / "dump.bin")
    subprocess.check_call([os.path.join(ROOT, "apps", "assets_dump"), obj, mtl, out], stdout=subprocess.DEVNULL)
    raw = open(out, "rb").read()
    n, nb, tw, th = np.frombuffer(raw[:16], np.int32)
    tris = np.frombuffer(raw[16:16 + 200 * n], S.TRIANGLE)
    quel = np.frombuffer(raw[16 + 200 * n:16 + 204 * n], np.int32)
    mats = np.frombuffer(raw[16 + 204 * n:], S.MATERIAL)
    return tris, quel, mats, int(tw), int(th), int(nb)


def test_c_loader_on_the_mini_scene(tmp_path):
    tris, quel, mats, tw, th, nb = _dump(os.path.join(MINI, "mini.obj"), os.path.join(MINI, "mini.mtl"), tmp_path)
    assert (tris.shape[0], nb, tw, th) == (16, 5, 4, 4)
    assert quel.tolist() == [0, 0, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4]
    assert np.allclose(tris[0]["A"], [-2, 0, 2]) and np.allclose(tris[0]["uvC"], [3, 3])       # tiled UVs kept
    m = mats.reshape(nb, th, tw)
    assert set(np.unique(m[2]["alpha"])) == {0.0, 1.0}                                         # plant: cut-outs
    assert np.all(m[0]["alpha"] == 1.0) and np.all(mats["emissionStrength"] == 0.0)
    # row 0 of the slab is the BOTTOM row of the file (texture.h:227)
    first_file_row = [int(x) for x in open(os.path.join(MINI, "tex", "ground.ppm")).read().split()[4:7]]
    assert np.allclose(m[0][th - 1][0]["diffuseColor"] * 255.0, first_file_row)


@pytest.mark.parametrize("obj,mtl", [
    ("tests/golden/mini_scene/mini.obj", "tests/golden/mini_scene/mini.mtl"),
    (REFERENCE + "/model3D/RTX_MAP/nature/mineways_doubleface_tri.obj", REFERENCE + "/model3D/RTX_MAP/nature/mineways_doubleface_tri.mtl"),
    (REFERENCE + "/model3D/water/mineways_tri.obj", REFERENCE + "/model3D/water/mineways_tri.mtl"),
])
def test_c_loader_matches_reference_loader_bit_for_bit(ref, tmp_path, obj, mtl):
    obj, mtl = (p if os.path.isabs(p) else os.path.join(ROOT, p) for p in (obj, mtl))
    if not os.path.exists(obj):
        pytest.skip("reference assets not present on this machine")
    tris, quel, mats, tw, th, nb = _dump(obj, mtl, tmp_path)
    rt, rq, rm, rtw, rth, rnb, _ = ref.load_mesh(obj, mtl)          # list_of_mesh + create_mat_list_mtl themselves
    assert (tw, th, nb) == (rtw, rth, rnb) and np.array_equal(quel, rq)
    for f in ("A", "B", "C", "uvA", "uvB", "uvC"):
        assert np.array_equal(tris[f], rt[f]), f
    assert np.array_equal(mats["diffuseColor"], rm["diffuseColor"]) and np.array_equal(mats["alpha"], rm["alpha"])


def test_fixture_scenes_rebuild_reference_arrays():
    nat = S.nature_mesh()
    assert nat.triangles.shape[0] == 5812 and nat.nb_materials == 31 and (nat.tex_w, nat.tex_h) == (16, 16)
    assert nat.mat_list.shape[0] == 31 * 256 and nat.quel_mat_pour_tri.max() == 30
    tree = S.tree_mesh()
    assert tree.triangles.shape[0] == 1792 and (tree.tex_w, tree.tex_h) == (1024, 1024)
    big = S.tree_mesh(grid=3)
    assert big.triangles.shape[0] == 9 * 1792 and big.name.endswith("synthetic")


def test_params_marshalling_and_main_c_truncation():
    sc = S.cornell_spheres()
    sc.aperture = (0.7, 1.9)
    rp = S.params_for(sc, width=100, spp=3)
    p = trt.make_params(rp)
    assert (p.largeur_image, p.hauteur_image, p.nbRayonParPixel, p.nbRebondMax, p.useAO) == (100, 75, 3, 5, 1)
    assert (p.AO_intensity, p.focus_distance, p.ouverture_x, p.ouverture_y) == (2.0, 3.0, 0.0, 1.0)     # main.c:42-43
    rp.main_c_truncation = False
    p = trt.make_params(rp)
    assert (p.AO_intensity, p.ouverture_x, p.ouverture_y) == (2.5, 0.7, 1.9)
    assert list(p.cam) == np.ascontiguousarray(sc.camera()).view(np.float64).reshape(12).tolist()


def test_c_host_program_builds():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "apps"), "main_b200"], stderr=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(ROOT, "apps", "main_b200"))


def test_integration_stub_compiles_inside_the_reference_main_c(tmp_path):
    """INTEGRATION.md §1 is a patch for main.c: splice its replacement block over main.c:402-453 (the pthread fan-out) and let
    gcc check it against the reference's own declarations (-fsyntax-only, stub OIDN header): the casts from the reference's
    `sphere*` / `triangle*` / `material*` / `color*` to the trt_* types and the names the stub uses are proven to exist."""
    import re
    import subprocess
    ref_main = "/root/reference/main.c"
    if not os.path.exists(ref_main):
        pytest.skip("reference tree not mounted (GPU box)")
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```c\n(.*?)```", doc, re.S).group(1)
    body = block[block.index("    /* ---- replaces main.c:402-453"):block.index("    /* ---- main.c:455 onwards unchanged")]
    src = open(ref_main).read().split("\n")
    assert "pthread_t threads[NUM_THREADS];" in src[403] and "pthread_join" in src[451], "main.c:402-453 moved"
    patched = ['#include "tipe_rt.h"'] + src[:401] + body.split("\n") + src[453:]
    out = tmp_path / "main_patched.c"
    out.write_text("\n".join(patched))
    r = subprocess.run(["gcc", "-fsyntax-only", "-w", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "oidn_stub"),
                        "-I", "/root/reference", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_c_host_program_options_and_scene_file(tmp_path):
    """apps/main_b200.c: the constants of main.c:293-347 and the overrides of texture.h:71-87 as data (--config file,
    --override, --no-overrides, --sphere), checked through --dry-run (no GPU needed)."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "apps"), "main_b200"], stderr=subprocess.DEVNULL)
    exe = os.path.join(ROOT, "apps", "main_b200")
    cfg = tmp_path / "scene.cfg"
    cfg.write_text("# a scene file\nscene = none\nsphere = 0 0 0 100000  0 0 0  0.784 0.965 1  1 0 1 1\nwidth = 320\nratio = 2\nspp = 7\n"
                   "no-overrides\noverride = 4:alpha=0.5,index=1.2\norigin = 1,2,3\nao = 0\n")
    out = subprocess.run([exe, "--config", str(cfg), "--override", "1:emission=1/0.5/0.25,strength=2", "--bounces", "3", "--dry-run"],
                         capture_output=True, text=True, check=True).stdout
    assert "image 320x160 spp 7 bounces 3 ao 0" in out and "origin 1,2,3" in out
    assert "spheres 1" in out and "sphere 0 0 0 100000  0 0 0  0.784 0.965 1  1 0 1 1" in out
    assert "overrides 2" in out and "override 4 mask 24" in out and "alpha 0.5 index 1.2" in out
    assert "override 1 mask 3 emission 1/0.5/0.25 strength 2" in out
    # defaults: the README scene and the reference's three built-in overrides
    out = subprocess.run([exe, "--dry-run"], capture_output=True, text=True, check=True).stdout
    assert "image 1000x750 spp 100 bounces 5 ao 1 ao_intensity 2.5" in out and "spheres 10" in out and "overrides 3" in out
    assert "override 1 mask 11 emission 1/1/1 strength 1.85" in out and "override 3 mask 28" in out
    bad = subprocess.run([exe, "--override", "4:colour=1", "--dry-run"], capture_output=True, text=True)
    assert bad.returncode == 2
