"""Host-side scene description: the reference's C structs as numpy dtypes, the camera set-up and
the five BASELINE configurations.

Mirrors (layout and meaning, not code) of the reference's C surface:
  vec3/material/sphere/UV/triangle/camera -> vec3.h:7-9, hitinfo.h:6-13, sphere.h:7-11,
  mesh.h:9-22, camera.h:10-15; init_camera -> camera.h:17-40; render knobs -> main.c:293-317.
Arrays built here are handed to libtipe_rt.so by pointer, exactly like main.c's own arrays.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Optional

import numpy as np

VEC3 = np.dtype([("e", "<f8", (3,))])
MATERIAL = np.dtype([
    ("diffuseColor", "<f8", (3,)),
    ("emissionColor", "<f8", (3,)),
    ("emissionStrength", "<f8"),
    ("reflectionStrength", "<f8"),
    ("alpha", "<f8"),
    ("materialIndex", "<f8"),
])
SPHERE = np.dtype([("center", "<f8", (3,)), ("radius", "<f8"), ("mat", MATERIAL)])
TRIANGLE = np.dtype([
    ("A", "<f8", (3,)), ("B", "<f8", (3,)), ("C", "<f8", (3,)),
    ("mat", MATERIAL),
    ("uvA", "<f8", (2,)), ("uvB", "<f8", (2,)), ("uvC", "<f8", (2,)),
])
CAMERA = np.dtype([
    ("origin", "<f8", (3,)), ("horizontal", "<f8", (3,)),
    ("vertical", "<f8", (3,)), ("coin_bas_gauche", "<f8", (3,)),
])
MAT_OVERRIDE = np.dtype([
    ("material_ordinal", "<i4"), ("mask", "<u4"),
    ("emissionColor", "<f8", (3,)), ("emissionStrength", "<f8"),
    ("reflectionStrength", "<f8"), ("alpha", "<f8"), ("materialIndex", "<f8"),
])

assert MATERIAL.itemsize == 80 and SPHERE.itemsize == 112 and TRIANGLE.itemsize == 200
assert CAMERA.itemsize == 96 and MAT_OVERRIDE.itemsize == 64

OV_EMISSION_COLOR, OV_EMISSION_STRENGTH, OV_REFLECTION, OV_ALPHA, OV_MATERIAL_INDEX = 1, 2, 4, 8, 16

# rtutility.h:16-21
RED, GREEN, BLUE = (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)
WHITE, BLACK, SKY = (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), (0.784, 0.965, 1.0)


def material(diffuse, emission=BLACK, emission_strength=0.0, reflection=0.0, alpha=1.0, index=1.0):
    m = np.zeros((), MATERIAL)
    m["diffuseColor"] = diffuse
    m["emissionColor"] = emission
    m["emissionStrength"] = emission_strength
    m["reflectionStrength"] = reflection
    m["alpha"] = alpha
    m["materialIndex"] = index
    return m


def spheres(rows) -> np.ndarray:
    """rows: iterable of (center, radius, material()) -> SPHERE array."""
    rows = list(rows)
    out = np.zeros(len(rows), SPHERE)
    for k, (c, r, m) in enumerate(rows):
        out[k]["center"] = c
        out[k]["radius"] = r
        out[k]["mat"] = m
    return out


def default_overrides() -> np.ndarray:
    """The three per-ordinal overrides hard-coded at texture.h:71-87."""
    ov = np.zeros(3, MAT_OVERRIDE)
    ov[0] = (1, OV_EMISSION_COLOR | OV_EMISSION_STRENGTH | OV_ALPHA, WHITE, 1.85, 0.0, 1.0, 0.0)
    ov[1] = (4, OV_ALPHA | OV_MATERIAL_INDEX | OV_REFLECTION, BLACK, 0.0, 0.93, 0.6, 1.33)
    ov[2] = (3, OV_ALPHA | OV_MATERIAL_INDEX | OV_REFLECTION, BLACK, 0.0, 0.3, 0.1, 1.50)
    return ov


def init_camera(origin, target, up, vfov: float, ratio: float) -> np.ndarray:
    """Look-at camera, camera.h:17-40 (same operation order, FP64)."""
    def norm(v):
        return v / math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])

    def cross(a, b):
        return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])

    origin = np.asarray(origin, np.float64)
    target = np.asarray(target, np.float64)
    up = np.asarray(up, np.float64)
    theta = vfov * 3.1415926535897932385 / 180.0
    h = math.tan(theta / 2)
    viewport_h = 2.0 * h
    viewport_w = ratio * viewport_h
    w = norm(origin - target)
    u = norm(cross(up, w))
    v = cross(w, u)
    cam = np.zeros((), CAMERA)
    cam["origin"] = origin
    cam["horizontal"] = u * viewport_w
    cam["vertical"] = v * viewport_h
    cam["coin_bas_gauche"] = origin - (cam["horizontal"] / 2 + (cam["vertical"] / 2 + w))
    return cam


@dataclasses.dataclass
class Scene:
    """Everything trt_scene_set takes, plus a default camera/knob set for the config."""
    name: str
    spheres: np.ndarray                      # SPHERE[nbSpheres]
    triangles: np.ndarray                    # TRIANGLE[nbTriangles]
    mat_list: np.ndarray                     # MATERIAL[nbMaterials*tex_h*tex_w]
    tex_w: int = 0
    tex_h: int = 0
    nb_materials: int = 0
    quel_mat_pour_tri: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, np.int32))
    overrides: Optional[np.ndarray] = None   # None => reference built-ins (texture.h:71-87)
    # sky environment (ThreadData.sky_mat_list / sky_width / sky_height, main.c:29,34): MATERIAL[sky_h*sky_w] as
    # create_mat_list loads it (texture.h:145-173), row 0 = bottom PPM row.  None => lookup off, like main.c:64-71 as shipped
    sky_mat_list: Optional[np.ndarray] = None
    sky_w: int = 0
    sky_h: int = 0
    # default view / knobs of the config (main.c:293-317)
    cam_origin: tuple = (0.0, 0.0, 0.0)
    cam_target: tuple = (0.0, 0.0, -1.0)
    cam_up: tuple = (0.0, 1.0, 0.0)
    vfov: float = 70.0
    ratio: float = 4.0 / 3.0
    width: int = 1000
    spp: int = 100
    bounces: int = 5
    use_ao: bool = False
    ao_intensity: float = 2.5
    focus_distance: float = 3.0
    aperture: tuple = (0.0, 0.0)

    @property
    def height(self) -> int:
        return int(self.width / self.ratio)          # main.c:295

    def camera(self) -> np.ndarray:
        return init_camera(self.cam_origin, self.cam_target, self.cam_up, self.vfov, self.ratio)


@dataclasses.dataclass
class RenderParams:
    """trt_params.  main_c_truncation=True reproduces ThreadData's int fields (main.c:42-43)."""
    width: int
    height: int
    spp: int
    bounces: int
    cam: np.ndarray
    use_ao: bool = False
    ao_intensity: float = 2.5
    focus_distance: float = 3.0
    aperture: tuple = (0.0, 0.0)
    seed: int = 1010
    sample_offset: int = 0
    sample_total: int = 0
    flags: int = 0
    tile_part: int = 0                 # 0 = whole image, else part | n_parts << 16 (interleaved 8x4 tiles, trt_params.tile_part)
    main_c_truncation: bool = True

    def effective(self):
        ao, fd, ax, ay = self.ao_intensity, self.focus_distance, self.aperture[0], self.aperture[1]
        if self.main_c_truncation:
            ao, fd, ax, ay = float(int(ao)), float(int(fd)), float(int(ax)), float(int(ay))
        return ao, fd, ax, ay


def params_for(scene: Scene, width: Optional[int] = None, spp: Optional[int] = None,
               bounces: Optional[int] = None, **kw) -> RenderParams:
    w = scene.width if width is None else width
    h = int(w / scene.ratio)
    return RenderParams(width=w, height=h, spp=scene.spp if spp is None else spp,
                        bounces=scene.bounces if bounces is None else bounces, cam=scene.camera(),
                        use_ao=scene.use_ao, ao_intensity=scene.ao_intensity,
                        focus_distance=scene.focus_distance, aperture=scene.aperture, **kw)


# ---------------------------------------------------------------------------------------------
# Configurations (BASELINE.json configs, made concrete in SURVEY.md §8d)
# ---------------------------------------------------------------------------------------------
def _readme_spheres():
    """README.md:48-57 with alpha=1, materialIndex=1 appended (the 6-field form of main.c:335-346)."""
    return [
        ((-501, 0, 0), 500, material(GREEN, BLACK, 0.0, 0.96)),
        ((0, -501, 0), 500, material(WHITE, BLACK, 0.0, 0.0)),
        ((501, 0, 0), 500, material(RED, BLACK, 0.0, 0.96)),
        ((-0.5, 1.4, -1.2), 0.5, material(BLACK, (1.0, 0.6, 0.2), 4.0, 0.0)),
        ((0.5, 1.4, -2.2), 0.5, material(BLACK, (0.7, 0.2, 1.0), 4.0, 0.0)),
        ((0.6, -1.4, -1.0), 0.5, material(BLACK, (0.55, 0.863, 1.0), 2.5, 0.0)),
        ((-0.5, -1.4, -3.1), 0.5, material(BLACK, (0.431, 1.0, 0.596), 2.5, 0.0)),
        ((0, 0, -504), 500, material(WHITE, BLACK, 0.0, 0.0)),
        ((0, 501, 0), 500, material(WHITE, BLACK, 0.0, 0.0)),
        ((0.4, -0.5, -3.3), 0.5, material(SKY, BLACK, 0.0, 0.99)),
    ]


_NO_TRIS = dict(triangles=np.zeros(0, TRIANGLE), mat_list=np.zeros(0, MATERIAL))
_README_CAM = dict(cam_origin=(0.34, 0.3, 0.5), cam_target=(0.0, -0.5, -3.0), vfov=70.0, ratio=4.0 / 3.0)


def cornell_spheres() -> Scene:
    """C1: README Cornell-style 10-sphere scene, 1000x750, 100 spp, 5 bounces, AO on."""
    return Scene(name="cornell_spheres", spheres=spheres(_readme_spheres()), **_NO_TRIS, **_README_CAM,
                 width=1000, spp=100, bounces=5, use_ao=True, ao_intensity=2.5)


def readme_benchmark() -> Scene:
    """C5: README benchmark config, 1200x900, 1000 spp, 6 bounces, AO on (README.md:99-110)."""
    s = cornell_spheres()
    return dataclasses.replace(s, name="readme_benchmark", width=1200, spp=1000, bounces=6)


def nested_dielectrics() -> Scene:
    """C3: the C1 box with the mirror ball replaced by three concentric dielectric spheres
    (glass 1.5 / water 1.33 / air, constants of texture.h:77-87); exercises the stack pop branch."""
    rows = _readme_spheres()[:-1]
    c = (0.0, -0.35, -2.4)
    rows += [
        (c, 0.65, material(WHITE, BLACK, 0.0, 0.3, 0.1, 1.5)),
        (c, 0.40, material(SKY, BLACK, 0.0, 0.93, 0.6, 1.33)),
        (c, 0.15, material(WHITE, BLACK, 0.0, 0.3, 0.1, 1.0)),
    ]
    return Scene(name="nested_dielectrics", spheres=spheres(rows), **_NO_TRIS, **_README_CAM,
                 width=1000, spp=1000, bounces=19, use_ao=False)


def sky_sphere():
    """main.c:346 — emissive sky, always the last sphere."""
    return ((0.0, 0.0, 0.0), 100000, material(BLACK, SKY, 1.0, 0.0, 1.0, 1.0))


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def mesh_from_fixture(path: str):
    """Loads a mesh fixture written by tools/make_fixtures.py (arrays captured from the
    reference's own loaders): triangles, per-triangle material ordinal, textures as u8 RGBA.
    mat_list is rebuilt exactly as create_mat_list_mtl does (texture.h:231-246):
    diffuse = rgb/maxVal, alpha = a/maxVal, emissionStrength = 0, every other field 0."""
    z = np.load(path)
    n = z["A"].shape[0]
    tris = np.zeros(n, TRIANGLE)
    tris["A"], tris["B"], tris["C"] = z["A"], z["B"], z["C"]
    tris["uvA"], tris["uvB"], tris["uvC"] = z["uvA"], z["uvB"], z["uvC"]
    tris["mat"]["diffuseColor"] = SKY            # mesh.h:205, never read by the renderer
    tex = z["tex_rgba"]                          # [nbMat, tex_h, tex_w, 4] u8, row 0 = bottom row
    nb_mat, th, tw, _ = tex.shape
    maxval = float(z["maxval"])
    mats = np.zeros(nb_mat * th * tw, MATERIAL)
    flat = tex.reshape(-1, 4).astype(np.float64)
    mats["diffuseColor"] = flat[:, :3] / maxval
    mats["alpha"] = flat[:, 3] / maxval
    return tris, mats, tw, th, nb_mat, z["quel_mat"].astype(np.int32)


def nature_mesh(fixture: Optional[str] = None) -> Scene:
    """C2 (substitute for the absent books mesh): RTX_MAP/nature/mineways_doubleface_tri.obj,
    5 812 triangles, 31 16x16 textures with alpha maps + emissive sky sphere; 1000x750, 200 spp,
    4 bounces, AO off; the built-in overrides turn ordinals 1/3/4 into light/glass/water."""
    tris, mats, tw, th, nb, quel = mesh_from_fixture(fixture or os.path.join(GOLDEN_DIR, "scene_nature.npz"))
    return Scene(name="nature_mesh", spheres=spheres([sky_sphere()]), triangles=tris, mat_list=mats,
                 tex_w=tw, tex_h=th, nb_materials=nb, quel_mat_pour_tri=quel,
                 cam_origin=(3.0, 2.5, 4.0), cam_target=(-0.3, 0.9, 0.0), vfov=40.0, ratio=4.0 / 3.0,
                 width=1000, spp=200, bounces=4, use_ao=False)


def water_mesh(fixture: Optional[str] = None) -> Scene:
    """C2b, the secondary substitute SURVEY §8d names for the absent books mesh: model3D/water/mineways_tri.obj (1 740 triangles,
    17 16x16 textures; NOT double-faced: no twins, unlike the nature mesh) + emissive sky; C2's knobs (200 spp, 4 bounces, AO off).
    The built-in overrides make ordinal 1 a light, 3 glass and 4 water here too."""
    tris, mats, tw, th, nb, quel = mesh_from_fixture(fixture or os.path.join(GOLDEN_DIR, "scene_water.npz"))
    return Scene(name="water_mesh", spheres=spheres([sky_sphere()]), triangles=tris, mat_list=mats,
                 tex_w=tw, tex_h=th, nb_materials=nb, quel_mat_pour_tri=quel,
                 cam_origin=(2.8, 2.3, 3.6), cam_target=(0.0, 0.6, 0.0), vfov=40.0, ratio=4.0 / 3.0,
                 width=1000, spp=200, bounces=4, use_ao=False)


def tree_mesh(fixture: Optional[str] = None, grid: int = 1) -> Scene:
    """C4: 1st_mc_world/tree/tree_tri.obj (1 792 triangles, one 1024^2 RGBA texture) + emissive sky;
    1920x1080, 500 spp, 6 bounces.  grid > 1 instances the tree on a grid x grid lattice by
    translation (what move_mesh does, mesh.h:220-234) — a SYNTHETIC BVH stress variant."""
    tris, mats, tw, th, nb, quel = mesh_from_fixture(fixture or os.path.join(GOLDEN_DIR, "scene_tree.npz"))
    if grid > 1:
        n = tris.shape[0]
        big = np.tile(tris, grid * grid)
        pitch = 1.6
        k = 0
        for gx in range(grid):
            for gz in range(grid):
                off = np.array([(gx - (grid - 1) / 2) * pitch, 0.0, (gz - (grid - 1) / 2) * pitch])
                for f in ("A", "B", "C"):
                    big[f][k * n:(k + 1) * n] += off
                k += 1
        tris, quel = big, np.tile(quel, grid * grid)
    name = "tree_mesh" if grid == 1 else f"tree_grid{grid}x{grid}_synthetic"
    return Scene(name=name, spheres=spheres([sky_sphere()]), triangles=tris, mat_list=mats,
                 tex_w=tw, tex_h=th, nb_materials=nb, quel_mat_pour_tri=quel,
                 cam_origin=(2.2 * max(1, grid), 1.4 * max(1, grid), 3.4 * max(1, grid)),
                 cam_target=(0.05, -0.1, 0.0), vfov=40.0, ratio=16.0 / 9.0,
                 width=1920, spp=500, bounces=6, use_ao=False)


def pyramide_eau(fixture: Optional[str] = None) -> Scene:
    """C3b: the scene main.c ships configured for (main.c:293-322, "refraction stack" renders in results/):
    model3D/pyramide_eau/scene.obj, 34 triangles, no spheres, materials = ordinals 0 walls, 1 light, 2 floor/ceiling,
    3 glass, 4 water (the built-in overrides of texture.h:71-87), coordinates up to +-1800; 16:10, 1000 px, vfov 30.2,
    100 spp, nbRebondMax 20.  Textures: the stripped 1500x1227 PPMs regenerated from the PNGs at 1/4 resolution."""
    tris, mats, tw, th, nb, quel = mesh_from_fixture(fixture or os.path.join(GOLDEN_DIR, "scene_pyramide_eau.npz"))
    return Scene(name="pyramide_eau_quarter_res_textures", spheres=np.zeros(0, SPHERE), triangles=tris, mat_list=mats,
                 tex_w=tw, tex_h=th, nb_materials=nb, quel_mat_pour_tri=quel,
                 cam_origin=(237.6461, 144.496, -962.8788), cam_target=(-176.9382, 141.5864, 651.8203), vfov=30.2,
                 ratio=16.0 / 10.0, width=1000, spp=100, bounces=20, use_ao=False)


def procedural_sky(w: int = 256, h: int = 128) -> np.ndarray:
    """A synthetic lat-long sky in create_mat_list's layout (the reference's model3D/hdr/*.ppm are stripped from the mount):
    horizon-to-zenith gradient, a warm sun disc and a dark ground half, quantised to 8 bits like a P3 PPM."""
    v = (np.arange(h) + 0.5) / h                 # texture row y <-> theta = v*pi measured from -Y (texture.h:97,101): row 0 = straight down
    u = (np.arange(w) + 0.5) / w
    uu, vv = np.meshgrid(u, v)
    up = np.clip((vv - 0.5) * 2.0, 0.0, 1.0)[..., None]
    horizon = np.array([0.95, 0.85, 0.70]); zenith = np.array([0.25, 0.45, 0.95]); ground = np.array([0.18, 0.15, 0.12])
    img = np.where(vv[..., None] >= 0.5, horizon * (1 - up) + zenith * up, ground)
    sun = np.exp(-(((uu - 0.3) * 2.0) ** 2 + ((vv - 0.72) * 1.0) ** 2) / 0.0008)[..., None]
    img = np.clip(img + sun * np.array([1.0, 0.9, 0.6]), 0.0, 1.0)
    img = np.floor(img * 255.0) / 255.0
    mats = np.zeros(w * h, MATERIAL)
    mats["diffuseColor"] = img.reshape(-1, 3)
    return mats


def tree_mesh_sky(fixture: Optional[str] = None) -> Scene:
    """C4 with the sky lookup switched on ("BVH + sky" taken literally): the tree under a lat-long sky image."""
    sc = tree_mesh(fixture)
    sc.name = "tree_mesh_sky_texture"
    sc.sky_w, sc.sky_h = 256, 128
    sc.sky_mat_list = procedural_sky(sc.sky_w, sc.sky_h)
    return sc


CONFIGS = {
    "C1": cornell_spheres,
    "C2": nature_mesh,
    "C3": nested_dielectrics,
    "C4": tree_mesh,
    "C5": readme_benchmark,
    "C3b": pyramide_eau,
    "C2b": water_mesh,
    "C4sky": tree_mesh_sky,
    # BASELINE.md §4 "large triangle mesh": the tree instanced 24 x 24 times = 1 032 192 triangles (SYNTHETIC; the shipped
    # 1st_mc_world holds one 1 792-triangle tree), C4's resolution / bounces, camera pulled back to see the whole forest
    "C4big": lambda: tree_mesh(grid=24),
}
