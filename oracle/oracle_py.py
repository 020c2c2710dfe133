"""TEST INFRASTRUCTURE — ctypes bindings over the two CPU checkers.

  Oracle : oracle/libtipe_oracle.so, the C restatement (tipe_oracle.c).
  Ref    : oracle/_ref/libtipe_ref.so, the reference's own main.c compiled as a library.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; the product package (tipe_raytracer_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libtipe_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libtipe_ref.so")

RNG_LIBC, RNG_PHILOX = 0, 1


def build(force: bool = False) -> None:
    """make -C oracle: always the port; _ref only where /root/reference exists."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "tipe_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "libtipe_oracle.so"])
    if os.path.exists("/root/reference/main.c") and (force or not os.path.exists(REF_SO) or
            os.path.getmtime(REF_SO) < os.path.getmtime(os.path.join(HERE, "ref_wrapper.c"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class TrtParams(C.Structure):
    _fields_ = [
        ("largeur_image", C.c_int), ("hauteur_image", C.c_int),
        ("nbRayonParPixel", C.c_int), ("nbRebondMax", C.c_int),
        ("useAO", C.c_int),
        ("AO_intensity", C.c_double),
        ("focus_distance", C.c_double), ("ouverture_x", C.c_double), ("ouverture_y", C.c_double),
        ("cam", C.c_double * 12),
        ("seed", C.c_uint64),
        ("sample_offset", C.c_int), ("sample_total", C.c_int),
        ("flags", C.c_int), ("reserved", C.c_int),
    ]


def make_params(rp) -> TrtParams:
    """tipe_raytracer_b200.scene.RenderParams -> trt_params."""
    ao, fd, ax, ay = rp.effective()
    p = TrtParams()
    p.largeur_image, p.hauteur_image = rp.width, rp.height
    p.nbRayonParPixel, p.nbRebondMax = rp.spp, rp.bounces
    p.useAO = 1 if rp.use_ao else 0
    p.AO_intensity, p.focus_distance, p.ouverture_x, p.ouverture_y = ao, fd, ax, ay
    cam = np.ascontiguousarray(rp.cam).view(np.float64).reshape(12)
    for k in range(12):
        p.cam[k] = float(cam[k])
    p.seed = rp.seed
    p.sample_offset, p.sample_total, p.flags = rp.sample_offset, rp.sample_total, rp.flags
    return p


class OracleScene(C.Structure):
    _fields_ = [
        ("spheres", C.c_void_p), ("nbSpheres", C.c_int),
        ("tris", C.c_void_p), ("nbTriangles", C.c_int),
        ("mat_list", C.c_void_p), ("tex_w", C.c_int), ("tex_h", C.c_int),
        ("quelMatPourTri", C.c_void_p),
        ("overrides", C.c_void_p), ("n_overrides", C.c_int),
        ("sky_mat_list", C.c_void_p), ("sky_w", C.c_int), ("sky_h", C.c_int),
    ]


class OracleStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("tri_tests", C.c_uint64), ("seconds", C.c_double)]


def _scene_arrays(scene):
    sph = np.ascontiguousarray(scene.spheres)
    tri = np.ascontiguousarray(scene.triangles)
    mats = np.ascontiguousarray(scene.mat_list)
    quel = np.ascontiguousarray(scene.quel_mat_pour_tri, dtype=np.int32)
    return sph, tri, mats, quel


class Oracle:
    """The C restatement."""

    def __init__(self):
        build()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.oracle_render.restype = C.c_int
        self.lib.oracle_primary.restype = C.c_int
        self.lib.oracle_hit_sphere.restype = C.c_int
        self.lib.oracle_hit_triangle.restype = C.c_int

    def _scene(self, scene):
        keep = _scene_arrays(scene)
        sph, tri, mats, quel = keep
        sc = OracleScene()
        sc.spheres, sc.nbSpheres = _ptr(sph), sph.shape[0]
        sc.tris, sc.nbTriangles = _ptr(tri), tri.shape[0]
        sc.mat_list, sc.tex_w, sc.tex_h = _ptr(mats), scene.tex_w, scene.tex_h
        sc.quelMatPourTri = _ptr(quel)
        if scene.overrides is None:
            sc.overrides, sc.n_overrides = None, -1
        else:
            ov = np.ascontiguousarray(scene.overrides)
            keep = keep + (ov,)
            sc.overrides, sc.n_overrides = _ptr(ov), ov.shape[0]
        if getattr(scene, "sky_mat_list", None) is not None:
            sky = np.ascontiguousarray(scene.sky_mat_list)
            keep = keep + (sky,)
            sc.sky_mat_list, sc.sky_w, sc.sky_h = _ptr(sky), scene.sky_w, scene.sky_h
        else:
            sc.sky_mat_list, sc.sky_w, sc.sky_h = None, 0, 0
        return sc, keep

    def render(self, scene, rp, rng=RNG_PHILOX, threads=0, want_sums=False):
        """-> dict(canva, albedo, normal [H*W,3] float64, sums [H*W,9] or None, stats)."""
        if threads <= 0:
            threads = os.cpu_count() or 1
        sc, keep = self._scene(scene)
        p = make_params(rp)
        n = rp.width * rp.height
        canva, albedo, normal = (np.zeros((n, 3)) for _ in range(3))
        sums = np.zeros((n, 9)) if want_sums else None
        st = OracleStats()
        rc = self.lib.oracle_render(C.byref(sc), C.byref(p), C.c_int(rng), C.c_int(threads),
                                    _ptr(canva), _ptr(albedo), _ptr(normal), _ptr(sums), C.byref(st))
        if rc != 0:
            raise RuntimeError(f"oracle_render failed: {rc}")
        del keep
        return dict(canva=canva, albedo=albedo, normal=normal, sums=sums,
                    stats=dict(paths=st.paths, rays=st.rays, sphere_tests=st.sphere_tests,
                               tri_tests=st.tri_tests, seconds=st.seconds))

    def primary(self, scene, rp):
        sc, keep = self._scene(scene)
        p = make_params(rp)
        n = rp.width * rp.height
        t = np.zeros(n); ids = np.zeros(n, np.int32); normal = np.zeros((n, 3)); albedo = np.zeros((n, 3))
        rc = self.lib.oracle_primary(C.byref(sc), C.byref(p), _ptr(t), _ptr(ids), _ptr(normal), _ptr(albedo))
        if rc != 0:
            raise RuntimeError("oracle_primary failed")
        del keep
        return dict(t=t, id=ids, normal=normal, albedo=albedo)

    # --- leaf functions --------------------------------------------------------------------
    def init_camera(self, origin, target, up, vfov, ratio):
        out = np.zeros(12)
        a = [np.asarray(x, np.float64) for x in (origin, target, up)]
        self.lib.oracle_init_camera(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), C.c_double(vfov), C.c_double(ratio), _ptr(out))
        return out

    def get_ray(self, u, v, cam, focus, dx, dy):
        cam = np.ascontiguousarray(cam).view(np.float64).reshape(12).copy()
        o, d = np.zeros(3), np.zeros(3)
        self.lib.oracle_get_ray(C.c_double(u), C.c_double(v), _ptr(cam), C.c_double(focus),
                                C.c_double(dx), C.c_double(dy), _ptr(o), _ptr(d))
        return o, d

    def hit_sphere(self, center, radius, o, d):
        center, o, d = (np.asarray(x, np.float64).copy() for x in (center, o, d))
        dst = C.c_double(0); pt, nr = np.zeros(3), np.zeros(3)
        hit = self.lib.oracle_hit_sphere(_ptr(center), C.c_double(radius), _ptr(o), _ptr(d), C.byref(dst), _ptr(pt), _ptr(nr))
        return bool(hit), dst.value, pt, nr

    def hit_triangle(self, tri, o, d):
        tri = np.ascontiguousarray(tri); o, d = (np.asarray(x, np.float64).copy() for x in (o, d))
        dst = C.c_double(0); pt, nr = np.zeros(3), np.zeros(3)
        hit = self.lib.oracle_hit_triangle(_ptr(tri), _ptr(o), _ptr(d), C.byref(dst), _ptr(pt), _ptr(nr))
        return bool(hit), dst.value, pt, nr

    def tri_uvmapping(self, scene, tri_index, point, normal):
        from tipe_raytracer_b200.scene import MATERIAL
        sc, keep = self._scene(scene)
        point, normal = (np.asarray(x, np.float64).copy() for x in (point, normal))
        out = np.zeros(1, MATERIAL)
        self.lib.oracle_tri_uvmapping(C.byref(sc), C.c_int(tri_index), _ptr(point), _ptr(normal), _ptr(out))
        del keep
        return out[0]

    def sphere_uvmapping(self, center, radius, point, sky_w, sky_h) -> int:
        center, point = (np.asarray(x, np.float64).copy() for x in (center, point))
        self.lib.oracle_sphere_uvmapping.restype = C.c_int
        return int(self.lib.oracle_sphere_uvmapping(_ptr(center), C.c_double(radius), _ptr(point), C.c_int(sky_w), C.c_int(sky_h)))

    def reflect(self, v, n):
        v, n = (np.asarray(x, np.float64).copy() for x in (v, n)); out = np.zeros(3)
        self.lib.oracle_reflect(_ptr(v), _ptr(n), _ptr(out)); return out

    def refract(self, v, n, n1, n2):
        v, n = (np.asarray(x, np.float64).copy() for x in (v, n)); out = np.zeros(3)
        self.lib.oracle_refract(_ptr(v), _ptr(n), C.c_double(n1), C.c_double(n2), _ptr(out)); return out

    def tonemap(self, s, spp):
        s = np.asarray(s, np.float64).copy(); out = np.zeros(3)
        self.lib.oracle_tonemap(_ptr(s), C.c_int(spp), _ptr(out)); return out

    def hsl_roundtrip(self, rgb):
        rgb = np.asarray(rgb, np.float64).copy(); out = np.zeros(3)
        self.lib.oracle_hsl_roundtrip(_ptr(rgb), _ptr(out)); return out

    def pile(self, kinds, values):
        kinds = np.asarray(kinds, np.int32).copy(); values = np.asarray(values, np.float64).copy()
        out = np.zeros(2 * len(kinds))
        self.lib.oracle_pile(_ptr(kinds), _ptr(values), C.c_int(len(kinds)), _ptr(out))
        return out.reshape(-1, 2)

    def tracer_libc(self, scene, o, d, bounces, ao, use_ao):
        sc, keep = self._scene(scene)
        o, d = (np.asarray(x, np.float64).copy() for x in (o, d)); out = np.zeros(9)
        self.lib.oracle_tracer_libc(C.byref(sc), _ptr(o), _ptr(d), C.c_int(bounces), C.c_double(ao), C.c_int(1 if use_ao else 0), _ptr(out))
        del keep
        return out

    def philox(self, ctr, key):
        ctr = np.asarray(ctr, np.uint32).copy(); key = np.asarray(key, np.uint32).copy(); out = np.zeros(4, np.uint32)
        self.lib.oracle_philox4x32_10(_ptr(ctr), _ptr(key), _ptr(out)); return out


class Ref:
    """The reference's own code (oracle/_ref/libtipe_ref.so)."""

    @staticmethod
    def available() -> bool:
        try:
            build()
        except Exception:
            pass
        return os.path.exists(REF_SO)

    def __init__(self):
        build()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = C.CDLL(REF_SO)
        self.lib.ref_render.restype = C.c_double
        self.lib.ref_ray_count.restype = C.c_uint64
        self.lib.ref_kat_random_double.restype = C.c_double
        self.lib.ref_kat_hit_sphere.restype = C.c_int
        self.lib.ref_kat_hit_triangle.restype = C.c_int
        self.lib.ref_load_mesh.restype = C.c_void_p
        self.lib.ref_load_textures.restype = C.c_void_p
        self.lib.ref_mat_path.restype = C.c_char_p

    def sizes(self):
        out = (C.c_int * 12)()
        self.lib.ref_sizes(out)
        return list(out)

    def srand(self, seed):
        self.lib.ref_srand(C.c_uint(seed))

    def init_camera(self, origin, target, up, vfov, ratio):
        out = np.zeros(12)
        a = [np.asarray(x, np.float64) for x in (origin, target, up)]
        self.lib.ref_init_camera(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), C.c_double(vfov), C.c_double(ratio), _ptr(out))
        return out

    def render(self, scene, rp, threads=1, seed=None, count_rays=False, quiet=True):
        """fill_canva fan-out of main.c:402-453. seed=None keeps the current rand() state."""
        sph, tri, mats, quel = _scene_arrays(scene)
        if scene.overrides is not None:
            raise ValueError("the reference hard-codes its overrides (texture.h:71-87)")
        ao, fd, ax, ay = rp.effective()
        cam = np.ascontiguousarray(rp.cam).view(np.float64).reshape(12).copy()
        n = rp.width * rp.height
        canva, albedo, normal = (np.zeros((n, 3)) for _ in range(3))
        if seed is not None:
            self.srand(seed)
        self.lib.ref_ray_count_reset()
        self.lib.ref_count_rays(C.c_int(1 if count_rays else 0))
        secs = self.lib.ref_render(_ptr(sph), C.c_int(sph.shape[0]), _ptr(tri), C.c_int(tri.shape[0]),
                                   _ptr(mats), C.c_int(scene.tex_w), C.c_int(scene.tex_h), _ptr(quel), _ptr(cam),
                                   C.c_int(rp.width), C.c_int(rp.height), C.c_int(rp.spp), C.c_int(rp.bounces),
                                   C.c_int(1 if rp.use_ao else 0), C.c_int(int(ao)), C.c_int(int(fd)),
                                   C.c_int(int(ax)), C.c_int(int(ay)), C.c_int(threads), C.c_int(1 if quiet else 0),
                                   _ptr(canva), _ptr(albedo), _ptr(normal))
        rays = int(self.lib.ref_ray_count())
        self.lib.ref_count_rays(C.c_int(0))
        return dict(canva=canva, albedo=albedo, normal=normal,
                    stats=dict(paths=n * rp.spp, rays=rays, seconds=secs))

    def primary(self, scene, rp):
        sph, tri, mats, quel = _scene_arrays(scene)
        _, fd, _, _ = rp.effective()
        cam = np.ascontiguousarray(rp.cam).view(np.float64).reshape(12).copy()
        n = rp.width * rp.height
        t = np.zeros(n); ids = np.zeros(n, np.int32); normal = np.zeros((n, 3)); albedo = np.zeros((n, 3))
        self.lib.ref_primary(_ptr(sph), C.c_int(sph.shape[0]), _ptr(tri), C.c_int(tri.shape[0]),
                             _ptr(mats), C.c_int(scene.tex_w), C.c_int(scene.tex_h), _ptr(quel), _ptr(cam),
                             C.c_int(rp.width), C.c_int(rp.height), C.c_double(fd),
                             _ptr(t), _ptr(ids), _ptr(normal), _ptr(albedo))
        return dict(t=t, id=ids, normal=normal, albedo=albedo)

    # --- leaf functions --------------------------------------------------------------------
    def get_ray(self, u, v, cam, focus, dx, dy):
        cam = np.ascontiguousarray(cam).view(np.float64).reshape(12).copy()
        out = np.zeros(6)
        self.lib.ref_kat_get_ray(C.c_double(u), C.c_double(v), _ptr(cam), C.c_double(focus), C.c_double(dx), C.c_double(dy), _ptr(out))
        return out[:3], out[3:]

    def hit_sphere(self, center, radius, o, d):
        center = np.asarray(center, np.float64).copy(); r = np.concatenate([o, d]).astype(np.float64)
        dst = C.c_double(0); pt, nr = np.zeros(3), np.zeros(3)
        hit = self.lib.ref_kat_hit_sphere(_ptr(center), C.c_double(radius), _ptr(r), C.byref(dst), _ptr(pt), _ptr(nr))
        return bool(hit), dst.value, pt, nr

    def hit_triangle(self, tri, o, d):
        tri = np.ascontiguousarray(tri); r = np.concatenate([o, d]).astype(np.float64)
        dst = C.c_double(0); pt, nr = np.zeros(3), np.zeros(3)
        hit = self.lib.ref_kat_hit_triangle(_ptr(tri), _ptr(r), C.byref(dst), _ptr(pt), _ptr(nr))
        return bool(hit), dst.value, pt, nr

    def tri_uvmapping(self, scene, tri_index, point, normal):
        from tipe_raytracer_b200.scene import MATERIAL
        sph, tri, mats, quel = _scene_arrays(scene)
        point, normal = (np.asarray(x, np.float64).copy() for x in (point, normal))
        out = np.zeros(1, MATERIAL)
        one = np.ascontiguousarray(tri[tri_index:tri_index + 1])
        self.lib.ref_kat_tri_uvmapping(_ptr(one), _ptr(point), _ptr(normal), _ptr(mats), C.c_int(scene.tex_w),
                                       C.c_int(scene.tex_h), C.c_int(tri_index), _ptr(quel), _ptr(out))
        return out[0]

    def sphere_uvmapping(self, center, radius, point, sky_w, sky_h) -> int:
        """texel index sphere_uvmapping (texture.h:92-112) reads, recovered through a sky whose texel k has diffuse.r = k."""
        from tipe_raytracer_b200.scene import MATERIAL, SPHERE
        s = np.zeros(1, SPHERE); s[0]["center"] = center; s[0]["radius"] = radius
        sky = np.zeros(sky_w * sky_h, MATERIAL); sky["diffuseColor"][:, 0] = np.arange(sky_w * sky_h)
        point = np.asarray(point, np.float64).copy(); out = np.zeros(1, MATERIAL)
        self.lib.ref_kat_sphere_uvmapping(_ptr(s), _ptr(point), _ptr(sky), C.c_int(sky_w), C.c_int(sky_h), _ptr(out))
        return int(out[0]["diffuseColor"][0])

    def reflect(self, v, n):
        v, n = (np.asarray(x, np.float64).copy() for x in (v, n)); out = np.zeros(3)
        self.lib.ref_kat_reflect(_ptr(v), _ptr(n), _ptr(out)); return out

    def refract(self, v, n, n1, n2):
        v, n = (np.asarray(x, np.float64).copy() for x in (v, n)); out = np.zeros(3)
        self.lib.ref_kat_refract(_ptr(v), _ptr(n), C.c_double(n1), C.c_double(n2), _ptr(out)); return out

    def tonemap(self, s, spp):
        s = np.asarray(s, np.float64).copy(); out = np.zeros(3)
        self.lib.ref_kat_tonemap(_ptr(s), C.c_int(spp), _ptr(out)); return out

    def hsl_roundtrip(self, rgb):
        rgb = np.asarray(rgb, np.float64).copy(); out = np.zeros(3)
        self.lib.ref_kat_hsl_roundtrip(_ptr(rgb), _ptr(out)); return out

    def pile(self, kinds, values):
        kinds = np.asarray(kinds, np.int32).copy(); values = np.asarray(values, np.float64).copy()
        out = np.zeros(2 * len(kinds))
        self.lib.ref_kat_pile(_ptr(kinds), _ptr(values), C.c_int(len(kinds)), _ptr(out))
        return out.reshape(-1, 2)

    def tracer(self, scene, o, d, bounces, ao, use_ao):
        sph, tri, mats, quel = _scene_arrays(scene)
        r = np.concatenate([o, d]).astype(np.float64); out = np.zeros(9)
        self.lib.ref_kat_tracer(_ptr(r), C.c_int(bounces), _ptr(sph), C.c_int(sph.shape[0]), _ptr(tri), C.c_int(tri.shape[0]),
                                C.c_double(ao), C.c_int(1 if use_ao else 0), _ptr(mats), C.c_int(scene.tex_w),
                                C.c_int(scene.tex_h), _ptr(quel), _ptr(out))
        return out

    def load_mesh(self, obj_path: str, mtl_path: str):
        """The reference loaders: list_of_mesh (mesh.h:96-218) + create_mat_list_mtl (texture.h:175-354).
        -> (TRIANGLE[n], quel_mat int32[n], MATERIAL[nbMat*th*tw] or None, tw, th, nbMat, [texture paths])."""
        from tipe_raytracer_b200.scene import TRIANGLE, MATERIAL
        nt, nm = C.c_int(0), C.c_int(0)
        paths, quel = C.c_void_p(), C.c_void_p()
        tptr = self.lib.ref_load_mesh(obj_path.encode(), mtl_path.encode(), C.byref(nt), C.byref(nm), C.byref(paths), C.byref(quel))
        n, m = nt.value, nm.value
        tris = np.frombuffer((C.c_char * (n * TRIANGLE.itemsize)).from_address(tptr), dtype=TRIANGLE).copy()
        quel_mat = np.frombuffer((C.c_char * (n * 4)).from_address(quel.value), dtype=np.int32).copy()
        names = [self.lib.ref_mat_path(paths, C.c_int(k)) for k in range(m)]
        names = [None if s is None else s.decode() for s in names]
        tw, th = C.c_int(0), C.c_int(0)
        mats = None
        if all(s is not None for s in names):
            mptr = self.lib.ref_load_textures(paths, C.c_int(m), C.byref(tw), C.byref(th))
            if mptr:
                cnt = m * tw.value * th.value
                mats = np.frombuffer((C.c_char * (cnt * MATERIAL.itemsize)).from_address(mptr), dtype=MATERIAL).copy()
                self.lib.ref_free(C.c_void_p(mptr))
        self.lib.ref_free(C.c_void_p(tptr)); self.lib.ref_free(quel)
        return tris, quel_mat, mats, tw.value, th.value, m, names
