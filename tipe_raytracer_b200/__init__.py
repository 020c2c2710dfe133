"""tipe_raytracer_b200 — host-side Python mirror of the C ABI (include/tipe_rt.h) of libtipe_rt.so,
the sm_100a replacement for tipe-raytracer's per-pixel path-tracing hot path (main.c:245-284 and
everything it calls).

The product is the CUDA library; this module only marshals numpy arrays laid out like the
reference's C structs into it (ctypes, plain pointers).  There is NO CPU rendering path here:
if libtipe_rt.so is missing or no B200 is visible, construction of `Renderer` raises.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
import subprocess
from typing import Optional

import numpy as np

from . import scene as scene  # noqa: F401  (re-export)
from .scene import RenderParams, Scene

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtipe_rt.so")

FLAG_COUNT_RAYS, FLAG_BRUTE_FORCE, FLAG_COUNT_TESTS, FLAG_TILE_SPLIT = 1, 2, 4, 8
ACCUM_CHANNELS, ACCUM_FRAC_BITS = 9, 28

# every symbol include/tipe_rt.h declares
ABI_SYMBOLS = [
    "trt_default_overrides", "trt_init", "trt_destroy", "trt_last_error", "trt_version",
    "trt_scene_set", "trt_sky_set", "trt_render", "trt_primary", "trt_primary_fp32", "trt_accumulate", "trt_resolve", "trt_resolve_device", "trt_resolve_hdr", "trt_resolve_hdr_device",
    "trt_nccl_version", "trt_nccl_unique_id", "trt_nccl_comm_init", "trt_nccl_comm_destroy", "trt_reduce_nccl", "trt_device_of", "trt_render_distributed", "trt_host_alloc", "trt_host_free",
    "trt_stream", "trt_synchronize", "trt_bvh_info_get", "trt_debug_schedule",
    "trt_kat_hit_sphere", "trt_kat_hit_triangle", "trt_kat_refract", "trt_kat_philox", "trt_kat_tonemap",
]


class TrtError(RuntimeError):
    pass


def build_extension(verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (csrc/Makefile), in-tree."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise TrtError("building libtipe_rt.so failed")
    return LIB_PATH


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
        ("flags", C.c_int), ("tile_part", C.c_int),
    ]


class TrtStats(C.Structure):
    _fields_ = [
        ("paths", C.c_uint64), ("rays", C.c_uint64),
        ("box_tests", C.c_uint64), ("tri_tests", C.c_uint64), ("sphere_tests", C.c_uint64),
        ("ms_trace", C.c_double), ("ms_reduce_resolve", C.c_double), ("ms_d2h", C.c_double), ("ms_total", C.c_double),
        ("kernel_launches", C.c_int), ("reserved", C.c_int),
        ("rays_skipped", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class TrtBvhInfo(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("n_leaves", C.c_int), ("max_depth", C.c_int), ("n_twins", C.c_int),
                ("sah_cost", C.c_double), ("ms_build", C.c_double)]


def make_params(rp: RenderParams) -> TrtParams:
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
    p.tile_part = rp.tile_part
    return p


def load_library() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise TrtError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.trt_last_error.restype = C.c_char_p
    lib.trt_stream.restype = C.c_void_p
    for name in ABI_SYMBOLS:
        getattr(lib, name)
    return lib


def host_arrays(n_pixels: int, lib: Optional[C.CDLL] = None) -> dict:
    """canva / albedo / normal [n_pixels, 3] float64 in PAGE-LOCKED host memory (trt_host_alloc): the GPUs write them directly."""
    lib = lib or load_library()
    out = {}
    for k in ("canva", "albedo", "normal"):
        p = C.c_void_p()
        rc = lib.trt_host_alloc(C.c_size_t(n_pixels * 24), C.byref(p))
        if rc != 0:
            raise TrtError(f"trt_host_alloc failed ({rc})")
        buf = (C.c_double * (3 * n_pixels)).from_address(p.value)
        out[k] = np.frombuffer(buf, dtype=np.float64).reshape(n_pixels, 3)
    return out          # (kept for the life of the process; trt_host_free(ptr) releases one)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Renderer:
    """One trt_ctx.  devices: list of CUDA ordinals (default [0]); more than one device renders
    disjoint sample ranges per device and sums them on the first through NVLink peer loads."""

    def __init__(self, devices: Optional[list] = None):
        self.lib = load_library()
        devices = [0] if devices is None else list(devices)
        ids = (C.c_int * len(devices))(*devices)
        self.ctx = C.c_void_p()
        rc = self.lib.trt_init(ids, C.c_int(len(devices)), C.byref(self.ctx))
        if rc != 0:
            msg = self.lib.trt_last_error(None)
            raise TrtError(f"trt_init failed ({rc}): {msg.decode() if msg else ''}")
        self.devices = devices
        self._scene: Optional[Scene] = None

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.trt_last_error(self.ctx)
            raise TrtError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.trt_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene ---------------------------------------------------------------------------------
    def scene_set(self, sc: Scene):
        sph = np.ascontiguousarray(sc.spheres)
        tri = np.ascontiguousarray(sc.triangles)
        mats = np.ascontiguousarray(sc.mat_list)
        quel = np.ascontiguousarray(sc.quel_mat_pour_tri, dtype=np.int32)
        if sc.overrides is None:
            ov, n_ov = None, -1
        else:
            ov = np.ascontiguousarray(sc.overrides)
            n_ov = ov.shape[0]
        rc = self.lib.trt_scene_set(self.ctx, _ptr(sph), C.c_int(sph.shape[0]), _ptr(tri), C.c_int(tri.shape[0]),
                                    _ptr(mats), C.c_int(sc.tex_w), C.c_int(sc.tex_h), C.c_int(sc.nb_materials),
                                    _ptr(quel), _ptr(ov), C.c_int(n_ov))
        self._check(rc, "trt_scene_set")
        if sc.sky_mat_list is not None:
            sky = np.ascontiguousarray(sc.sky_mat_list)
            assert sky.shape[0] == sc.sky_w * sc.sky_h
            self._check(self.lib.trt_sky_set(self.ctx, _ptr(sky), C.c_int(sc.sky_w), C.c_int(sc.sky_h)), "trt_sky_set")
        else:
            self._check(self.lib.trt_sky_set(self.ctx, None, C.c_int(0), C.c_int(0)), "trt_sky_set")
        self._scene = sc

    def bvh_info(self) -> dict:
        info = TrtBvhInfo()
        self._check(self.lib.trt_bvh_info_get(self.ctx, C.byref(info)), "trt_bvh_info_get")
        return {k: getattr(info, k) for k, _ in info._fields_}

    # -- the drop-in call (main.c:402-453) -------------------------------------------------------
    def render(self, rp: RenderParams, out: Optional[dict] = None) -> dict:
        """-> dict(canva, albedo, normal: float64 [H*W, 3]; stats). `out` may supply the three arrays."""
        n = rp.width * rp.height
        canva = out["canva"] if out else np.empty((n, 3))
        albedo = out["albedo"] if out else np.empty((n, 3))
        normal = out["normal"] if out else np.empty((n, 3))
        p = make_params(rp)
        st = TrtStats()
        self._check(self.lib.trt_render(self.ctx, C.byref(p), _ptr(canva), _ptr(albedo), _ptr(normal), C.byref(st)), "trt_render")
        return dict(canva=canva, albedo=albedo, normal=normal, stats=st.as_dict())

    def primary(self, rp: RenderParams) -> dict:
        n = rp.width * rp.height
        t = np.empty(n); ids = np.empty(n, np.int32); normal = np.empty((n, 3)); albedo = np.empty((n, 3))
        p = make_params(rp)
        self._check(self.lib.trt_primary(self.ctx, C.byref(p), _ptr(t), _ptr(ids), _ptr(normal), _ptr(albedo)), "trt_primary")
        return dict(t=t, id=ids, normal=normal, albedo=albedo)

    def primary_fp32(self, rp: RenderParams) -> dict:
        """First hit of the pixel-centre rays through the path tracer's own FP32 / BVH arithmetic."""
        n = rp.width * rp.height
        t = np.empty(n, np.float32); ids = np.empty(n, np.int32)
        p = make_params(rp)
        self._check(self.lib.trt_primary_fp32(self.ctx, C.byref(p), _ptr(t), _ptr(ids)), "trt_primary_fp32")
        return dict(t=t, id=ids)

    # -- device-resident pieces ------------------------------------------------------------------
    def accumulate(self, rp: RenderParams, d_accum_ptr: int) -> dict:
        """Adds rp.spp samples starting at rp.sample_offset into the int64 [9][H*W] device buffer."""
        p = make_params(rp)
        st = TrtStats()
        self._check(self.lib.trt_accumulate(self.ctx, C.byref(p), C.c_void_p(d_accum_ptr), C.byref(st)), "trt_accumulate")
        return st.as_dict()

    def resolve(self, rp: RenderParams, d_accum_ptr: int, out: Optional[dict] = None) -> dict:
        n = rp.width * rp.height
        canva = out["canva"] if out else np.empty((n, 3))
        albedo = out["albedo"] if out else np.empty((n, 3))
        normal = out["normal"] if out else np.empty((n, 3))
        p = make_params(rp)
        self._check(self.lib.trt_resolve(self.ctx, C.byref(p), C.c_void_p(d_accum_ptr), _ptr(canva), _ptr(albedo), _ptr(normal)), "trt_resolve")
        return dict(canva=canva, albedo=albedo, normal=normal)

    def resolve_hdr(self, rp: RenderParams, d_accum_ptr: int) -> dict:
        """Un-quantised float32 means for an external denoiser (linear radiance, albedo, normal)."""
        n = rp.width * rp.height
        rad, alb, nrm = (np.empty((n, 3), np.float32) for _ in range(3))
        p = make_params(rp)
        self._check(self.lib.trt_resolve_hdr(self.ctx, C.byref(p), C.c_void_p(d_accum_ptr), _ptr(rad), _ptr(alb), _ptr(nrm)), "trt_resolve_hdr")
        return dict(radiance=rad, albedo=alb, normal=nrm)

    def resolve_hdr_device(self, rp: RenderParams, d_accum_ptr: int, d_radiance: int, d_albedo: int, d_normal: int):
        """Float32 means into device buffers (3*W*H floats each), for a denoiser that reads device memory."""
        p = make_params(rp)
        self._check(self.lib.trt_resolve_hdr_device(self.ctx, C.byref(p), C.c_void_p(d_accum_ptr), C.c_void_p(d_radiance),
                                                    C.c_void_p(d_albedo), C.c_void_p(d_normal)), "trt_resolve_hdr_device")

    def resolve_device(self, rp: RenderParams, d_accum_ptr: int, d_canva: int, d_albedo: int, d_normal: int):
        p = make_params(rp)
        self._check(self.lib.trt_resolve_device(self.ctx, C.byref(p), C.c_void_p(d_accum_ptr), C.c_void_p(d_canva),
                                                C.c_void_p(d_albedo), C.c_void_p(d_normal)), "trt_resolve_device")

    # -- one process per GPU: NCCL inside the library (nccl_glue.cu) ----------------------------------
    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.trt_nccl_unique_id(buf)
        if rc != 0:
            msg = self.lib.trt_last_error(None)
            raise TrtError(f"trt_nccl_unique_id failed ({rc}): {msg.decode() if msg else ''}")
        return buf.raw

    def nccl_comm_init(self, n_ranks: int, rank: int, unique_id: bytes) -> int:
        comm = C.c_void_p()
        self._check(self.lib.trt_nccl_comm_init(self.ctx, C.c_int(n_ranks), C.c_int(rank), C.c_char_p(unique_id), C.byref(comm)), "trt_nccl_comm_init")
        return int(comm.value or 0)

    def nccl_comm_destroy(self, comm: int):
        self.lib.trt_nccl_comm_destroy(C.c_void_p(comm))

    def reduce_nccl(self, comm: int, d_accum_ptr: int, count: int, root: int = 0):
        """ncclReduce(int64, SUM) of the sample sums onto `root` (root < 0: all-reduce), asynchronous on the ctx's stream."""
        self._check(self.lib.trt_reduce_nccl(self.ctx, C.c_void_p(comm), C.c_void_p(d_accum_ptr), C.c_size_t(count), C.c_int(root)), "trt_reduce_nccl")

    def render_distributed(self, rp: RenderParams, comm: int, rank: int, n_ranks: int, out: Optional[dict] = None) -> dict:
        """trt_render_distributed: this rank's share of rp.spp, NCCL reduce, rank 0 gets the image (other ranks: stats only)."""
        n = rp.width * rp.height
        res = {}
        if rank == 0:
            res = dict(canva=out["canva"] if out else np.empty((n, 3)), albedo=out["albedo"] if out else np.empty((n, 3)),
                       normal=out["normal"] if out else np.empty((n, 3)))
        p = make_params(rp)
        st = TrtStats()
        self._check(self.lib.trt_render_distributed(self.ctx, C.byref(p), C.c_void_p(comm), C.c_int(rank), C.c_int(n_ranks),
                                                    _ptr(res.get("canva")), _ptr(res.get("albedo")), _ptr(res.get("normal")), C.byref(st)),
                    "trt_render_distributed")
        res["stats"] = st.as_dict()
        return res

    def debug_schedule(self) -> dict:
        out = (C.c_uint64 * 7)()
        self._check(self.lib.trt_debug_schedule(self.ctx, out), "trt_debug_schedule")
        return dict(zip(("trips", "inside", "node", "leaf", "sessions", "hk", "consume"), [int(x) for x in out]))

    def stream_ptr(self) -> int:
        return int(self.lib.trt_stream(self.ctx) or 0)

    def synchronize(self):
        self._check(self.lib.trt_synchronize(self.ctx), "trt_synchronize")

    # -- known-answer entry points -----------------------------------------------------------------
    def kat_hit_sphere(self, center_radius, rays):
        cr = np.ascontiguousarray(center_radius, np.float32); rays = np.ascontiguousarray(rays, np.float32)
        n = cr.shape[0]
        hit = np.zeros(n, np.int32); dst = np.zeros(n, np.float32); pt = np.zeros((n, 3), np.float32); nr = np.zeros((n, 3), np.float32)
        self._check(self.lib.trt_kat_hit_sphere(self.ctx, C.c_int(n), _ptr(cr), _ptr(rays), _ptr(hit), _ptr(dst), _ptr(pt), _ptr(nr)), "trt_kat_hit_sphere")
        return hit.astype(bool), dst, pt, nr

    def kat_hit_triangle(self, abc, rays):
        abc = np.ascontiguousarray(abc, np.float32); rays = np.ascontiguousarray(rays, np.float32)
        n = abc.shape[0]
        hit = np.zeros(n, np.int32); out = np.zeros((n, 3), np.float32)
        self._check(self.lib.trt_kat_hit_triangle(self.ctx, C.c_int(n), _ptr(abc), _ptr(rays), _ptr(hit), _ptr(out)), "trt_kat_hit_triangle")
        return hit.astype(bool), out

    def kat_refract(self, v_n_n1_n2):
        a = np.ascontiguousarray(v_n_n1_n2, np.float32)
        out = np.zeros((a.shape[0], 3), np.float32)
        self._check(self.lib.trt_kat_refract(self.ctx, C.c_int(a.shape[0]), _ptr(a), _ptr(out)), "trt_kat_refract")
        return out

    def kat_philox(self, ctr_key):
        a = np.ascontiguousarray(ctr_key, np.uint32)
        out = np.zeros((a.shape[0], 4), np.uint32)
        self._check(self.lib.trt_kat_philox(self.ctx, C.c_int(a.shape[0]), _ptr(a), _ptr(out)), "trt_kat_philox")
        return out

    def kat_tonemap(self, sums, spp):
        a = np.ascontiguousarray(sums, np.int64)
        out = np.zeros(a.shape[0], np.float64)
        self._check(self.lib.trt_kat_tonemap(self.ctx, C.c_int(a.shape[0]), _ptr(a), C.c_int(spp), _ptr(out)), "trt_kat_tonemap")
        return out


# -------------------------------------------------------------------------------------------------
# The reference's own operator for this path: fill_canva(ThreadData*) (main.c:22-46, 245-284).
# -------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class ThreadData:
    """Field-for-field mirror of struct ThreadData (main.c:22-46).  tex_list / sky_mat_list /
    sky_width / sky_height exist in the reference but are never read on this path."""
    start_row: int
    end_row: int
    canva: np.ndarray
    albedo_tab: np.ndarray
    normal_tab: np.ndarray
    mat_list: np.ndarray
    cam: np.ndarray
    largeur_image: int
    hauteur_image: int
    tex_width: int
    tex_height: int
    quelMatPourTri: np.ndarray
    nbRayonParPixel: int
    nbRebondMax: int
    sphere_list: np.ndarray
    triangle_list: np.ndarray
    nbSpheres: int
    nbTriangles: int
    ouverture_x: int = 0          # int in the reference: doubles are truncated on assignment (main.c:429-433)
    ouverture_y: int = 0
    focus_distance: int = 3
    AO_intensity: int = 2
    useAO: bool = False
    total_pixels: int = 0
    seed: int = 1010              # not in the reference (it uses the global rand() stream)


def fill_canva(data: ThreadData, renderer: Optional[Renderer] = None) -> None:
    """fill_canva(void* arg) of main.c:245-284: renders rows end_row..start_row (inclusive, row 0 =
    bottom) into data.canva / albedo_tab / normal_tab.  The whole frame is traced on the GPU and the
    requested band copied out; calling it once with start_row=H-1, end_row=0 is the drop-in use."""
    own = renderer is None
    r = renderer or Renderer()
    try:
        nb_mat = 0
        if data.nbTriangles > 0:
            nb_mat = data.mat_list.shape[0] // (data.tex_width * data.tex_height)
        sc = Scene(name="thread_data", spheres=data.sphere_list[:data.nbSpheres], triangles=data.triangle_list[:data.nbTriangles],
                   mat_list=data.mat_list, tex_w=data.tex_width, tex_h=data.tex_height, nb_materials=nb_mat,
                   quel_mat_pour_tri=np.asarray(data.quelMatPourTri, np.int32))
        r.scene_set(sc)
        rp = RenderParams(width=data.largeur_image, height=data.hauteur_image, spp=data.nbRayonParPixel,
                          bounces=data.nbRebondMax, cam=data.cam, use_ao=bool(data.useAO),
                          ao_intensity=float(int(data.AO_intensity)), focus_distance=float(int(data.focus_distance)),
                          aperture=(float(int(data.ouverture_x)), float(int(data.ouverture_y))), seed=data.seed)
        res = r.render(rp)
        W = data.largeur_image
        lo, hi = data.end_row * W, (data.start_row + 1) * W
        for name, dst in (("canva", data.canva), ("albedo", data.albedo_tab), ("normal", data.normal_tab)):
            dst.reshape(-1, 3)[lo:hi] = res[name][lo:hi]
    finally:
        if own:
            r.close()
