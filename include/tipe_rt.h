/* tipe_rt.h — C ABI of libtipe_rt.so, the B200 (sm_100a) replacement for the per-pixel
 * path-tracing hot path of xelema/tipe-raytracer's main.c.
 *
 * What it replaces: the pthread fan-out `for t in NUM_THREADS: pthread_create(fill_canva)` +
 * `pthread_join` of main.c:402-453, i.e. everything reached from fill_canva (main.c:245-284):
 * get_ray (camera.h:42-54), tracer (main.c:118-242), closest_hit (main.c:52-92),
 * ambient_occlusion (main.c:94-116), hit_sphere (sphere.h:13-47), hit_triangle (mesh.h:70-94),
 * tri_uvmapping (texture.h:44-90), the refraction stack (pile.h), write_color_canva
 * (rtutility.h:56-71).  Everything else in main.c (OBJ/MTL/PPM loading, init_camera, the
 * OIDN post-pass, the PPM writer) stays host C and feeds / consumes this interface.
 *
 * Plain C, plain pointers and sizes.  The scene structs below are layout-identical to the
 * reference's (sizes and offsets asserted at the bottom), so a caller inside the reference
 * tree passes its own `sphere*`, `triangle*`, `material*`, `camera` by cast — see
 * INTEGRATION.md for the exact stub.
 *
 * Conventions (all from the reference):
 *   - pixel_index = j*W + i with j = 0 the BOTTOM image row (main.c:261, file written j=H-1..0, main.c:460);
 *   - canva holds integer values 0..255 stored in doubles (write_color_canva), albedo_tab /
 *     normal_tab hold per-pixel means (main.c:275-279);
 *   - object id: sphere k -> k, triangle k -> nbSpheres + k, miss -> -1;
 *   - ThreadData truncates AO_intensity, focus_distance, ouverture_x/y to int (main.c:42-43):
 *     this ABI takes doubles, a caller that wants main.c's behaviour passes (double)(int)x.
 *
 * Errors: every call returns TRT_OK (0) or a negative code; the text is in trt_last_error().
 * The library never calls exit().  There is no CPU fallback: without a CUDA device trt_init fails.
 */
#ifndef TIPE_RT_H
#define TIPE_RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#  define TRT_API __attribute__((visibility("default")))
#else
#  define TRT_API
#endif

/* ---- scene types: byte-compatible with the reference ---------------------------------- */
typedef struct trt_vec3 { double e[3]; } trt_vec3;                 /* vec3.h:7-9 (vec3/point3/color) */

typedef struct trt_material {                                       /* hitinfo.h:6-13 */
    trt_vec3 diffuseColor;
    trt_vec3 emissionColor;
    double   emissionStrength;
    double   reflectionStrength;
    double   alpha;
    double   materialIndex;
} trt_material;

typedef struct trt_sphere {                                         /* sphere.h:7-11 */
    trt_vec3     center;
    double       radius;
    trt_material mat;
} trt_sphere;

typedef struct trt_uv { double u, v; } trt_uv;                      /* mesh.h:9-12 */

typedef struct trt_triangle {                                       /* mesh.h:14-22 */
    trt_vec3     A, B, C;
    trt_material mat;      /* ignored: the reference always overwrites it from the texture (main.c:86-88) */
    trt_uv       uvA, uvB, uvC;
} trt_triangle;

typedef struct trt_camera {                                         /* camera.h:10-15 */
    trt_vec3 origin, horizontal, vertical, coin_bas_gauche;
} trt_camera;

/* ---- per-material overrides (texture.h:71-87 hard-codes three of them) ------------------ */
enum {
    TRT_OV_EMISSION_COLOR    = 1u << 0,
    TRT_OV_EMISSION_STRENGTH = 1u << 1,
    TRT_OV_REFLECTION        = 1u << 2,
    TRT_OV_ALPHA             = 1u << 3,
    TRT_OV_MATERIAL_INDEX    = 1u << 4
};
typedef struct trt_mat_override {
    int      material_ordinal;       /* numero_mat = quelMatPourTri[tri] (texture.h:64) */
    unsigned mask;                   /* which of the fields below replace the texel's */
    double   emissionColor[3];
    double   emissionStrength;
    double   reflectionStrength;
    double   alpha;
    double   materialIndex;
} trt_mat_override;

/* Fills `out` (room for >= 3) with the reference's built-in table (texture.h:71-87):
 * ordinal 1 -> emission WHITE x1.85, alpha 1; ordinal 4 -> alpha .6, index 1.33, refl .93;
 * ordinal 3 -> alpha .1, index 1.50, refl .3.  Returns the count (3). */
TRT_API int trt_default_overrides(trt_mat_override* out, int capacity);

/* ---- render parameters: the knobs of main.c:293-317 as carried by ThreadData ----------- */
enum {
    TRT_FLAG_COUNT_RAYS  = 1 << 0,   /* fill trt_stats.rays (one 64-bit atomic per warp per item) */
    TRT_FLAG_BRUTE_FORCE = 1 << 1,   /* triangles by linear scan like main.c:81-90 instead of the BVH */
    TRT_FLAG_COUNT_TESTS = 1 << 2,   /* also count box / triangle / sphere tests (slower; for the roofline) */
    TRT_FLAG_TILE_SPLIT  = 1 << 3    /* trt_render on several devices: split the image by interleaved 8x4 tiles (device k takes the
                                        tiles k mod N with ALL samples) instead of by sample ranges; same image bit for bit */
};
typedef struct trt_params {
    int        largeur_image, hauteur_image;     /* W, H */
    int        nbRayonParPixel;                  /* samples rendered by THIS call */
    int        nbRebondMax;                      /* loop trip count of main.c:132 */
    int        useAO;
    double     AO_intensity;
    double     focus_distance, ouverture_x, ouverture_y;
    trt_camera cam;                              /* as produced by init_camera (camera.h:21-40) */
    uint64_t   seed;                             /* Philox key; same seed => bit-identical image */
    int        sample_offset;                    /* global index of this call's first sample (spp split) */
    int        sample_total;                     /* divisor used by the resolve; 0 => nbRayonParPixel */
    int        flags;
    int        tile_part;                        /* 0: the whole image.  Else part | n_parts << 16: this call renders only the 8x4
                                                    tiles t with t mod n_parts == part (tile split across processes) */
} trt_params;

typedef struct trt_stats {
    uint64_t paths;            /* W*H*nbRayonParPixel traced by this call (all devices of the ctx) */
    uint64_t rays;             /* closest_hit evaluations: primary + bounce + AO (TRT_FLAG_COUNT_RAYS) */
    uint64_t box_tests, tri_tests, sphere_tests;   /* TRT_FLAG_COUNT_TESTS */
    double   ms_trace;         /* path-tracing kernel(s), CUDA events, max over the ctx's devices */
    double   ms_reduce_resolve;/* peer reduce + tonemap kernel */
    double   ms_d2h;           /* device -> caller buffers */
    double   ms_total;         /* whole call, host clock */
    int      kernel_launches;  /* kernels of this library launched by the call */
    int      reserved;
    uint64_t rays_skipped;     /* of `rays`: ambient-occlusion rays of a path's LAST bounce, which main.c:220 traces although their
                                  result (a factor on rayColor) is never read; counted for parity with the reference, not traced */
} trt_stats;

typedef struct trt_ctx trt_ctx;

/* ---- life cycle ------------------------------------------------------------------------ */
/* device_ids == NULL => devices 0..n_devices-1.  n_devices > 1 renders disjoint sample ranges
 * on each device and sums them on device_ids[0] through peer (NVLink) loads. */
TRT_API int         trt_init(const int* device_ids, int n_devices, trt_ctx** out);
TRT_API void        trt_destroy(trt_ctx* ctx);
TRT_API const char* trt_last_error(const trt_ctx* ctx);     /* ctx may be NULL: last trt_init error */
TRT_API int         trt_version(void);

/* ---- scene ----------------------------------------------------------------------------- */
/* Copies and converts the caller's arrays (read-only, not retained) to the device layout and
 * builds the BVH on the GPU.  mat_list has nbMaterials*tex_w*tex_h entries (texture.h:64-68).
 * overrides == NULL && n_overrides < 0 => the reference's built-in table. */
TRT_API int trt_scene_set(trt_ctx* ctx,
                          const trt_sphere* sphere_list, int nbSpheres,
                          const trt_triangle* triangle_list, int nbTriangles,
                          const trt_material* mat_list, int tex_width, int tex_height, int nbMaterials,
                          const int* quelMatPourTri,
                          const trt_mat_override* overrides, int n_overrides);

/* Sky environment: enables the lat-long lookup the reference keeps commented out at main.c:64-71 (sphere_uvmapping,
 * texture.h:92-112; sky_mat_list as loaded by create_mat_list, texture.h:145-173: sky_w*sky_h materials whose
 * diffuseColor is the texel, index y*sky_w + x with row 0 = bottom PPM row).  The LAST sphere of the scene then emits
 * the texel under the hit point (emissionColor = texel, alpha = 1), like ThreadData.sky_mat_list (main.c:29,34) would.
 * NULL switches it off (the default: main.c as shipped).  Independent of trt_scene_set; kept until changed. */
TRT_API int trt_sky_set(trt_ctx* ctx, const trt_material* sky_mat_list, int sky_width, int sky_height);

/* ---- the drop-in call: replaces main.c:402-453 ------------------------------------------- */
/* canva / albedo_tab / normal_tab: caller-owned host arrays of W*H colors; stats may be NULL. */
TRT_API int trt_render(trt_ctx* ctx, const trt_params* params,
                       trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab, trt_stats* stats);

/* Primary-hit buffers through pixel centres (u = i/(W-1), v = j/(H-1), no jitter, no aperture):
 * FP64, linear scan in the reference's order, so ids match the CPU path exactly. Any output may be NULL. */
TRT_API int trt_primary(trt_ctx* ctx, const trt_params* params,
                        double* t, int* id, trt_vec3* normal, trt_vec3* albedo);

/* The same pixel-centre rays through the arithmetic trt_render actually uses: FP32 camera ray, FP32 sphere / triangle tests, the
 * BVH walk.  Not bit-exact against the FP64 pass by construction (a ray through a silhouette may fall on the other side); the
 * tests measure how often.  t in units of |d| (float), id in the caller's numbering.  Either output may be NULL. */
TRT_API int trt_primary_fp32(trt_ctx* ctx, const trt_params* params, float* t, int* id);

/* ---- device-resident pieces (what trt_render is made of; used by the multi-process path
 *      where the sample sums cross GPUs with an NCCL reduce, and by bench.py) -------------- */
#define TRT_ACCUM_CHANNELS 9      /* radiance rgb, albedo rgb, normal xyz */
#define TRT_ACCUM_FRAC_BITS 28    /* sums are int64 fixed point, value * 2^28 per sample */

/* Adds this call's samples into d_accum (device pointer on the ctx's first device, int64,
 * planar [9][W*H], must be zeroed by the caller before the first call of an image). */
TRT_API int trt_accumulate(trt_ctx* ctx, const trt_params* params, int64_t* d_accum, trt_stats* stats);

/* Tonemaps device sums into host (or, _device, device) color arrays like main.c:275-279. */
TRT_API int trt_resolve(trt_ctx* ctx, const trt_params* params, const int64_t* d_accum,
                        trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab);
TRT_API int trt_resolve_device(trt_ctx* ctx, const trt_params* params, const int64_t* d_accum,
                               double* d_canva, double* d_albedo, double* d_normal);

/* Un-quantised float3 means (linear radiance before the sqrt tonemap, albedo, normal), W*H*3 floats each, in the
 * layout OIDN's OIDN_FORMAT_FLOAT3 images use: the hand-off for an external denoiser, instead of the 8-bit round
 * trip of denoiser.h:43-45. The denoiser itself (denoiser.h:31-91) stays outside this library. */
TRT_API int trt_resolve_hdr(trt_ctx* ctx, const trt_params* params, const int64_t* d_accum,
                            float* radiance, float* albedo, float* normal);
/* Same, into DEVICE buffers on the ctx's first device (3*W*H floats each): what a GPU denoiser build (oidnNewSharedBuffer on
 * device memory; denoiser.h:41-59 fills host arrays) consumes without the frame ever leaving HBM.  Returns when the buffers are ready. */
TRT_API int trt_resolve_hdr_device(trt_ctx* ctx, const trt_params* params, const int64_t* d_accum,
                                   float* d_radiance, float* d_albedo, float* d_normal);

/* ---- the collective of the multi-process path (one process per GPU; SURVEY §8e, BASELINE config 5) -------------------
 * Every rank renders its sample range with trt_accumulate, then trt_reduce_nccl adds the int64 sums of all ranks onto
 * `root` (root < 0: onto every rank, ncclAllReduce) with ncclReduce(ncclInt64, ncclSum) on the ctx's stream, and the root
 * calls trt_resolve.  Asynchronous like NCCL itself: ordered with the ctx's other work, complete after trt_synchronize or
 * the next blocking call.  `comm` is an ncclComm_t: the caller's own, or one made by trt_nccl_comm_init from a 128-byte
 * ncclUniqueId that rank 0 obtained with trt_nccl_unique_id and handed to the other ranks by any means (file, socket,
 * MPI).  NCCL is loaded at run time (libnccl.so.2); the reference has no counterpart (pthreads, main.c:402-453). */
#define TRT_NCCL_UNIQUE_ID_BYTES 128
TRT_API int trt_nccl_version(void);                           /* 0 when NCCL cannot be loaded */
TRT_API int trt_nccl_unique_id(void* out128);
TRT_API int trt_nccl_comm_init(trt_ctx* ctx, int n_ranks, int rank, const void* unique_id128, void** comm_out);
TRT_API int trt_nccl_comm_destroy(void* comm);
TRT_API int trt_reduce_nccl(trt_ctx* ctx, void* comm, int64_t* d_accum, size_t count, int root);
/* The drop-in call for ONE PROCESS PER GPU: every rank calls it with the same params; rank r traces the samples
 * [r*spp/N, (r+1)*spp/N) of every pixel (or, with TRT_FLAG_TILE_SPLIT, all samples of the tiles r mod N), the sums meet on
 * rank 0 through trt_reduce_nccl, rank 0 tonemaps into canva / albedo_tab / normal_tab (ignored on the other ranks, may be
 * NULL there).  The image is bit-identical to trt_render on one GPU.  stats describe this rank's share. */
TRT_API int trt_render_distributed(trt_ctx* ctx, const trt_params* params, void* nccl_comm, int rank, int n_ranks,
                                   trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab, trt_stats* stats);
/* Optional: page-locked host memory for canva / albedo_tab / normal_tab (instead of the three malloc()s of main.c:385-400).  When all
 * three arrays of a render / resolve call are page-locked (these, or cudaHostAlloc / cudaHostRegister of the caller's own), the GPUs
 * write them directly over PCIe - with N devices N bands at once - instead of going through the library's staging buffer and
 * host copy threads.  Ordinary malloc()ed arrays keep working as before. */
TRT_API int  trt_host_alloc(size_t bytes, void** out);
TRT_API void trt_host_free(void* p);
/* CUDA ordinal of the ctx's k-th device (-1 if out of range) */
TRT_API int trt_device_of(const trt_ctx* ctx, int k);

/* Stream the ctx's first device works on (a cudaStream_t), for callers that interleave their own work. */
TRT_API void* trt_stream(trt_ctx* ctx);
TRT_API int   trt_synchronize(trt_ctx* ctx);

/* BVH facts of the current scene: nodes, leaves, max depth, root surface-area cost. */
typedef struct trt_bvh_info { int n_nodes, n_leaves, max_depth, n_twins; double sah_cost; double ms_build; } trt_bvh_info;
TRT_API int trt_bvh_info_get(trt_ctx* ctx, trt_bvh_info* out);
/* Scheduling counters of the last TRT_FLAG_COUNT_TESTS launch (tools/): BVH-loop trips per warp, lanes inside / at a node /
   at a leaf summed over trips, loop sessions, housekeeping passes, lanes consuming a hit summed over passes. */
TRT_API int trt_debug_schedule(trt_ctx* ctx, uint64_t* out7);

/* ---- device known-answer entry points (FP32 leaf functions of the path-tracing kernel,
 *      run on the GPU over n independent inputs; outputs are float) -------------------------- */
TRT_API int trt_kat_hit_sphere(trt_ctx* ctx, int n, const float* center_radius /*[n][4]*/,
                               const float* ray_o_d /*[n][6]*/, int* hit, float* dst, float* point, float* normal);
TRT_API int trt_kat_hit_triangle(trt_ctx* ctx, int n, const float* abc /*[n][9]*/,
                                 const float* ray_o_d /*[n][6]*/, int* hit, float* dst_u_v /*[n][3]*/);
TRT_API int trt_kat_refract(trt_ctx* ctx, int n, const float* v_n_n1_n2 /*[n][8]*/, float* out /*[n][3]*/);
TRT_API int trt_kat_philox(trt_ctx* ctx, int n, const uint32_t* ctr_key /*[n][6]*/, uint32_t* out /*[n][4]*/);
TRT_API int trt_kat_tonemap(trt_ctx* ctx, int n, const int64_t* sums /*[n]*/, int spp, double* out /*[n]*/);

/* ---- layout contract -------------------------------------------------------------------- */
#if defined(__cplusplus)
#  define TRT_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#  define TRT_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
TRT_STATIC_ASSERT(sizeof(trt_vec3) == 24, "vec3 is 3 doubles");
TRT_STATIC_ASSERT(sizeof(trt_material) == 80, "material layout (hitinfo.h:6-13)");
TRT_STATIC_ASSERT(offsetof(trt_material, alpha) == 64, "material.alpha offset");
TRT_STATIC_ASSERT(sizeof(trt_sphere) == 112 && offsetof(trt_sphere, mat) == 32, "sphere layout (sphere.h:7-11)");
TRT_STATIC_ASSERT(sizeof(trt_triangle) == 200 && offsetof(trt_triangle, uvA) == 152, "triangle layout (mesh.h:14-22)");
TRT_STATIC_ASSERT(sizeof(trt_camera) == 96, "camera layout (camera.h:10-15)");

#define TRT_OK                 0
#define TRT_ERR_INVALID       -1
#define TRT_ERR_NO_DEVICE     -2
#define TRT_ERR_CUDA          -3
#define TRT_ERR_NO_SCENE      -4
#define TRT_ERR_OOM           -5
#define TRT_ERR_NCCL          -6

#ifdef __cplusplus
}
#endif
#endif /* TIPE_RT_H */
