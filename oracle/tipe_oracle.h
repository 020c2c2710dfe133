/* TEST INFRASTRUCTURE — CPU restatement ("port") of the reference hot path. Not shipped.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product library (libtipe_rt.so) never links, loads or calls it.
 *
 * Pinning: in TRT_ORACLE_RNG_LIBC mode every function here reproduces the reference's FP64
 * arithmetic in the reference's operation order and draws from glibc rand() in the reference's
 * call order, so a 1-thread render is BIT-IDENTICAL to oracle/_ref (the reference's own main.c
 * compiled as a library) — tests/test_oracle_pin.py checks that, plus the golden vectors under
 * tests/golden/ that were generated from oracle/_ref.
 * In TRT_ORACLE_RNG_PHILOX mode the only change is where the uniforms come from: the
 * counter-based stream the GPU uses (Philox4x32-10, key = seed, counter = (pixel, sample, slot)),
 * which makes GPU-vs-oracle comparisons sample-exact instead of merely statistical.
 */
#ifndef TIPE_ORACLE_H
#define TIPE_ORACLE_H

#include "../include/tipe_rt.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TRT_ORACLE_RNG_LIBC   0
#define TRT_ORACLE_RNG_PHILOX 1

typedef struct oracle_scene {
    const trt_sphere*   spheres;   int nbSpheres;
    const trt_triangle* tris;      int nbTriangles;
    const trt_material* mat_list;  int tex_w, tex_h;
    const int*          quelMatPourTri;
    const trt_mat_override* overrides; int n_overrides;   /* NULL / <0 => reference built-ins */
    /* Sky environment: the lat-long lookup the reference keeps commented out at main.c:65-71.
     * NULL => exactly main.c as shipped.  Non-NULL => the LAST sphere takes emissionColor from the sky texel, alpha = 1. */
    const trt_material* sky_mat_list; int sky_w, sky_h;
} oracle_scene;

typedef struct oracle_stats {
    uint64_t paths, rays, sphere_tests, tri_tests;
    double   seconds;
} oracle_stats;

/* fill_canva (main.c:245-284) over the whole image. rng_mode LIBC forces 1 thread. Any of the
 * outputs may be NULL. sums9 (optional, [W*H][9]) receives the un-tonemapped per-pixel sums. */
int oracle_render(const oracle_scene* sc, const trt_params* p, int rng_mode, int threads,
                  trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab, double* sums9,
                  oracle_stats* stats);

/* Primary-hit buffers through pixel centres, ids captured like main.c:59-90. */
int oracle_primary(const oracle_scene* sc, const trt_params* p,
                   double* t, int* id, trt_vec3* normal, trt_vec3* albedo);

/* leaf functions */
void oracle_init_camera(const double* origin, const double* target, const double* up,
                        double vfov, double ratio, trt_camera* out);                 /* camera.h:21-40 */
void oracle_get_ray(double u, double v, const trt_camera* cam, double focus, double dx, double dy,
                    double* origin, double* dir);                                    /* camera.h:42-54 */
int  oracle_hit_sphere(const double* center, double radius, const double* o, const double* d,
                       double* dst, double* point, double* normal);                  /* sphere.h:13-47 */
int  oracle_hit_triangle(const trt_triangle* tri, const double* o, const double* d,
                         double* dst, double* point, double* normal);                /* mesh.h:70-94 */
void oracle_tri_uvmapping(const oracle_scene* sc, int tri_index, const double* point, const double* normal,
                          trt_material* out);                                        /* texture.h:44-90 */
int  oracle_sphere_uvmapping(const double* center, double radius, const double* point, int sky_w, int sky_h); /* texture.h:92-112, texel index */
void oracle_reflect(const double* v, const double* n, double* out);                  /* rtutility.h:205-208 */
void oracle_refract(const double* v, const double* n, double n1, double n2, double* out); /* rtutility.h:210-227 */
void oracle_tonemap(const double* sum, int spp, double* out);                        /* rtutility.h:56-71 */
void oracle_hsl_roundtrip(const double* rgb, double* out);                           /* rtutility.h:81-165 */
void oracle_pile(const int* kind, const double* value, int n, double* out);          /* pile.h + main.c:169-181 */
void oracle_tracer_libc(const oracle_scene* sc, const double* o, const double* d, int nbRebondMax,
                        double AO_intensity, int useAO, double* out9);               /* main.c:118-242 */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
