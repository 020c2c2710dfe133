/* TEST INFRASTRUCTURE — CPU restatement of the reference hot path (see tipe_oracle.h).
 *
 * Written from the behaviour of /root/reference (file:line cited per function), not copied
 * from it: vectors are bare double[3], the refraction stack is an array, the RNG is pluggable.
 * What IS kept, deliberately, is the order of every floating-point operation, because the
 * libc-RNG mode has to be bit-identical to the reference binary.  Compile with
 * -ffp-contract=off (no FMA contraction) to keep that true on any -march.
 */
#define _GNU_SOURCE
#include "tipe_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * vec3 helpers — same operation order as vec3.h:57-139 (dot is ((x0*y0 + x1*y1) + x2*y2),
 * normalize divides each component by sqrt(dot), never multiplies by a reciprocal).
 * ---------------------------------------------------------------------------------------- */
typedef struct { double x[3]; } v3;

static inline v3 V(double a, double b, double c) { v3 r = {{a, b, c}}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x[0] + b.x[0], a.x[1] + b.x[1], a.x[2] + b.x[2]); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x[0] - b.x[0], a.x[1] - b.x[1], a.x[2] - b.x[2]); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x[0] * b.x[0], a.x[1] * b.x[1], a.x[2] * b.x[2]); }
static inline v3 vscale(v3 a, double t) { return V(a.x[0] * t, a.x[1] * t, a.x[2] * t); }
static inline v3 vdiv(v3 a, double t) { return V(a.x[0] / t, a.x[1] / t, a.x[2] / t); }
static inline v3 vneg(v3 a) { return V(-a.x[0], -a.x[1], -a.x[2]); }
static inline double vdot(v3 a, v3 b) { return a.x[0] * b.x[0] + a.x[1] * b.x[1] + a.x[2] * b.x[2]; }
static inline v3 vcross(v3 a, v3 b)
{
    return V(a.x[1] * b.x[2] - a.x[2] * b.x[1],
             a.x[2] * b.x[0] - a.x[0] * b.x[2],
             a.x[0] * b.x[1] - a.x[1] * b.x[0]);
}
static inline double vlen(v3 a) { return sqrt(vdot(a, a)); }
static inline v3 vnormalize(v3 a) { return vdiv(a, vlen(a)); }
static inline v3 vlerp(v3 a, v3 b, double t) { return vadd(a, vscale(vsub(b, a), t)); }   /* rtutility.h:32-34 */
static inline v3 from_trt(const trt_vec3* p) { return V(p->e[0], p->e[1], p->e[2]); }
static inline v3 from_ptr(const double* p) { return V(p[0], p[1], p[2]); }
static inline void to_ptr(v3 a, double* p) { p[0] = a.x[0]; p[1] = a.x[1]; p[2] = a.x[2]; }

/* ------------------------------------------------------------------------------------------
 * RNG.  LIBC: glibc rand() in call order (rtutility.h:192-193, 229-231).
 * PHILOX: Philox4x32-10 (Salmon et al., SC'11), key = (seed lo, seed hi),
 *         counter = (pixel_index, global sample index, slot, 0); uniform = (x >> 8) * 2^-24.
 *   slot 0        : x0 pixel jitter u, x1 pixel jitter v, x2 aperture dx, x3 aperture dy (main.c:265-269)
 *   slot 1 + 2*i  : bounce i — x0,x1 diffuse direction (main.c:163); x2,x3 AO direction (main.c:100)
 *                   refraction coin (main.c:185) = the low bytes of x0, x1, x2 of the same block (24 bits the uniforms above never read)
 *   slot 2 + 2*i  : unused (held the coin until round 2)
 * ---------------------------------------------------------------------------------------- */
API void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
    int mode;
    uint32_t key[2];
    uint32_t pixel, sample;
    uint32_t cached_slot; int cache_valid;
    uint32_t cache[4];
} rng_t;

enum { SLOT_CAMERA = 0 };
static inline uint32_t slot_bounce(int i) { return 1u + 2u * (uint32_t)i; }

static double rng_u01(rng_t* g, uint32_t slot, int lane)
{
    if (g->mode == TRT_ORACLE_RNG_LIBC) return rand() / (RAND_MAX + 1.0);
    if (!g->cache_valid || g->cached_slot != slot) {
        uint32_t ctr[4] = {g->pixel, g->sample, slot, 0u};
        oracle_philox4x32_10(ctr, g->key, g->cache);
        g->cached_slot = slot; g->cache_valid = 1;
    }
    return (double)(g->cache[lane] >> 8) * (1.0 / 16777216.0);
}

/* the refraction coin: one randomDouble(0,1) in libc mode; in Philox mode 24 bits taken from the low bytes of words 0-2 of the
 * bounce's block (rng_u01 uses the top 24 bits of a word, so these are independent of every direction drawn from the block) */
static double rng_coin(rng_t* g, uint32_t slot)
{
    if (g->mode == TRT_ORACLE_RNG_LIBC) return rand() / (RAND_MAX + 1.0);
    (void)rng_u01(g, slot, 0);                                   /* fills the cache for this slot */
    uint32_t bits = (g->cache[0] & 0xffu) | ((g->cache[1] & 0xffu) << 8) | ((g->cache[2] & 0xffu) << 16);
    return (double)bits * (1.0 / 16777216.0);
}

/* randomDouble(min,max) = min + (max-min)*u  (rtutility.h:229-231) */
static inline double rng_range(rng_t* g, uint32_t slot, int lane, double lo, double hi)
{
    return lo + (hi - lo) * rng_u01(g, slot, lane);
}

/* random_dir_no_norm (rtutility.h:189-203): uniform on the sphere via theta = 2*pi*u,
 * phi = acos(2v-1); the three products are FLOAT products of cosf/sinf results. */
static v3 rng_unit_vector(rng_t* g, uint32_t slot, int lane_u, int lane_v)
{
    const double PI_ = 3.1415926535897932385;
    double u = rng_u01(g, slot, lane_u);
    double v = rng_u01(g, slot, lane_v);
    double theta = 2 * PI_ * u;
    double phi = acos(2 * v - 1);
    v3 d;
    d.x[0] = cosf(theta) * sinf(phi);
    d.x[1] = sinf(theta) * sinf(phi);
    d.x[2] = cosf(phi);
    return vnormalize(d);
}

/* ------------------------------------------------------------------------------------------
 * Leaf geometry
 * ---------------------------------------------------------------------------------------- */
typedef struct { int hit; double dst; v3 point, normal; trt_material mat; } hit_t;

/* sphere.h:13-47 — near root if >= 1e-4, else far root if >= 1e-4; normal always outward. */
static hit_t sphere_hit(v3 center, double radius, v3 o, v3 d)
{
    hit_t h; h.hit = 0;
    v3 oc = vsub(o, center);
    double a = vdot(d, d);
    double b = 2.0 * vdot(oc, d);
    double c = vdot(oc, oc) - radius * radius;
    double disc = b * b - 4 * a * c;
    if (disc > 0) {
        double t1 = (-b - sqrt(disc)) / (2 * a);
        if (t1 >= 0.0001) {
            h.hit = 1; h.dst = t1;
            h.point = vadd(o, vscale(d, t1));
            h.normal = vnormalize(vsub(h.point, center));
            return h;
        }
        double t2 = (-b + sqrt(disc)) / (2 * a);
        if (t2 >= 0.0001) {
            h.hit = 1; h.dst = t2;
            h.point = vadd(o, vscale(d, t2));
            h.normal = vnormalize(vsub(h.point, center));
            return h;
        }
    }
    return h;
}

/* mesh.h:70-94 — Moller-Trumbore in the "dao" form, front faces only (det >= 1e-6),
 * dst/u/v/w all >= 1e-7; point and normal are computed unconditionally. */
static hit_t triangle_hit(const trt_triangle* t, v3 o, v3 d)
{
    v3 A = from_trt(&t->A);
    v3 ab = vsub(from_trt(&t->B), A);
    v3 ac = vsub(from_trt(&t->C), A);
    v3 n = vcross(ab, ac);
    v3 ao = vsub(o, A);
    v3 dao = vcross(ao, d);
    double det = -vdot(d, n);
    double inv = 1 / det;
    double dst = vdot(ao, n) * inv;
    double u = vdot(ac, dao) * inv;
    double v = -vdot(ab, dao) * inv;
    double w = 1 - u - v;
    hit_t h;
    h.hit = det >= 1E-6 && dst >= 0.0000001 && u >= 0.0000001 && v >= 0.0000001 && w >= 0.0000001;
    h.point = vadd(o, vscale(d, dst));
    h.normal = vnormalize(n);
    h.dst = dst;
    return h;
}

static const trt_mat_override k_builtin_overrides[3] = {
    /* texture.h:71-87 */
    {1, TRT_OV_EMISSION_COLOR | TRT_OV_EMISSION_STRENGTH | TRT_OV_ALPHA, {1, 1, 1}, 1.85, 0, 1.0, 0},
    {4, TRT_OV_ALPHA | TRT_OV_MATERIAL_INDEX | TRT_OV_REFLECTION, {0, 0, 0}, 0, 0.93, 0.6, 1.33},
    {3, TRT_OV_ALPHA | TRT_OV_MATERIAL_INDEX | TRT_OV_REFLECTION, {0, 0, 0}, 0, 0.3, 0.1, 1.50},
};

/* texture.h:16-27 + 44-90 — barycentrics by signed sub-areas, uv tiling by fmod, nearest texel,
 * slab selected by the triangle's material ordinal, then the per-ordinal overrides. */
static trt_material texel_material(const oracle_scene* sc, int tri_index, v3 P, v3 nrm)
{
    const trt_triangle* t = &sc->tris[tri_index];
    v3 A = from_trt(&t->A), B = from_trt(&t->B), C = from_trt(&t->C);
    double areaABC = vdot(nrm, vcross(vsub(B, A), vsub(C, A)));
    double areaPBC = vdot(nrm, vcross(vsub(B, P), vsub(C, P)));
    double areaPCA = vdot(nrm, vcross(vsub(C, P), vsub(A, P)));
    double b0 = areaPBC / areaABC;
    double b1 = areaPCA / areaABC;
    double b2 = 1.0 - b0 - b1;

    double uu = (b0 * t->uvA.u + b1 * t->uvB.u + b2 * t->uvC.u);
    double vv = (b0 * t->uvA.v + b1 * t->uvB.v + b2 * t->uvC.v);
    uu = fmod(uu, 1.0);
    vv = fmod(vv, 1.0);
    if (uu < 0) uu += 1.0;
    if (vv < 0) vv += 1.0;

    int x = (int)(uu * (double)(sc->tex_w));
    int y = (int)(vv * (double)(sc->tex_h));
    int ordinal = sc->quelMatPourTri[tri_index];
    int index = (y * sc->tex_w + x) + (sc->tex_h * sc->tex_w * ordinal);
    trt_material m = sc->mat_list[index];

    const trt_mat_override* ov = sc->overrides; int n_ov = sc->n_overrides;
    if (ov == NULL && n_ov < 0) { ov = k_builtin_overrides; n_ov = 3; }
    for (int k = 0; k < n_ov; k++) {
        if (ov[k].material_ordinal != ordinal) continue;
        if (ov[k].mask & TRT_OV_EMISSION_COLOR) memcpy(m.emissionColor.e, ov[k].emissionColor, 24);
        if (ov[k].mask & TRT_OV_EMISSION_STRENGTH) m.emissionStrength = ov[k].emissionStrength;
        if (ov[k].mask & TRT_OV_REFLECTION) m.reflectionStrength = ov[k].reflectionStrength;
        if (ov[k].mask & TRT_OV_ALPHA) m.alpha = ov[k].alpha;
        if (ov[k].mask & TRT_OV_MATERIAL_INDEX) m.materialIndex = ov[k].materialIndex;
    }
    return m;
}

/* sphere_uvmapping (texture.h:92-112): lat-long texel of the sky image under a hit point on sphere (c, r).
 * The reference indexes mat_list unchecked; theta = pi (straight down) or phi = 2*pi would read one row / one texel
 * past the intended one, so the index is clamped to the last texel here and on the GPU. */
static int sky_texel_index(v3 center, double radius, v3 P, int sky_w, int sky_h)
{
    const double PI_ = 3.1415926535897932385;              /* rtutility.h:12 */
    v3 d = vdiv(vsub(P, center), radius);
    double theta = acos(-d.x[1]);
    double phi = atan2(-d.x[2], d.x[0]) + PI_;
    double u = phi / (2 * PI_);
    double v = theta / PI_;
    int x = (int)(u * (double)sky_w);
    int y = (int)(v * (double)sky_h);
    long index = (long)y * sky_w + x;
    if (index < 0) index = 0;
    if (index > (long)sky_w * sky_h - 1) index = (long)sky_w * sky_h - 1;
    return (int)index;
}

typedef struct { uint64_t rays, sphere_tests, tri_tests; } counters_t;

/* main.c:52-92 — spheres first, then triangles, strict '<' so the first primitive wins ties.
 * The reference leaves mat/normal uninitialised on a miss; here they are zero. *id_out gets the
 * winning primitive (sphere k -> k, triangle k -> nbSpheres+k, miss -> -1). */
static hit_t nearest_hit(const oracle_scene* sc, v3 o, v3 d, counters_t* cnt, int* id_out)
{
    hit_t best; memset(&best, 0, sizeof best);
    best.hit = 0; best.dst = INFINITY;
    int id = -1;
    if (cnt) { cnt->rays++; cnt->sphere_tests += (uint64_t)sc->nbSpheres; cnt->tri_tests += (uint64_t)sc->nbTriangles; }
    for (int k = 0; k < sc->nbSpheres; k++) {
        hit_t h = sphere_hit(from_trt(&sc->spheres[k].center), sc->spheres[k].radius, o, d);
        if (h.hit && h.dst < best.dst) {
            best = h; best.mat = sc->spheres[k].mat; id = k;
            if (sc->sky_mat_list && k == sc->nbSpheres - 1) {          /* "cas du ciel", main.c:64-71 (commented out upstream) */
                int index = sky_texel_index(from_trt(&sc->spheres[k].center), sc->spheres[k].radius, h.point, sc->sky_w, sc->sky_h);
                best.mat.emissionColor = sc->sky_mat_list[index].diffuseColor;
                best.mat.alpha = 1.0;
            }
        }
    }
    for (int k = 0; k < sc->nbTriangles; k++) {
        hit_t h = triangle_hit(&sc->tris[k], o, d);
        if (h.hit && h.dst < best.dst) {
            trt_material m = texel_material(sc, k, h.point, h.normal);
            best = h; best.mat = m; id = sc->nbSpheres + k;
        }
    }
    if (id_out) *id_out = id;
    return best;
}

/* ------------------------------------------------------------------------------------------
 * Shading helpers
 * ---------------------------------------------------------------------------------------- */
static v3 reflect_dir(v3 v, v3 n) { return vsub(v, vscale(n, 2 * vdot(v, n))); }     /* rtutility.h:205-208 */

/* rtutility.h:210-227 — note both indices are squared before the ratio is taken, and the
 * refracted direction is not re-normalised. Total internal reflection falls back to the mirror. */
static v3 refract_dir(v3 v, v3 n, double n1, double n2)
{
    n1 *= n1;
    n2 *= n2;
    double radical = 1 - ((n1 / n2) * (n1 / n2)) * (1 - (vdot(n, v) * vdot(n, v)));
    if (radical > 0) {
        v3 tangential = vscale(vsub(v, vscale(n, vdot(v, n))), (n1 / n2));
        v3 along = vscale(vneg(n), sqrt(radical));
        return vadd(tangential, along);
    }
    return reflect_dir(v, n);
}

/* rtutility.h:81-165 — RGB -> HSL -> RGB; applied with unit gains at main.c:155-158. */
static v3 rgb_to_hsl(v3 c)
{
    double r = c.x[0], g = c.x[1], b = c.x[2];
    double mx = (r > g) ? ((r > b) ? r : b) : ((g > b) ? g : b);
    double mn = (r < g) ? ((r < b) ? r : b) : ((g < b) ? g : b);
    double h = 0.0, s, l = (mx + mn) / 2.0;
    if (mx == mn) { h = 0.0; s = 0.0; }
    else {
        double d = mx - mn;
        s = (l < 0.5) ? (d / (mx + mn)) : (d / (2.0 - mx - mn));
        if (mx == r) h = (g - b) / d + ((g < b) ? 6.0 : 0.0);
        else if (mx == g) h = (b - r) / d + 2.0;
        else if (mx == b) h = (r - g) / d + 4.0;
        h /= 6.0;
    }
    return V(h, s, l);
}
static double hue_channel(double t1, double t2, double hue)
{
    if (hue < 0.0) hue += 1.0;
    if (hue > 1.0) hue -= 1.0;
    if (6.0 * hue < 1.0) return t1 + (t2 - t1) * 6.0 * hue;
    if (2.0 * hue < 1.0) return t2;
    if (3.0 * hue < 2.0) return t1 + (t2 - t1) * ((2.0 / 3.0) - hue) * 6.0;
    return t1;
}
static v3 hsl_to_rgb(v3 hsl)
{
    double h = hsl.x[0], s = hsl.x[1], l = hsl.x[2];
    if (s == 0.0) return V(l, l, l);
    double t2 = (l < 0.5) ? (l * (1.0 + s)) : (l + s - l * s);
    double t1 = 2.0 * l - t2;
    return V(hue_channel(t1, t2, h + 1.0 / 3.0), hue_channel(t1, t2, h), hue_channel(t1, t2, h - 1.0 / 3.0));
}

/* Refraction stack, pile.h:23-72 restated on an array.  Element = (n[0], n[1]). */
typedef struct { double (*e)[2]; int top; } stack_t;   /* top = number of elements */
static void stack_push(stack_t* s, double a, double b) { s->e[s->top][0] = a; s->e[s->top][1] = b; s->top++; }
static void stack_pop(stack_t* s, double out[2])
{   /* depiler on an empty list returns an unset IndRef (pile.h:39-47); never happens here */
    if (s->top > 0) { s->top--; out[0] = s->e[s->top][0]; out[1] = s->e[s->top][1]; }
}
static void stack_enter(stack_t* s, double index)      /* index_suivant_pile, pile.h:61-66 */
{
    double old[2] = {0, 0};
    stack_pop(s, old);
    stack_push(s, old[0], old[1]);
    stack_push(s, old[1], index);
}
static void stack_peek(stack_t* s, double out[2])       /* info_pile_actuelle, pile.h:68-72 */
{
    stack_pop(s, out);
    stack_push(s, out[0], out[1]);
}

/* main.c:94-116 — one cosine-ish occlusion ray; (d/dst)^AO / 1 / AO if it hits anything, else 0. */
static v3 occlusion_term(const oracle_scene* sc, v3 P, v3 n, double AO, rng_t* g, int bounce, counters_t* cnt)
{
    v3 occ = V(0, 0, 0);
    v3 rd = rng_unit_vector(g, slot_bounce(bounce), 2, 3);
    v3 dir = vnormalize(vadd(n, rd));
    hit_t h = nearest_hit(sc, P, dir, cnt, NULL);
    if (h.hit) {
        double distance = vlen(vsub(h.point, P));
        double att = distance / h.dst;
        att = pow(att, AO);
        occ = vadd(occ, V(att, att, att));
    }
    return vdiv(vdiv(occ, 1), AO);
}

/* main.c:118-242 — one path.  out = (radiance, albedo AOV, normal AOV). */
static void trace_path(const oracle_scene* sc, v3 o, v3 d, int nbRebondMax, double AO, int useAO,
                       rng_t* g, counters_t* cnt, v3 out[3])
{
    v3 light = V(0, 0, 0), throughput = V(1, 1, 1), albedo = V(0, 0, 0), normal_aov = V(0, 0, 0);
    int is_alpha = 0, alpha_depth = 0;
    double n1, n2 = 1.0;
    double (*store)[2] = (double (*)[2])malloc(sizeof(double[2]) * (size_t)(2 * nbRebondMax + 4));
    stack_t st = {store, 0};
    stack_push(&st, 1.0, 1.0);

    for (int i = 0; i < nbRebondMax; i++) {
        hit_t h = nearest_hit(sc, o, d, cnt, NULL);
        trt_material m = h.mat;

        if (i == 0) { albedo = from_trt(&m.diffuseColor); normal_aov = h.normal; }
        if (i == alpha_depth && is_alpha) {
            albedo = from_trt(&m.diffuseColor);
            if (m.emissionStrength > 0) albedo = from_trt(&m.emissionColor);
            normal_aov = h.normal;
            is_alpha = 0;
        }
        if (!h.hit) break;

        if (i == alpha_depth && m.emissionStrength > 0) {       /* light seen directly: main.c:154-160 */
            v3 hsl = rgb_to_hsl(from_trt(&m.emissionColor));
            hsl.x[2] *= 1.0;
            hsl.x[1] *= 1.0;
            v3 c = hsl_to_rgb(hsl);
            out[0] = c; out[1] = c; out[2] = h.normal;
            free(store);
            return;
        }

        o = h.point;
        v3 diffuse_dir = vnormalize(vadd(h.normal, rng_unit_vector(g, slot_bounce(i), 0, 1)));
        v3 mirror_dir = reflect_dir(d, h.normal);
        v3 mixed = vlerp(diffuse_dir, mirror_dir, m.reflectionStrength);

        if (m.alpha <= 0.99 && m.alpha >= 0.0001) {              /* refraction: main.c:167-193 */
            v3 n = h.normal;
            double top[2] = {0, 0};
            stack_enter(&st, m.materialIndex);
            stack_peek(&st, top);
            n1 = top[0]; n2 = top[1];
            if (vdot(d, h.normal) > 0) {                         /* leaving the object */
                n = vneg(h.normal);
                stack_pop(&st, top);
                n1 = top[1]; n2 = top[0];
            }
            v3 refracted = refract_dir(d, n, n1, n2);
            double coin = rng_coin(g, slot_bounce(i));
            if (coin > m.alpha) { d = refracted; continue; }
            else d = mixed;
        }
        if (m.alpha > 0.99) { is_alpha = 0; d = mixed; }
        if (m.alpha < 0.0001) { o = h.point; is_alpha = 1; alpha_depth++; continue; }   /* texture hole */

        v3 ecol = from_trt(&m.emissionColor), dcol = from_trt(&m.diffuseColor);
        if (useAO) {
            v3 emitted = vscale(ecol, m.emissionStrength * 1.5 * AO);
            light = vadd(light, vmul(emitted, throughput));
            if (throughput.x[0] > 0.5 || throughput.x[1] > 0.5 || throughput.x[2] > 0.5)
                throughput = vmul(dcol, vscale(throughput, 1.3));
            throughput = vmul(dcol, throughput);
            v3 occ = occlusion_term(sc, h.point, h.normal, AO, g, i, cnt);
            throughput = vmul(throughput, occ);
        } else {
            v3 emitted = vscale(ecol, m.emissionStrength);
            light = vadd(light, vmul(emitted, throughput));
            if (throughput.x[0] > 0.5 || throughput.x[1] > 0.5 || throughput.x[2] > 0.5)
                throughput = vmul(dcol, vscale(throughput, 1.3));
            throughput = vmul(dcol, throughput);
        }
    }
    free(store);
    out[0] = light; out[1] = albedo; out[2] = normal_aov;
}

/* camera.h:42-54 — the aperture offset is applied along WORLD x/y, not in the lens plane. */
static void camera_ray(double u, double v, const trt_camera* cam, double focus, double dx, double dy, v3* o, v3* d)
{
    v3 org = from_trt(&cam->origin);
    v3 direction = vadd(from_trt(&cam->coin_bas_gauche),
                        vadd(vscale(from_trt(&cam->horizontal), u), vsub(vscale(from_trt(&cam->vertical), v), org)));
    v3 destination = vadd(org, vscale(direction, focus));
    v3 new_origin = vadd(org, V(dx, dy, 0));
    *o = new_origin;
    *d = vnormalize(vsub(destination, new_origin));
}

/* rtutility.h:56-71 — sqrtf (FLOAT) of the mean, clamp to .999, scale by 256, truncate. */
static double clamp01(double x, double lo, double hi) { if (x < lo) return lo; if (x > hi) return hi; return x; }
static v3 tonemap(v3 sum, int spp)
{
    double r = sum.x[0], g = sum.x[1], b = sum.x[2];
    double rapport = 1.0 / spp;
    r = sqrtf(rapport * r);
    g = sqrtf(rapport * g);
    b = sqrtf(rapport * b);
    return V((int)(256 * clamp01(r, 0.0, 0.999)), (int)(256 * clamp01(g, 0.0, 0.999)), (int)(256 * clamp01(b, 0.0, 0.999)));
}

/* ------------------------------------------------------------------------------------------
 * Image driver — main.c:245-284 (rows top to bottom, columns left to right, spp inner).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const oracle_scene* sc; const trt_params* p; int rng_mode;
    int row_first, row_step;         /* this worker renders rows H-1-row_first, then -row_step, ... */
    trt_vec3 *canva, *albedo, *normal; double* sums9;
    counters_t cnt;
} worker_t;

static void* render_rows(void* arg)
{
    worker_t* w = (worker_t*)arg;
    const trt_params* p = w->p;
    const int W = p->largeur_image, H = p->hauteur_image, spp = p->nbRayonParPixel;
    const int divisor = p->sample_total > 0 ? p->sample_total : spp;
    rng_t g; memset(&g, 0, sizeof g);
    g.mode = w->rng_mode;
    g.key[0] = (uint32_t)(p->seed & 0xffffffffu); g.key[1] = (uint32_t)(p->seed >> 32);

    for (int j = H - 1 - w->row_first; j >= 0; j -= w->row_step) {
        for (int i = 0; i < W; i++) {
            int pixel = j * W + i;
            v3 total[3] = {V(0, 0, 0), V(0, 0, 0), V(0, 0, 0)};
            for (int s = 0; s < spp; s++) {
                g.pixel = (uint32_t)pixel; g.sample = (uint32_t)(p->sample_offset + s); g.cache_valid = 0;
                double u = ((double)i + rng_range(&g, SLOT_CAMERA, 0, -0.5, 0.5)) / (W - 1);
                double v = ((double)j + rng_range(&g, SLOT_CAMERA, 1, -0.5, 0.5)) / (H - 1);
                double dx = rng_range(&g, SLOT_CAMERA, 2, -0.5, 0.5) * p->ouverture_x;
                double dy = rng_range(&g, SLOT_CAMERA, 3, -0.5, 0.5) * p->ouverture_y;
                v3 o, d, res[3];
                camera_ray(u, v, &p->cam, p->focus_distance, dx, dy, &o, &d);
                trace_path(w->sc, o, d, p->nbRebondMax, p->AO_intensity, p->useAO, &g, &w->cnt, res);
                total[0] = vadd(total[0], res[0]);
                total[1] = vadd(total[1], res[1]);
                total[2] = vadd(total[2], res[2]);
            }
            if (w->canva) to_ptr(tonemap(total[0], divisor), w->canva[pixel].e);
            if (w->albedo) to_ptr(vdiv(total[1], divisor), w->albedo[pixel].e);
            if (w->normal) to_ptr(vdiv(total[2], divisor), w->normal[pixel].e);
            if (w->sums9) for (int c = 0; c < 3; c++) to_ptr(total[c], w->sums9 + (size_t)pixel * 9 + 3 * c);
        }
    }
    return NULL;
}

API int oracle_render(const oracle_scene* sc, const trt_params* p, int rng_mode, int threads,
                      trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab, double* sums9,
                      oracle_stats* stats)
{
    if (!sc || !p || p->largeur_image < 2 || p->hauteur_image < 2 || p->nbRayonParPixel < 1) return -1;
    if (rng_mode == TRT_ORACLE_RNG_LIBC || threads < 1) threads = 1;   /* rand() order is only defined serially */
    if (threads > p->hauteur_image) threads = p->hauteur_image;
    if (threads > 256) threads = 256;

    struct timeval t0, t1;
    gettimeofday(&t0, NULL);
    worker_t* w = (worker_t*)calloc((size_t)threads, sizeof(worker_t));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    for (int k = 0; k < threads; k++) {
        w[k].sc = sc; w[k].p = p; w[k].rng_mode = rng_mode;
        w[k].row_first = k; w[k].row_step = threads;
        w[k].canva = canva; w[k].albedo = albedo_tab; w[k].normal = normal_tab; w[k].sums9 = sums9;
    }
    if (threads == 1) render_rows(&w[0]);
    else {
        for (int k = 0; k < threads; k++) pthread_create(&th[k], NULL, render_rows, &w[k]);
        for (int k = 0; k < threads; k++) pthread_join(th[k], NULL);
    }
    gettimeofday(&t1, NULL);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->paths = (uint64_t)p->largeur_image * (uint64_t)p->hauteur_image * (uint64_t)p->nbRayonParPixel;
        for (int k = 0; k < threads; k++) {
            stats->rays += w[k].cnt.rays; stats->sphere_tests += w[k].cnt.sphere_tests; stats->tri_tests += w[k].cnt.tri_tests;
        }
        stats->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-6 * (double)(t1.tv_usec - t0.tv_usec);
    }
    free(w); free(th);
    return 0;
}

API int oracle_primary(const oracle_scene* sc, const trt_params* p,
                       double* t, int* id, trt_vec3* normal, trt_vec3* albedo)
{
    if (!sc || !p) return -1;
    const int W = p->largeur_image, H = p->hauteur_image;
    for (int j = 0; j < H; j++) for (int i = 0; i < W; i++) {
        double u = (double)i / (W - 1), v = (double)j / (H - 1);
        v3 o, d; int prim;
        camera_ray(u, v, &p->cam, p->focus_distance, 0.0, 0.0, &o, &d);
        hit_t h = nearest_hit(sc, o, d, NULL, &prim);
        int px = j * W + i;
        if (t) t[px] = h.dst;
        if (id) id[px] = prim;
        if (normal) to_ptr(h.normal, normal[px].e);
        if (albedo) memcpy(albedo[px].e, h.mat.diffuseColor.e, 24);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Leaf entry points for the known-answer tests
 * ---------------------------------------------------------------------------------------- */
API void oracle_init_camera(const double* origin, const double* target, const double* up,
                            double vfov, double ratio, trt_camera* out)
{   /* camera.h:17-40 */
    double theta = vfov * 3.1415926535897932385 / 180.0;
    double h = tan(theta / 2);
    double viewport_h = 2.0 * h;
    double viewport_w = ratio * viewport_h;
    v3 org = from_ptr(origin);
    v3 w = vnormalize(vsub(org, from_ptr(target)));
    v3 u = vnormalize(vcross(from_ptr(up), w));
    v3 v = vcross(w, u);
    v3 hor = vscale(u, viewport_w), ver = vscale(v, viewport_h);
    v3 corner = vsub(org, vadd(vdiv(hor, 2), vadd(vdiv(ver, 2), w)));
    to_ptr(org, out->origin.e); to_ptr(hor, out->horizontal.e);
    to_ptr(ver, out->vertical.e); to_ptr(corner, out->coin_bas_gauche.e);
}

API void oracle_get_ray(double u, double v, const trt_camera* cam, double focus, double dx, double dy,
                        double* origin, double* dir)
{ v3 o, d; camera_ray(u, v, cam, focus, dx, dy, &o, &d); to_ptr(o, origin); to_ptr(d, dir); }

API int oracle_hit_sphere(const double* center, double radius, const double* o, const double* d,
                          double* dst, double* point, double* normal)
{
    hit_t h = sphere_hit(from_ptr(center), radius, from_ptr(o), from_ptr(d));
    if (h.hit) { *dst = h.dst; to_ptr(h.point, point); to_ptr(h.normal, normal); }
    return h.hit;
}

API int oracle_hit_triangle(const trt_triangle* tri, const double* o, const double* d,
                            double* dst, double* point, double* normal)
{
    hit_t h = triangle_hit(tri, from_ptr(o), from_ptr(d));
    *dst = h.dst; to_ptr(h.point, point); to_ptr(h.normal, normal);
    return h.hit;
}

API void oracle_tri_uvmapping(const oracle_scene* sc, int tri_index, const double* point, const double* normal,
                              trt_material* out)
{ *out = texel_material(sc, tri_index, from_ptr(point), from_ptr(normal)); }

API int oracle_sphere_uvmapping(const double* center, double radius, const double* point, int sky_w, int sky_h)
{ return sky_texel_index(from_ptr(center), radius, from_ptr(point), sky_w, sky_h); }

API void oracle_reflect(const double* v, const double* n, double* out)
{ to_ptr(reflect_dir(from_ptr(v), from_ptr(n)), out); }

API void oracle_refract(const double* v, const double* n, double n1, double n2, double* out)
{ to_ptr(refract_dir(from_ptr(v), from_ptr(n), n1, n2), out); }

API void oracle_tonemap(const double* sum, int spp, double* out)
{ to_ptr(tonemap(from_ptr(sum), spp), out); }

API void oracle_hsl_roundtrip(const double* rgb, double* out)
{ to_ptr(hsl_to_rgb(rgb_to_hsl(from_ptr(rgb))), out); }

/* Same protocol as ref_kat_pile in ref_wrapper.c. */
API void oracle_pile(const int* kind, const double* value, int n, double* out)
{
    double (*store)[2] = (double (*)[2])malloc(sizeof(double[2]) * (size_t)(2 * n + 4));
    stack_t st = {store, 0};
    stack_push(&st, 1.0, 1.0);
    for (int k = 0; k < n; k++) {
        double top[2] = {0, 0};
        stack_enter(&st, value[k]);
        stack_peek(&st, top);
        double n1 = top[0], n2 = top[1];
        if (kind[k] == 1) { stack_pop(&st, top); n1 = top[1]; n2 = top[0]; }
        out[2 * k] = n1; out[2 * k + 1] = n2;
    }
    free(store);
}

API void oracle_tracer_libc(const oracle_scene* sc, const double* o, const double* d, int nbRebondMax,
                            double AO_intensity, int useAO, double* out9)
{
    rng_t g; memset(&g, 0, sizeof g); g.mode = TRT_ORACLE_RNG_LIBC;
    v3 res[3];
    trace_path(sc, from_ptr(o), from_ptr(d), nbRebondMax, AO_intensity, useAO, &g, NULL, res);
    for (int c = 0; c < 3; c++) to_ptr(res[c], out9 + 3 * c);
}
