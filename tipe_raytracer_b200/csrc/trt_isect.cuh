// trt_isect.cuh — FP32 intersection routines of the path-tracing kernel.
//
// Semantics follow the reference's CPU path (cited per function); the arithmetic is re-derived
// for FP32 on the GPU:
//   * rays restart exactly on the surface they left (main.c:162).  In FP64 the reference gets away
//     with absolute thresholds (1e-4 spheres, 1e-7 triangles); in FP32 the primitive the ray
//     starts on is identified by id instead: a source sphere is solved with c = 0 exactly, a source
//     triangle (and its coincident opposite-facing twin in double-faced meshes) is skipped —
//     both are what the reference's thresholds amount to in exact arithmetic;
//   * the quadratic uses the cancellation-free root pair (q/a, c/q);
//   * un-normalised directions are legal (main.c:165 lerps without renormalising): distances are
//     in units of |d| everywhere, exactly like the reference.
#pragma once

#include "trt_device.cuh"

#define TRT_NO_PRIM (-1)

struct Ray {
    float3 o, d;
};

struct TraceCounters {
    unsigned box_tests, tri_tests, sphere_tests;
};

// sphere.h:13-47.  Returns the accepted root (near if >= 1e-4, else far if >= 1e-4) or -1.
// `on_surface`: the ray starts on this very sphere => c = |oc|^2 - r^2 is 0 by construction.
__device__ __forceinline__ float sphere_root(float4 cr, float3 o, float3 d, float a, float inv_a, bool on_surface)
{
    const float3 oc = f3(o.x - cr.x, o.y - cr.y, o.z - cr.z);
    const float hb = dot3(oc, d);                                   // b / 2
    float c = fmaf(oc.x, oc.x, fmaf(oc.y, oc.y, fmaf(oc.z, oc.z, -cr.w * cr.w)));
    if (on_surface) c = 0.0f;
    const float qd = fmaf(hb, hb, -a * c);                          // discriminant / 4
    if (!(qd > 0.0f)) return -1.0f;                                 // sphere.h:27 (strict)
    const float sq = qd * rsq_approx(fmaxf(qd, 1e-36f));      // sqrt(qd), 2 ulp: MUFU.RSQ + FMUL (qd > 0 here; the floor keeps a denormal qd from flushing to 1/0)                               // sqrt(qd), 2 ulp: MUFU.RSQ + FMUL
    const float q = -(hb + copysignf(sq, hb));                      // never cancels
    const float ra = q * inv_a;
    const float rb = (q != 0.0f) ? c * rcp_approx(q) : 0.0f;         // MUFU.RCP + FMUL (__fdividef adds a 5-instruction range rescue for |q| > 2^126)
    const float t1 = fminf(ra, rb), t2 = fmaxf(ra, rb);
    if (t1 >= 1e-4f) return t1;                                     // sphere.h:29
    if (t2 >= 1e-4f) return t2;                                     // sphere.h:38
    return -1.0f;
}

// mesh.h:70-94, with A, AB, AC, N = ABxAC precomputed.  Front faces only (det >= 1e-6),
// dst >= tmin (FP32 stand-in for 1e-7), u, v, w >= 1e-7.
__device__ __forceinline__ bool triangle_test(float4 r0, float4 r1, float4 r2, float3 o, float3 d, float tmin,
                                              float& t, float& u, float& v)
{
    const float3 A = f3(r0.x, r0.y, r0.z), AB = f3(r0.w, r1.x, r1.y), AC = f3(r1.z, r1.w, r2.x), N = f3(r2.y, r2.z, r2.w);
    const float det = -dot3(d, N);
    const float3 ao = o - A;
    const float3 dao = cross3(ao, d);
    const float inv = rcp_approx(det);              // MUFU.RCP, 1 ulp: the accepted-hit thresholds below are 1e-6 / 1e-7, far coarser
    t = dot3(ao, N) * inv;
    u = dot3(AC, dao) * inv;
    v = -dot3(AB, dao) * inv;
    const float w = 1.0f - u - v;
    return det >= 1e-6f && t >= tmin && u >= 1e-7f && v >= 1e-7f && w >= 1e-7f;
}

// ---- resumable closest-hit query ------------------------------------------------------------------
// closest_hit (main.c:52-92): spheres in order, then triangles; strict '<' keeps the first primitive
// on exact ties (for triangles: the lower index in the caller's array).
// The query is split so that a warp can leave the BVH loop while some lanes are still inside it:
//   ray_begin        : the sphere scan (uniform trip count, no divergence) + traversal set-up
//   bvh_step_unified : one node visit or one leaf test; state lives in Trav + the lane's local stack
#define TRT_TRAV_DONE 0x7fffffff

struct Trav {
    int   node;     // TRT_TRAV_DONE when the query is complete; >= 0 internal node; < 0 leaf ~p
    int   sp;       // stack pointer (stack[0] holds TRT_TRAV_DONE)
    float t;        // best distance so far, in units of |d|
    int   prim;     // TRT_NO_PRIM, sphere k -> k, triangle at leaf position p -> n_spheres + p
    float u, v;     // Moller-Trumbore barycentrics of the best triangle (weights of B and C)
};

__device__ __forceinline__ float3 safe_inverse(float3 d)
{
    // MUFU.RCP (1 ulp): the slab tests only have to be conservative, and the boxes are padded by 2^-21 of the scene size
    const float e = 1e-30f;
    return f3(rcp_approx(fabsf(d.x) > e ? d.x : copysignf(e, d.x)),
              rcp_approx(fabsf(d.y) > e ? d.y : copysignf(e, d.y)),
              rcp_approx(fabsf(d.z) > e ? d.z : copysignf(e, d.z)));
}

// any_hit: the caller only needs "hit anything" (ambient_occlusion, main.c:104-106).
// ray_scan = the part of a closest_hit query that needs no traversal state: the sphere scan, and whether the BVH has to be
// walked at all (returns true) — false when there are no triangles, when an any-hit query is already answered by a sphere,
// or when the ray misses the bounds of all triangles (or reaches them beyond the best sphere hit).
template <bool HAS_TRIS, bool BRUTE, bool COUNT>
__device__ __forceinline__ bool ray_scan(const DevScene& sc, const float4* __restrict__ s_sph, float3 o, float3 d,
                                         int src_prim, int src_twin, bool any_hit, Trav& tr, TraceCounters& cnt)
{
    tr.t = __int_as_float(0x7f800000);
    tr.prim = TRT_NO_PRIM;
    tr.u = tr.v = 0.0f;
    tr.node = TRT_TRAV_DONE;
    tr.sp = 0;
    const float a = dot3(d, d);
    const float inv_a = __frcp_rn(a);
    const int ns = sc.n_spheres;
    for (int k = 0; k < ns; k++) {
        const float t = sphere_root(s_sph[k], o, d, a, inv_a, k == src_prim);
        if (t >= 0.0f && t < tr.t) { tr.t = t; tr.prim = k; }
    }
    if (COUNT) cnt.sphere_tests += ns;
    if (!HAS_TRIS || sc.n_tris == 0) return false;
    if (any_hit && tr.prim != TRT_NO_PRIM) return false;
    if (BRUTE) {
        // linear scan like main.c:81-90 (leaf order; ties resolved by the caller's index)
        const int src_leaf = src_prim - ns;
        for (int p = 0; p < sc.n_tris; p++) {
            if (p == src_leaf || p == src_twin) continue;
            float t, u, v;
            const float4 r0 = __ldg(sc.tri + 4 * p), r1 = __ldg(sc.tri + 4 * p + 1), r2 = __ldg(sc.tri + 4 * p + 2);
            if (COUNT) cnt.tri_tests++;
            if (triangle_test(r0, r1, r2, o, d, sc.tri_tmin, t, u, v)) {
                bool take = t < tr.t;
                if (t == tr.t && tr.prim >= ns) take = __ldg(sc.tri_orig + p) < __ldg(sc.tri_orig + (tr.prim - ns));
                if (take) { tr.t = t; tr.prim = ns + p; tr.u = u; tr.v = v; }
            }
        }
        return false;
    }
    // scene-box cull: a ray that misses the bounds of all triangles (or reaches them beyond the best
    // sphere hit) never enters the traversal loop
    const float3 idir = safe_inverse(d);
    const float ax = (sc.root_lo.x - o.x) * idir.x, bx = (sc.root_hi.x - o.x) * idir.x;
    const float ay = (sc.root_lo.y - o.y) * idir.y, by = (sc.root_hi.y - o.y) * idir.y;
    const float az = (sc.root_lo.z - o.z) * idir.z, bz = (sc.root_hi.z - o.z) * idir.z;
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), sc.box_tmin));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tr.t));
    if (COUNT) cnt.box_tests += 1;
    return tf >= tn;
}

template <bool HAS_TRIS, bool BRUTE, bool COUNT>
__device__ __forceinline__ void ray_begin(const DevScene& sc, const float4* __restrict__ s_sph, float3 o, float3 d,
                                          int src_prim, int src_twin, bool any_hit, Trav& tr, int* stack, TraceCounters& cnt)
{
    if (ray_scan<HAS_TRIS, BRUTE, COUNT>(sc, s_sph, o, d, src_prim, src_twin, any_hit, tr, cnt)) {
        stack[0] = TRT_TRAV_DONE;
        tr.sp = 1;
        tr.node = sc.root;
    }
}

// Sphere-only scenes with ambient occlusion: the AO ray of bounce i and the bounce ray of bounce i+1 leave the same
// point (main.c:162 / main.c:102), so both are resolved in ONE scan that shares oc and c = |oc|^2 - r^2 per sphere.
// The AO query only needs "any hit" (main.c:106): with t2 = (-hb + sqrt(qd))/a the larger root,
//   qd > 0 and (t1 >= 1e-4 or t2 >= 1e-4)  <=>  qd > 0 and sqrt(qd) >= 1e-4*a + hb,   no square root needed.
template <bool COUNT>
__device__ __forceinline__ void spheres_pair(const DevScene& sc, const float4* __restrict__ s_sph, float3 o, float3 d, float3 d_ao,
                                             bool do_bounce, bool do_ao, int src_prim, Trav& tr, bool& occluded, TraceCounters& cnt)
{
    tr.t = __int_as_float(0x7f800000);
    tr.prim = TRT_NO_PRIM;
    tr.u = tr.v = 0.0f;
    tr.node = TRT_TRAV_DONE;
    tr.sp = 0;
    occluded = false;
    const float a = dot3(d, d), inv_a = __frcp_rn(a);
    const float a2 = dot3(d_ao, d_ao), eps2 = 1e-4f * a2;
    const int ns = sc.n_spheres;
    // unrolled by 5 (same-call A/B: C1 11.45 -> 11.07 ms, C5 -3.6 %; by 2: no change, by 10: 11.26)
#pragma unroll 5
    for (int k = 0; k < ns; k++) {
        const float4 cr = s_sph[k];
        const float3 oc = f3(o.x - cr.x, o.y - cr.y, o.z - cr.z);
        float c = fmaf(oc.x, oc.x, fmaf(oc.y, oc.y, fmaf(oc.z, oc.z, -cr.w * cr.w)));
        if (k == src_prim) c = 0.0f;
        // both queries are evaluated for every lane and masked by do_bounce / do_ao (a lane usually has both; branching
        // around each costs a reconvergence barrier per sphere); only the root computation stays behind a branch
        {
            const float hb = dot3(oc, d);
            const float qd = fmaf(hb, hb, -a * c);
            if (qd > 0.0f && do_bounce) {
                const float sq = qd * rsq_approx(fmaxf(qd, 1e-36f));      // sqrt(qd), 2 ulp: MUFU.RSQ + FMUL (qd > 0 here; the floor keeps a denormal qd from flushing to 1/0)
                const float q = -(hb + copysignf(sq, hb));
                const float ra = q * inv_a;
                const float rb = (q != 0.0f) ? c * rcp_approx(q) : 0.0f;         // MUFU.RCP + FMUL (__fdividef adds a 5-instruction range rescue for |q| > 2^126)
                const float t1 = fminf(ra, rb), t2 = fmaxf(ra, rb);
                const float t = (t1 >= 1e-4f) ? t1 : ((t2 >= 1e-4f) ? t2 : -1.0f);
                if (t >= 0.0f && t < tr.t) { tr.t = t; tr.prim = k; }
            }
        }
        {
            const float hb = dot3(oc, d_ao);
            const float qd = fmaf(hb, hb, -a2 * c);
            const float e = eps2 + hb;
            occluded = occluded | (do_ao & (qd > 0.0f) & ((e <= 0.0f) | (qd >= e * e)));
        }
    }
    if (COUNT) cnt.sphere_tests += ns * ((do_bounce ? 1 : 0) + (do_ao ? 1 : 0));
}

// compare-exchange of (key, ref) pairs: the 5-comparator network below sorts the four children by entry distance
__device__ __forceinline__ void order2(float& ka, int& ra, float& kb, int& rb)
{
    const bool sw = kb < ka;
    const float k0 = sw ? kb : ka, k1 = sw ? ka : kb;
    const int r0 = sw ? rb : ra, r1 = sw ? ra : rb;
    ka = k0; kb = k1; ra = r0; rb = r1;
}

// One traversal step with a UNIFIED FETCH: whatever the lane is at — a 4-wide node (128 B) or a leaf's triangle
// (64 B) — its record is requested first, with 256-bit loads, and only then do the two code paths (slab tests /
// Moller-Trumbore) run.  In the plain if-if loop the leaf path's loads were issued after the node path had
// finished: two dependent L1/L2 round trips per trip, and 40 % of all stall samples sat on the triangle fetch.
template <bool COUNT>
__device__ __forceinline__ void bvh_step_unified(const DevScene& sc, float3 o, float3 d, float3 idir, float3 ood,
                                                 int src_leaf, int twin_leaf, bool any_hit, Trav& tr, int* stack, TraceCounters& cnt)
{
    const bool leaf = tr.node < 0;
    const int idx = leaf ? (~tr.node & TRT_LEAF_POS_MASK) : tr.node;              // leaf reference: position | (run length - 1) << 27
    TRT_CHECK(leaf ? (idx + ((~tr.node) >> TRT_LEAF_POS_BITS) < sc.n_tris) : (idx < sc.n_nodes), 104);
    TRT_CHECK(tr.sp >= 1 && tr.sp + 3 <= TRT_STACK_DEPTH, 105);
    const float4* rec = leaf ? sc.tri + 4 * (size_t)idx : sc.nodes + 8 * (size_t)idx;
    const int rot = leaf ? 0 : (idx & 3);
    float4 f0a, f0b, f1a, f1b, f2a = make_float4(0, 0, 0, 0), f2b = f2a, f3a = f2a, f3b = f2a;
    ldg256(rec + 2 * ((0 + rot) & 3), f0a, f0b);
    ldg256(rec + 2 * ((1 + rot) & 3), f1a, f1b);
    if (!leaf) {
        ldg256(rec + 2 * ((2 + rot) & 3), f2a, f2b);
        ldg256(rec + 2 * ((3 + rot) & 3), f3a, f3b);
        const float INF = __int_as_float(0x7f800000);
#define TRT_SLAB(LX, HX, LY, HY, LZ, HZ, REF, KEY)                                                        \
    {                                                                                                      \
        const float ax = fmaf(LX, idir.x, -ood.x), bx = fmaf(HX, idir.x, -ood.x);                          \
        const float ay = fmaf(LY, idir.y, -ood.y), by = fmaf(HY, idir.y, -ood.y);                          \
        const float az = fmaf(LZ, idir.z, -ood.z), bz = fmaf(HZ, idir.z, -ood.z);                          \
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), sc.box_tmin));           \
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tr.t));           \
        KEY = (tf >= tn && REF != TRT_CHILD_EMPTY) ? tn : INF;                                             \
        if (COUNT && REF != TRT_CHILD_EMPTY) cnt.box_tests++;                                               \
    }
        float k0, k1, k2, k3;
        int r0 = __float_as_int(f3a.x), r1 = __float_as_int(f3a.y), r2 = __float_as_int(f3a.z), r3 = __float_as_int(f3a.w);
        TRT_SLAB(f0a.x, f0b.x, f1a.x, f1b.x, f2a.x, f2b.x, r0, k0)
        TRT_SLAB(f0a.y, f0b.y, f1a.y, f1b.y, f2a.y, f2b.y, r1, k1)
        TRT_SLAB(f0a.z, f0b.z, f1a.z, f1b.z, f2a.z, f2b.z, r2, k2)
        TRT_SLAB(f0a.w, f0b.w, f1a.w, f1b.w, f2a.w, f2b.w, r3, k3)
#undef TRT_SLAB
        order2(k0, r0, k1, r1); order2(k2, r2, k3, r3);
        order2(k0, r0, k2, r2); order2(k1, r1, k3, r3);
        order2(k1, r1, k2, r2);
        int sp = tr.sp;
        if (k3 < INF) stack[sp++] = r3;
        if (k2 < INF) stack[sp++] = r2;
        if (k1 < INF) stack[sp++] = r1;
        if (k0 < INF) { tr.node = r0; tr.sp = sp; }
        else { tr.node = stack[sp - 1]; tr.sp = sp - 1; }
    } else {
        // a leaf is a RUN of triangles that share one box (bvh_build.cu, group_heads): 64-byte records idx, idx+1, ...;
        // the head record carries the count.  The first record came with the unified fetch above.
        int next = stack[--tr.sp];
        int left = ((~tr.node) >> TRT_LEAF_POS_BITS) + 1;
        int p = idx;
        for (;;) {
            if (p != src_leaf && p != twin_leaf) {
                float t, u, v;
                if (COUNT) cnt.tri_tests++;
                if (triangle_test(f0a, f0b, f1a, o, d, sc.tri_tmin, t, u, v)) {
                    const int ns = sc.n_spheres;
                    bool take = t < tr.t;
                    if (t == tr.t && tr.prim >= ns) take = __float_as_int(f1b.x) < __ldg(sc.tri_orig + (tr.prim - ns));
                    if (take) {
                        tr.t = t; tr.prim = ns + p; tr.u = u; tr.v = v;
                        if (any_hit) { next = TRT_TRAV_DONE; left = 0; }
                    }
                }
            }
            if (--left <= 0) break;
            p++; rec += 4;
            ldg256(rec, f0a, f0b); ldg256(rec + 2, f1a, f1b);
        }
        tr.node = next;
    }
}

// One traversal trip with DEFERRED leaves: a lane that reaches a leaf parks it in `pend` and goes on with its next stack
// entry; the parked leaves of the whole warp are tested together (see the trigger below).  The triangle loop then runs with
// several lanes instead of the 3-4 that happen to be at a leaf in any one trip (ncu: a third of all warp instructions at
// 3.5 lanes).  Nodes visited in between do not see the parked leaf's hit distance yet, so a few more of them are visited;
// the result is unchanged.  Must be called by all 32 lanes.
// TOP: the first top_k wide nodes (the top levels of the tree: the collapse numbers nodes level by level) were staged into
// shared memory at kernel start (TMA bulk copy, pathtrace.cu) and are read from there, same rotated layout.
// N64: 64-byte nodes with 8-bit quantised child boxes (bvh_build.cu, compress_nodes) instead of the 128-byte ones: two 256-bit
// loads per visit instead of four, and a byte -> float conversion folded into each slab FMA (experiment, TRT_TUNE_NODE64).
template <bool COUNT, bool TOP = false, bool N64 = false>
__device__ __forceinline__ void bvh_trip_deferred(const DevScene& sc, float3 o, float3 d, float3 idir, float3 ood,
                                                  int src_leaf, int twin_leaf, bool any_hit, int share,
                                                  Trav& tr, int& pend, int* stack, TraceCounters& cnt, int exp_flags = 0,
                                                  const float4* s_top = nullptr, int top_k = 0)
{
    if (tr.node >= 0 && tr.node != TRT_TRAV_DONE) {
        TRT_CHECK(tr.node < sc.n_nodes, 101);
        TRT_CHECK(tr.sp >= 1 && tr.sp + 3 <= TRT_STACK_DEPTH, 102);
        const float INF = __int_as_float(0x7f800000);
        float k0, k1, k2, k3;
        int r0, r1, r2, r3;
        if (N64) {
            const float4* rec = sc.nodes64 + 4 * (size_t)tr.node;
            float4 w0, w1, w2, w3;
            ldg256(rec, w0, w1); ldg256(rec + 2, w2, w3);
            const unsigned ex = __float_as_uint(w0.w);
            // plane = origin + step * 2^(e-127); t = (plane - o) * idir = step * A + B.  float(step) is formed as the bits 0x4B0000ss = 2^23 + ss
            // (one PRMT), and the 2^23 is taken out again in the constant term (the half step this can cost is in the boxes' outward rounding)
            const float Ax = __uint_as_float((ex & 0xffu) << 23) * idir.x, Ay = __uint_as_float(((ex >> 8) & 0xffu) << 23) * idir.y,
                        Az = __uint_as_float(((ex >> 16) & 0xffu) << 23) * idir.z;
            const float Bx = fmaf(-8388608.0f, Ax, fmaf(w0.x, idir.x, -ood.x)), By = fmaf(-8388608.0f, Ay, fmaf(w0.y, idir.y, -ood.y)),
                        Bz = fmaf(-8388608.0f, Az, fmaf(w0.z, idir.z, -ood.z));
            const unsigned lx = __float_as_uint(w1.x), hx = __float_as_uint(w1.y), ly = __float_as_uint(w1.z), hy = __float_as_uint(w1.w),
                           lz = __float_as_uint(w2.x), hz = __float_as_uint(w2.y);
            r0 = __float_as_int(w3.x); r1 = __float_as_int(w3.y); r2 = __float_as_int(w3.z); r3 = __float_as_int(w3.w);
#define TRT_Q(WORD, SEL) __uint_as_float(__byte_perm(WORD, 0x4B000000u, SEL))
#define TRT_SLAB64(SEL, REF, KEY)                                                                                  \
    {                                                                                                              \
        const float ax = fmaf(TRT_Q(lx, SEL), Ax, Bx), bx = fmaf(TRT_Q(hx, SEL), Ax, Bx);                          \
        const float ay = fmaf(TRT_Q(ly, SEL), Ay, By), by = fmaf(TRT_Q(hy, SEL), Ay, By);                          \
        const float az = fmaf(TRT_Q(lz, SEL), Az, Bz), bz = fmaf(TRT_Q(hz, SEL), Az, Bz);                          \
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), sc.box_tmin));            \
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tr.t));                   \
        KEY = (tf >= tn && REF != TRT_CHILD_EMPTY) ? tn : INF;                                                     \
        if (COUNT && REF != TRT_CHILD_EMPTY) cnt.box_tests++;                                                      \
    }
            // __byte_perm selector: result bytes (b3 b2 b1 b0) = (magic.3, magic.2, magic.1, word.k) -> 0x7 6 5 k with the magic as the second operand
            TRT_SLAB64(0x7650, r0, k0)
            TRT_SLAB64(0x7651, r1, k1)
            TRT_SLAB64(0x7652, r2, k2)
            TRT_SLAB64(0x7653, r3, k3)
#undef TRT_SLAB64
#undef TRT_Q
        } else {
        const int rot = tr.node & 3;
        float4 f0a, f0b, f1a, f1b, f2a, f2b, f3a, f3b;
        if (TOP && tr.node < top_k) {
            const float4* rec = s_top + 8 * tr.node;
            f0a = rec[2 * ((0 + rot) & 3)]; f0b = rec[2 * ((0 + rot) & 3) + 1]; f1a = rec[2 * ((1 + rot) & 3)]; f1b = rec[2 * ((1 + rot) & 3) + 1];
            f2a = rec[2 * ((2 + rot) & 3)]; f2b = rec[2 * ((2 + rot) & 3) + 1]; f3a = rec[2 * ((3 + rot) & 3)]; f3b = rec[2 * ((3 + rot) & 3) + 1];
        } else {
            const float4* rec = sc.nodes + 8 * (size_t)tr.node;
            ldg256(rec + 2 * ((0 + rot) & 3), f0a, f0b); ldg256(rec + 2 * ((1 + rot) & 3), f1a, f1b);
            ldg256(rec + 2 * ((2 + rot) & 3), f2a, f2b); ldg256(rec + 2 * ((3 + rot) & 3), f3a, f3b);
        }
#define TRT_SLAB(LX, HX, LY, HY, LZ, HZ, REF, KEY)                                                        \
    {                                                                                                      \
        const float ax = fmaf(LX, idir.x, -ood.x), bx = fmaf(HX, idir.x, -ood.x);                          \
        const float ay = fmaf(LY, idir.y, -ood.y), by = fmaf(HY, idir.y, -ood.y);                          \
        const float az = fmaf(LZ, idir.z, -ood.z), bz = fmaf(HZ, idir.z, -ood.z);                          \
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), sc.box_tmin));    \
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tr.t));           \
        KEY = (tf >= tn && REF != TRT_CHILD_EMPTY) ? tn : INF;                                             \
        if (COUNT && REF != TRT_CHILD_EMPTY) cnt.box_tests++;                                               \
    }
        r0 = __float_as_int(f3a.x); r1 = __float_as_int(f3a.y); r2 = __float_as_int(f3a.z); r3 = __float_as_int(f3a.w);
        TRT_SLAB(f0a.x, f0b.x, f1a.x, f1b.x, f2a.x, f2b.x, r0, k0)
        TRT_SLAB(f0a.y, f0b.y, f1a.y, f1b.y, f2a.y, f2b.y, r1, k1)
        TRT_SLAB(f0a.z, f0b.z, f1a.z, f1b.z, f2a.z, f2b.z, r2, k2)
        TRT_SLAB(f0a.w, f0b.w, f1a.w, f1b.w, f2a.w, f2b.w, r3, k3)
#undef TRT_SLAB
        }
        order2(k0, r0, k1, r1); order2(k2, r2, k3, r3);
        order2(k0, r0, k2, r2); order2(k1, r1, k3, r3);
        order2(k1, r1, k2, r2);
        int sp = tr.sp;
        if (k3 < INF) stack[sp++] = r3;
        if (k2 < INF) stack[sp++] = r2;
        if (k1 < INF) stack[sp++] = r1;
        if (k0 < INF) { tr.node = r0; tr.sp = sp; }
        else { tr.node = stack[sp - 1]; tr.sp = sp - 1; }
    }
    if (tr.node < 0 && pend == 0) {
        pend = tr.node; tr.node = stack[--tr.sp];
        if (exp_flags & 1) prefetch_l1(sc.tri + 4 * (size_t)(~pend & TRT_LEAF_POS_MASK));
    }
    if ((exp_flags & 2) && tr.node >= 0 && tr.node != TRT_TRAV_DONE) prefetch_l1(sc.nodes + 8 * (size_t)tr.node);
    // the parked leaves are tested once a 1/share part of the lanes that are still inside hold one, or no lane can advance
    // (every one of them holds a leaf and has a second leaf or nothing in hand): a lane that is stuck WAITS for company
    // instead of forcing the triangle loop on the whole warp (same-call A/B on C2: 36.6 -> 35.3 ms)
    const bool can_advance = tr.node >= 0 && tr.node != TRT_TRAV_DONE;
    const unsigned pend_mask = __ballot_sync(0xffffffffu, pend != 0);
    const unsigned adv_mask = __ballot_sync(0xffffffffu, can_advance);
    if (__popc(pend_mask) * share >= __popc(adv_mask | pend_mask) || (pend_mask != 0 && adv_mask == 0)) {
#ifdef TRT_COOP_LEAF
        // WARP-COOPERATIVE leaf tests: the triangles of all parked runs are dealt out to all 32 lanes (prefix sum of the run
        // lengths; lane w of a pass takes the w-th triangle and fetches the ray of the lane that owns it through shuffles),
        // tested in one Moller-Trumbore pass, and every owner collects the nearest hit of its own run (first-wins ties by the
        // caller's index, exactly like the sequential loop) - instead of a loop as long as the longest run at ~6 of 32 lanes.
        {
            const int run_n = pend != 0 ? (((~pend) >> TRT_LEAF_POS_BITS) + 1) : 0;
            const int p0 = ~pend & TRT_LEAF_POS_MASK;
            int incl = run_n;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, off); if ((int)(threadIdx.x & 31) >= off) incl += y; }
            const int excl = incl - run_n;
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int ns = sc.n_spheres;
            const int lane_id = threadIdx.x & 31;
            float best_t = tr.t; int best_orig = -2, best_p = -1; float best_u = 0.0f, best_v = 0.0f;      // -2: the incumbent's caller index is not loaded yet
            for (int base = 0; base < total; base += 32) {
                const int w = base + lane_id;
                int owner = 0;
#pragma unroll
                for (int k = 16; k >= 1; k >>= 1) { const int v = __shfl_sync(0xffffffffu, incl, owner + k - 1); if (v <= w) owner += k; }
                owner = min(owner, 31);
                const int tri = __shfl_sync(0xffffffffu, p0 - excl, owner) + w;
                const float3 ro = f3(__shfl_sync(0xffffffffu, o.x, owner), __shfl_sync(0xffffffffu, o.y, owner), __shfl_sync(0xffffffffu, o.z, owner));
                const float3 rdir = f3(__shfl_sync(0xffffffffu, d.x, owner), __shfl_sync(0xffffffffu, d.y, owner), __shfl_sync(0xffffffffu, d.z, owner));
                const int sl = __shfl_sync(0xffffffffu, src_leaf, owner), tl = __shfl_sync(0xffffffffu, twin_leaf, owner);
                float t = __int_as_float(0x7f800000), u = 0.0f, v = 0.0f; int orig = 0x7fffffff;
                if (w < total && tri != sl && tri != tl) {
                    const float4* rec = sc.tri + 4 * (size_t)tri;
                    float4 f0a, f0b, f1a, f1b;
                    ldg256(rec, f0a, f0b); ldg256(rec + 2, f1a, f1b);
                    if (COUNT) cnt.tri_tests++;
                    float tt, uu, vv;
                    if (triangle_test(f0a, f0b, f1a, ro, rdir, sc.tri_tmin, tt, uu, vv)) { t = tt; u = uu; v = vv; orig = __float_as_int(f1b.x); }
                }
                // owners collect from the lanes of their run that fall into this pass
                const int s0 = max(excl, base) - base, s1 = min(incl, base + 32) - base;
                const int seg = (pend != 0 && s1 > s0) ? s1 - s0 : 0;
                const int maxlen = __reduce_max_sync(0xffffffffu, seg);
                int got = -1;
                for (int j = 0; j < maxlen; j++) {
                    const int src = min(max(s0, 0) + j, 31);
                    const float tj = __shfl_sync(0xffffffffu, t, src);
                    const int oj = __shfl_sync(0xffffffffu, orig, src);
                    if (j < seg && tj <= best_t && oj != 0x7fffffff) {
                        bool take = tj < best_t;
                        if (!take) {                                 // tie: the lower index in the caller's array wins (main.c:85 strict '<'); a sphere keeps it
                            if (best_orig == -2) best_orig = tr.prim >= ns ? __ldg(sc.tri_orig + (tr.prim - ns)) : -1;
                            take = oj < best_orig;
                        }
                        if (take) { best_t = tj; best_orig = oj; got = src; }
                    }
                }
                const int gsrc = max(got, 0);
                const float gu = __shfl_sync(0xffffffffu, u, gsrc), gv = __shfl_sync(0xffffffffu, v, gsrc);
                if (got >= 0) { best_u = gu; best_v = gv; best_p = p0 + (base + got - excl); }
            }
            if (pend != 0) {
                pend = 0;
                if (best_p >= 0) {
                    tr.t = best_t; tr.prim = ns + best_p; tr.u = best_u; tr.v = best_v;
                    if (any_hit) { tr.node = TRT_TRAV_DONE; tr.sp = 0; }
                }
            }
        }
#else
        if (pend != 0) {
            int p = ~pend & TRT_LEAF_POS_MASK;
            const int run_len = ((~pend) >> TRT_LEAF_POS_BITS) + 1;
            TRT_CHECK(p >= 0 && p + run_len <= sc.n_tris && run_len <= 16, 103);
            pend = 0;
            const float4* rec = sc.tri + 4 * (size_t)p;
            float4 f0a, f0b, f1a, f1b;
            ldg256(rec, f0a, f0b); ldg256(rec + 2, f1a, f1b);
#define TRT_LEAF_TEST(P, R0, R1, R2, R3)                                                                                  \
    if ((P) != src_leaf && (P) != twin_leaf) {                                                                            \
        float t, u, v;                                                                                                    \
        if (COUNT) cnt.tri_tests++;                                                                                       \
        if (triangle_test(R0, R1, R2, o, d, sc.tri_tmin, t, u, v)) {                                                      \
            const int ns = sc.n_spheres;                                                                                  \
            bool take = t < tr.t;                                                                                         \
            if (t == tr.t && tr.prim >= ns) take = __float_as_int(R3.x) < __ldg(sc.tri_orig + (tr.prim - ns));            \
            if (take) {                                                                                                   \
                tr.t = t; tr.prim = ns + (P); tr.u = u; tr.v = v;                                                         \
                if (any_hit) { tr.node = TRT_TRAV_DONE; tr.sp = 0; left = 0; }                                            \
            }                                                                                                             \
        }                                                                                                                 \
    }
#ifdef TRT_LEAF_PAIRS
            // two triangles of the run per iteration: the runs of the reference's meshes are 2, 4 or 8 long (the halves of a
            // face, its back faces, crossed plant quads), and the second record's fetch overlaps the first one's test
            // (the record after the last one of a run is fetched too: the triangle buffer has 256 bytes of slack for that)
            float4 g0a, g0b, g1a, g1b;
            ldg256(rec + 4, g0a, g0b); ldg256(rec + 6, g1a, g1b);
            int left = run_len;
            for (;;) {
                TRT_LEAF_TEST(p, f0a, f0b, f1a, f1b)
                if (left >= 2) { TRT_LEAF_TEST(p + 1, g0a, g0b, g1a, g1b) }
                left -= 2;
                if (left <= 0) break;
                p += 2; rec += 8;
                ldg256(rec, f0a, f0b); ldg256(rec + 2, f1a, f1b);
                ldg256(rec + 4, g0a, g0b); ldg256(rec + 6, g1a, g1b);
            }
#else
            int left = run_len;
            for (;;) {
                TRT_LEAF_TEST(p, f0a, f0b, f1a, f1b)
                if (--left <= 0) break;
                p++; rec += 4;
                ldg256(rec, f0a, f0b); ldg256(rec + 2, f1a, f1b);
            }
#endif
#undef TRT_LEAF_TEST
        }
#endif
    }
}

// ---- material at the hit ----------------------------------------------------------------------
struct Material {
    float3 diffuse; float alpha;
    float3 emission; float emission_strength;
    float reflection, index;
};

__device__ __forceinline__ Material unpack_material(float4 m0, float4 m1, float4 m2)
{
    Material m;
    m.diffuse = xyz(m0); m.alpha = m0.w;
    m.emission = xyz(m1); m.emission_strength = m1.w;
    m.reflection = m2.x; m.index = m2.y;
    return m;
}

// tri_uvmapping (texture.h:44-90): uv from the barycentrics (weights w,u,v for A,B,C — equal to the
// reference's signed-area barycentrics), tiling by fmod, nearest texel, slab by
// material ordinal.  Per-ordinal overrides are already baked into the texel table.
__device__ __forceinline__ Material triangle_material(const DevScene& sc, int leaf_pos, float u, float v, int& twin_out)
{
    const float4 q0 = __ldg(sc.tri_uv + 2 * leaf_pos), q1 = __ldg(sc.tri_uv + 2 * leaf_pos + 1);
    const float w = 1.0f - u - v;
    float tu = fmaf(w, q0.x, fmaf(u, q0.z, v * q1.x));
    float tv = fmaf(w, q0.y, fmaf(u, q0.w, v * q1.y));
    tu -= truncf(tu); tv -= truncf(tv);                 // fmod(x, 1.0)
    if (tu < 0.0f) tu += 1.0f;
    if (tv < 0.0f) tv += 1.0f;
    const int x = min((int)(tu * (float)sc.tex_w), sc.tex_w - 1);
    const int y = min((int)(tv * (float)sc.tex_h), sc.tex_h - 1);
    const int ordinal = __float_as_int(q1.z);
    twin_out = __float_as_int(q1.w);
    TRT_CHECK(leaf_pos >= 0 && leaf_pos < sc.n_tris && ordinal >= 0 && ordinal < sc.n_materials && x >= 0 && y >= 0, 106);
    const size_t texel = (size_t)(y * sc.tex_w + x) + (size_t)sc.tex_h * (size_t)sc.tex_w * (size_t)ordinal;
    const float4* rec = sc.texels + 3 * texel;
    return unpack_material(__ldg(rec), __ldg(rec + 1), __ldg(rec + 2));
}

// reflected_vec (rtutility.h:205-208)
__device__ __forceinline__ float3 reflect3(float3 v, float3 n) { return fma3(n, -2.0f * dot3(v, n), v); }

// refracted_vec (rtutility.h:210-227): both indices squared first, result not renormalised,
// total internal reflection -> mirror.
__device__ __forceinline__ float3 refract3(float3 v, float3 n, float n1, float n2)
{
    n1 *= n1; n2 *= n2;
    const float rho = n1 / n2;
    const float vn = dot3(n, v);
    const float radical = 1.0f - (rho * rho) * (1.0f - vn * vn);
    if (radical > 0.0f) {
        const float3 tang = (v - n * vn) * rho;
        return tang - n * sqrtf(radical);
    }
    return reflect3(v, n);
}
