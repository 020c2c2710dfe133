// trt_shade.cuh — the per-hit logic of the reference's tracer() (main.c:118-242) and the per-sample
// set-up of fill_canva (main.c:264-272), shared by both execution engines (the register-resident
// megakernel of pathtrace.cu and the wavefront engine of wavefront.cu), so that there is exactly one
// FP32 restatement of the shading semantics.
#pragma once

#include "trt_isect.cuh"
#include "trt_kernels.h"

enum { PH_NEW = 0, PH_BOUNCE = 1, PH_AO = 2, PH_DONE = 3 };

struct PathState {
    float3 o, d;            // current ray (d is NOT normalised after a mirror/diffuse lerp, main.c:165)
    float3 d_ao;            // pending ambient-occlusion direction
    float3 T, L;            // rayColor, incomingLight (main.c:120-121)
    float3 aov_albedo, aov_normal;
    float  cur_index;       // the refraction stack collapsed to its top n2 (pile.h)
    int    bounce;          // i of main.c:132
    int    alpha_depth;
    int    src_prim, src_twin;
    bool   is_alpha;
    bool   ao_pending;      // fused mode: the AO ray of the last shaded bounce is traced together with the next bounce ray
};

struct PathResult {         // what one finished path adds to its pixel (col_alb_norm, denoiser.h:11-13)
    float3 L, albedo, normal;
    int dead_ao;            // 1: the path ended on a shaded bounce whose ambient-occlusion ray was counted but not traced (see consume_hit)
};

// item index -> (pixel slot p of an 8x4 tile, sample chunk, tile).  An item is (pixel, chunk of samples); 32*n_chunks
// consecutive items cover one tile.  order 0: pixel fastest (32 consecutive items = the 32 pixels of the tile, one
// chunk); order 1: chunk fastest (32 consecutive items = all chunks of 2-3 neighbouring pixels, so the lanes of a warp
// start on nearly identical primary rays: coherent traversal, node fetches that coalesce).
// x / d and x % d for x < 2^32 with a host-computed m = min(floor(2^32 / d), 2^32 - 1): the quotient estimate is at most
// one short, one correction step makes it exact (an integer division costs 20 instructions, the 64-bit one 80).
__device__ __forceinline__ unsigned fastdiv(unsigned x, unsigned d, unsigned m, unsigned& rem)
{
    unsigned q = __umulhi(x, m);
    unsigned r = x - q * d;
    if (r >= d) { q++; r -= d; }
    rem = r;
    return q;
}

__device__ __forceinline__ void decode_item(const RenderArgs& A, long long item, int& p, int& chunk, int& tile, int& tx, int& ty)
{
    if (A.items_fit32) {
        unsigned local, t, c, x;
        t = fastdiv((unsigned)item, A.per_tile, A.m_per_tile, local);
        if (A.item_order == 0) { p = (int)(local & 31u); chunk = (int)(local >> 5); }
        else { p = (int)fastdiv(local, (unsigned)A.n_chunks, A.m_chunks, c); chunk = (int)c; }
        t = t * (unsigned)A.tile_stride + (unsigned)A.tile_offset;       // tile split: the k-th tile of this launch
        tile = (int)t;
        ty = (int)fastdiv(t, (unsigned)A.tiles_x, A.m_tiles_x, x); tx = (int)x;
        return;
    }
    const int per_tile = 32 * A.n_chunks;
    tile = (int)(item / per_tile);
    const int local = (int)(item - (long long)tile * per_tile);
    if (A.item_order == 0) { p = local & 31; chunk = local >> 5; }
    else { chunk = local % A.n_chunks; p = local / A.n_chunks; }
    tile = tile * A.tile_stride + A.tile_offset;
    tx = tile % A.tiles_x; ty = tile / A.tiles_x;
}

// path index -> (pixel of an 8x4 tile, sample).  Paths are numbered sample-fastest, then the 32 pixels of a tile, then tiles.
__device__ __forceinline__ void decode_path(const RenderArgs& A, long long g, int& px, int& py, int& sample)
{
    int p, tile;
    if (A.pp_fit32) {
        unsigned rem, s, x;
        unsigned t = fastdiv((unsigned)g, A.pp_per_tile, A.pp_m_per_tile, rem);
        p = (int)fastdiv(rem, (unsigned)A.spp, A.pp_m_spp, s);
        sample = (int)s;
        t = t * (unsigned)A.tile_stride + (unsigned)A.tile_offset;
        const int ty = (int)fastdiv(t, (unsigned)A.tiles_x, A.m_tiles_x, x);
        px = (int)x * 8 + (p & 7); py = ty * 4 + (p >> 3);
        return;
    }
    const long long per_tile = 32LL * A.spp;
    const long long tl = g / per_tile;
    const long long rem = g - tl * per_tile;
    p = (int)(rem / A.spp); sample = (int)(rem - (long long)p * A.spp);
    tile = (int)(tl * A.tile_stride + A.tile_offset);
    px = (tile % A.tiles_x) * 8 + (p & 7); py = (tile / A.tiles_x) * 4 + (p >> 3);
}

// fill_canva's per-sample set-up (main.c:265-271) + get_ray (camera.h:42-54): the primary ray of (pixel, sample).
// A.cam_corner = coin_bas_gauche - origin, formed in FP64 on the host.
__device__ __forceinline__ void camera_ray(const RenderArgs& A, int pixel, int px, int py, int sample, float3& o, float3& d)
{
    const uint4 r = philox4x32_10((uint32_t)pixel, (uint32_t)(A.sample_offset + sample), 0u, 0u, A.key0, A.key1);
    const float u = ((float)px + (u01(r.x) - 0.5f)) * A.inv_wm1;
    const float v = ((float)py + (u01(r.y) - 0.5f)) * A.inv_hm1;
    const float dx = (u01(r.z) - 0.5f) * A.aperture_x;
    const float dy = (u01(r.w) - 0.5f) * A.aperture_y;
    const float3 direction = fma3(A.cam_hor, u, fma3(A.cam_ver, v, A.cam_corner));
    const float3 dest = fma3(direction, A.focus, A.cam_origin);
    o = f3(A.cam_origin.x + dx, A.cam_origin.y + dy, A.cam_origin.z);
    d = normalize3(dest - o);
}

// tracer()'s locals at entry (main.c:119-130)
__device__ __forceinline__ void init_path_state(PathState& P)
{
    P.T = f3(1, 1, 1); P.L = f3(0, 0, 0);
    P.aov_albedo = P.aov_normal = f3(0, 0, 0);
    P.cur_index = 1.0f; P.bounce = 0; P.alpha_depth = 0; P.is_alpha = false;
    P.src_prim = TRT_NO_PRIM; P.src_twin = -1;
    P.ao_pending = false;
}

__device__ __forceinline__ void start_path(const RenderArgs& A, int pixel, int px, int py, int sample, PathState& P)
{
    camera_ray(A, pixel, px, py, sample, P.o, P.d);
    init_path_state(P);
}

// One completed closest_hit query -> one trip of the bounce loop (main.c:132-239) or the end of an
// ambient-occlusion query (main.c:94-116).  `phase` is PH_BOUNCE or PH_AO on entry.
// Returns true when the path has ended (res is then what it contributes); otherwise P holds the
// next ray (P.o with P.d, or P.d_ao when phase == PH_AO on exit).
// sph / sph_mat may point to shared or global memory.
__device__ __forceinline__ bool consume_hit(const RenderArgs& A, const float4* __restrict__ sph, const float4* __restrict__ sph_mat,
                                            int pixel, int sample, float hit_t, int hit_prim, float hit_u, float hit_v,
                                            PathState& P, int& phase, PathResult& res, bool fuse_ao = false)
{
    const DevScene& sc = A.scene;
    const bool did_hit = hit_prim != TRT_NO_PRIM;
    res.dead_ao = 0;

    if (phase == PH_AO) {
        // ambient_occlusion (main.c:94-116): distance/dst == 1 for a unit direction, so the term is
        // 1/AO when anything is hit and 0 otherwise.
        P.T = P.T * (did_hit ? A.inv_ao : 0.0f);
        res.L = P.L; res.albedo = P.aov_albedo; res.normal = P.aov_normal;
        if (P.bounce >= A.bounces) return true;
        phase = PH_BOUNCE;
        return false;
    }

    const int i = P.bounce;
    Material m;
    float3 n = f3(0, 0, 0), hp = P.o;
    int twin = -1;
    if (did_hit) {
        hp = fma3(P.d, hit_t, P.o);
        if (hit_prim < sc.n_spheres) {
            const float4 cr = sph[hit_prim];
            n = normalize3(hp - xyz(cr));                                    // sphere.h:33 (always outward)
            m = unpack_material(sph_mat[3 * hit_prim], sph_mat[3 * hit_prim + 1], sph_mat[3 * hit_prim + 2]);
            if (sc.sky != nullptr && hit_prim == sc.n_spheres - 1) {
                // "cas du ciel" (main.c:64-71, commented out upstream) + sphere_uvmapping (texture.h:92-112): the last
                // sphere emits the lat-long sky texel under the hit point and is opaque.  Index clamped like the oracle's.
                const float inv_r = 1.0f / cr.w;
                const float dx = (hp.x - cr.x) * inv_r, dy = (hp.y - cr.y) * inv_r, dz = (hp.z - cr.z) * inv_r;
                const float theta = acosf(fminf(fmaxf(-dy, -1.0f), 1.0f));
                const float phi = atan2f(-dz, dx) + 3.14159265358979f;
                const int x = (int)(phi * 0.159154943091895f * (float)sc.sky_w);
                const int y = (int)(theta * 0.318309886183791f * (float)sc.sky_h);
                const int idx = min(max(y * sc.sky_w + x, 0), sc.sky_w * sc.sky_h - 1);
                m.emission = xyz(__ldg(sc.sky + idx));
                m.alpha = 1.0f;
            }
        } else {
            const int leaf = hit_prim - sc.n_spheres;
            const float4 r2 = __ldg(sc.tri + 4 * leaf + 2);
            n = normalize3(f3(r2.y, r2.z, r2.w));                            // mesh.h:90
            m = triangle_material(sc, leaf, hit_u, hit_v, twin);
        }
    } else {
        // the reference reads an uninitialised HitInfo here (main.c:135-139); defined as zeros
        m = unpack_material(make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0));
    }

    if (i == 0) { P.aov_albedo = m.diffuse; P.aov_normal = n; }                                   // main.c:137-140
    if (i == P.alpha_depth && P.is_alpha) {                                                       // main.c:142-150
        P.aov_albedo = (m.emission_strength > 0.0f) ? m.emission : m.diffuse;
        P.aov_normal = n;
        P.is_alpha = false;
    }
    res.L = P.L; res.albedo = P.aov_albedo; res.normal = P.aov_normal;

    if (!did_hit) return true;                                                                    // main.c:236-238
    if (i == P.alpha_depth && m.emission_strength > 0.0f) {
        // light seen directly (main.c:154-160): HSL round trip with unit gains == identity
        res.L = m.emission; res.albedo = m.emission; res.normal = n;
        return true;
    }

    // Last trip of main.c:132's loop: the ray this bounce sets up is never traced, so its direction (Philox block, two unit vectors,
    // reflect, lerp, refract) is dead code in the reference and is not computed here; what still reaches the pixel is this hit's
    // emission, and on a refractive material the coin that decides whether it counts.  Same image, bit for bit.
#ifdef TRT_KEEP_DEAD_AO
    const bool last = false;
#else
    const bool last = i + 1 >= A.bounces;
#endif
    const bool refractive = m.alpha <= 0.99f && m.alpha >= 0.0001f;
    const float3 d_in = P.d;
    P.o = hp;                                                                                     // main.c:162
    P.src_prim = hit_prim; P.src_twin = twin;
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    float3 mixed = d_in;
    if (!last || refractive) {
        r = philox4x32_10((uint32_t)pixel, (uint32_t)(A.sample_offset + sample), 1u + 2u * (uint32_t)i, 0u, A.key0, A.key1);
        const float3 diffuse_dir = normalize3(n + unit_sphere(u01(r.x), u01(r.y)));               // main.c:163
        const float3 mirror_dir = reflect3(d_in, n);
        mixed = lerp3(diffuse_dir, mirror_dir, m.reflection);                                     // main.c:165
    }
    bool shade = true;
    if (refractive) {                                                                             // main.c:167-193
        float3 nn = n;
        float n1 = P.cur_index, n2 = m.index;                  // entering: (top.n2, m), push
        const float before = P.cur_index;
        P.cur_index = m.index;
        if (dot3(d_in, n) > 0.0f) {                            // leaving: pop, swap
            nn = -n; n1 = m.index; n2 = before; P.cur_index = before;
        }
        const float3 refr = refract3(d_in, nn, n1, n2);
        // the refraction coin (main.c:185) comes out of the bounce's own Philox block: u01() reads the top 24 bits of a word, so the low
        // bytes of words 0-2 are 24 random bits that no direction has seen (a second block per refractive hit cost 7 % of C3)
        const uint32_t coin = (__byte_perm(__byte_perm(r.x, r.y, 0x0040), r.z, 0x0410) & 0x00ffffffu) << 8;
        if (u01(coin) > m.alpha) { P.d = refr; shade = false; }
        else P.d = mixed;
    }
    if (m.alpha > 0.99f) { P.is_alpha = false; P.d = mixed; }                                     // main.c:195-198
    if (shade && m.alpha < 0.0001f) { P.is_alpha = true; P.alpha_depth++; shade = false; }        // main.c:200-206
    P.bounce = i + 1;
    if (shade) {
        const float k = A.use_ao ? A.emit_scale_ao : 1.0f;                                        // 1.5 * AO (main.c:210)
        P.L = P.L + (m.emission * (m.emission_strength * k)) * P.T;
        if (P.T.x > 0.5f || P.T.y > 0.5f || P.T.z > 0.5f) P.T = m.diffuse * (P.T * 1.3f);         // main.c:214-216
        P.T = m.diffuse * P.T;
        if (A.use_ao) {
#ifndef TRT_KEEP_DEAD_AO          /* -DTRT_KEEP_DEAD_AO traces it like the reference (for A/B runs: same image, same ray count) */
            if (P.bounce >= A.bounces) {
                // last trip of main.c:132: ambient_occlusion (main.c:220-221) only scales rayColor, which nothing reads after the
                // loop.  The reference traces that ray all the same; here it is COUNTED (trt_stats.rays stays the reference's
                // closest_hit count, trt_stats.rays_skipped says how many of them were not traced) and the path ends now.
                res.dead_ao = 1;
            } else
#endif
            {
                P.d_ao = normalize3(n + unit_sphere(u01(r.z), u01(r.w)));                         // main.c:100-102
                if (fuse_ao) { P.ao_pending = true; if (P.bounce >= A.bounces) phase = PH_AO; }   // AO ray rides with the next bounce ray
                else phase = PH_AO;
            }
        }
    }
    res.L = P.L;
    return phase != PH_AO && P.bounce >= A.bounces;
}
