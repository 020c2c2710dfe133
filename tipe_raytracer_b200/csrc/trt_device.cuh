// trt_device.cuh — device-side building blocks of the sm_100a path tracer: float3 helpers,
// Philox4x32-10, fixed-point sample accumulation, the device scene layout.
//
// FP32 restatements of the reference's leaf maths live in trt_isect.cuh; this file only has
// what several kernels share.  Nothing here touches the CPU oracle.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define TRT_ACCUM_CHANNELS_D 9
#define TRT_ACCUM_SCALE 268435456.0f          /* 2^28, TRT_ACCUM_FRAC_BITS */
#define TRT_ACCUM_INV_SCALE (1.0 / 268435456.0)
#define TRT_SAMPLE_CLAMP 1048576.0f           /* |per-sample value| <= 2^20 before quantisation */

#define TRT_MAX_SMEM_SPHERES 1024
#define TRT_STACK_DEPTH 96                    /* traversal stack entries per lane (local memory) */
#define TRT_CHILD_EMPTY 0x7ffffffe             /* unused child slot of a 4-wide node */
/* leaf reference = ~(position of the run's first triangle | (run length - 1) << 27): the traversal knows how many records to
   test without waiting for the first one (run length <= 16, positions < 2^27) */
#define TRT_LEAF_POS_BITS 27
#define TRT_LEAF_POS_MASK ((1 << TRT_LEAF_POS_BITS) - 1)
#define TRT_LEAF_REF(pos, count) (~((pos) | (((count) - 1) << TRT_LEAF_POS_BITS)))

// ---- device scene (all pointers are device pointers owned by the ctx) -----------------------
struct DevScene {
    // spheres, in the caller's order (closest_hit scans them first, main.c:59-78)
    int n_spheres;
    const float4* sph;          // [n]   (cx, cy, cz, radius)
    const float4* sph_mat;      // [3n]  material record, see below
    const double* sph64;        // [14n] raw copy of the caller's sphere array (112 B each) for the FP64 primary pass
    // triangles, stored in BVH leaf order
    int n_tris;
    int n_nodes;                // wide nodes of the BVH (range checks of -DTRT_DEBUG_BOUNDS builds)
    const float4* tri;          // [4n]  64 B: (A.xyz, AB.x) (AB.yz, AC.xy) (AC.z, N.xyz) (caller index, twin leaf, ordinal, -), N = AB x AC un-normalised
    const float4* tri_uv;       // [2n]  (uvA, uvB) (uvC, bits(material ordinal), bits(twin leaf position or -1))
    const int*    tri_orig;     // [n]   leaf position -> index in the caller's triangle array
    const double* tri64;        // raw copy of the caller's triangle array (200 B each), caller order
    const int*    quel_mat;     // [n]   caller order
    // 4-wide BVH, 128-byte nodes (one L1 line): lo.x[4] hi.x[4] lo.y[4] hi.y[4] lo.z[4] hi.z[4] child[4] pad
    // child reference: >= 0 wide node, < 0 leaf ~(leaf position), TRT_CHILD_EMPTY unused slot;
    // the four 32-byte fields {x, y, z, children} of node i sit at slot (field + i) & 3 (L1 bank spreading)
    const float4* nodes;
    const float4* nodes64;      // the same tree as 64-byte nodes with 8-bit quantised child boxes (experiment; nullptr = not built)
    int root;                   // 0, or ~0 when there is a single triangle
    float3 root_lo, root_hi;    // bounds of all triangles (rays that miss them never enter the traversal loop)
    // textures: nbMaterials slabs of tex_h*tex_w texels, 3 float4 per texel, overrides baked in
    const float4* texels;
    int tex_w, tex_h, n_materials;
    float tri_tmin;             // FP32 stand-in for the 1e-7 distance threshold of mesh.h:88
    float box_tmin;             // start of the ray interval in the slab tests: tri_tmin (a box the ray leaves before that cannot hold an accepted hit) or 0
    // sky environment (trt_sky_set; nullptr = off): lat-long image, (r, g, b, 0) per texel, row 0 = bottom PPM row
    const float4* sky;
    int sky_w, sky_h;
};
// material record (sphere or texel), 3 x float4:
//   [0] diffuse.rgb, alpha   [1] emission.rgb, emissionStrength   [2] reflectionStrength, materialIndex, 0, 0

// ---- -DTRT_DEBUG_BOUNDS: range checks of every index the kernels form (compute-sanitizer stand-in) -----------------------
// A violation sets *trt_debug_flag (counters[12]; the render call then fails with the code) and the access is skipped / clamped.
#ifdef TRT_DEBUG_BOUNDS
static __device__ unsigned long long* trt_debug_flag_ptr;      // one per translation unit; every kernel sets it on entry (TRT_DEBUG_INIT)
#define TRT_DEBUG_INIT(A) do { if (threadIdx.x == 0) trt_debug_flag_ptr = (A).counters + 12; __syncthreads(); } while (0)
#define TRT_CHECK(cond, code) do { if (!(cond)) { if (trt_debug_flag_ptr) atomicMax(trt_debug_flag_ptr, (unsigned long long)(code)); } } while (0)
#else
#define TRT_CHECK(cond, code) do { } while (0)
#define TRT_DEBUG_INIT(A) do { } while (0)
#endif

// ---- float3 helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b)
{
    return f3(fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ float3 fma3(float3 a, float s, float3 b) { return f3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); }
__device__ __forceinline__ float rsq_approx(float x);
__device__ __forceinline__ float3 normalize3(float3 a) { float r = rsq_approx(dot3(a, a)); return a * r; }
__device__ __forceinline__ float3 lerp3(float3 a, float3 b, float t) { return fma3(b - a, t, a); }   // rtutility.h:32-34
__device__ __forceinline__ float3 xyz(float4 v) { return f3(v.x, v.y, v.z); }

// ---- Philox4x32-10 (Salmon et al. 2011) --------------------------------------------------------
// key = (seed lo, seed hi); counter = (pixel_index, global sample index, slot, 0).
//   slot 0       : pixel jitter u, v, aperture dx, dy                          (main.c:265-269)
//   slot 1 + 2*i : bounce i: .x .y diffuse direction, .z .w AO direction       (main.c:163, main.c:100)
//   slot 2 + 2*i : bounce i: .x refraction coin                                (main.c:185)
// A draw depends only on (seed, pixel, sample, bounce), never on the launch geometry or the GPU
// that computes it, which is what makes the image bit-identical across runs and GPU counts.
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1;
        c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// 24-bit uniform in [0,1): exactly representable in FP32, so the FP64 oracle sees the same value.
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// random_dir_no_norm (rtutility.h:189-203): uniform on the unit sphere from (u, v).
// z = cos(acos(2v-1)) = 2v-1, sin(phi) = sqrt(1-z^2); theta = 2*pi*u.
__device__ __forceinline__ float3 unit_sphere(float u, float v)
{
    const float z = fmaf(2.0f, v, -1.0f);
    const float s = sqrtf(fmaxf(0.0f, fmaf(-z, z, 1.0f)));
    float sn, cs;
    __sincosf(6.283185307179586f * u, &sn, &cs);       // MUFU.SIN/COS: abs error ~2^-21 over [0, 2*pi), ample for a sampling direction
    return f3(cs * s, sn * s, z);
}

// ---- fixed-point accumulation --------------------------------------------------------------------
// Integer sums are associative: the image does not depend on which lane / GPU added which sample.
// clamp = RenderArgs.sample_clamp: 2^20 for up to 2^14 samples per pixel, 2^34 / samples beyond, so that even a pixel whose
// every sample saturates cannot wrap the int64 sum (2^28 * 2^34 = 2^62)
__device__ __forceinline__ long long to_fixed(float v, float clamp)
{
    v = (v != v) ? 0.0f : fminf(fmaxf(v, -clamp), clamp);
    return __float2ll_rn(v * TRT_ACCUM_SCALE);
}

// 256-bit read-only load (sm_100a: LDG.E.256.CONSTANT).  A divergent gather is bound by L1 wavefronts, one per
// lane per load instruction: fetching a 128-byte node with four 256-bit loads instead of eight 128-bit ones
// doubles the node rate (tools/microbench/gather.cu: 36 -> 72 G records/s on B200).  p must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }

// MUFU.RSQ without rsqrtf()'s denormal-input rescue (4 more instructions); callers pass values far from 1e-38
__device__ __forceinline__ float rsq_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
