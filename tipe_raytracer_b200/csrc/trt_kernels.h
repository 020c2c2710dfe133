// trt_kernels.h — host-visible launch interface of the CUDA translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "trt_device.cuh"

#define TRT_MAX_RESOLVE_SOURCES 16

// Arguments of the path-tracing kernel (passed as one __grid_constant__ struct).
struct RenderArgs {
    DevScene scene;
    int W, H;
    int spp;                // samples rendered by this launch (of every pixel)
    int sample_offset;      // global index of the first of them (Philox counter)
    int bounces;            // nbRebondMax
    int use_ao;
    float inv_ao;           // 1 / AO_intensity
    float sample_clamp;     // |per-sample value| bound before quantisation, see to_fixed
    float emit_scale_ao;    // 1.5 * AO_intensity
    float focus, aperture_x, aperture_y;
    float inv_wm1, inv_hm1; // 1/(W-1), 1/(H-1)
    float3 cam_origin, cam_hor, cam_ver, cam_corner;   // cam_corner = coin_bas_gauche - origin (FP64 on host)
    uint32_t key0, key1;
    // work decomposition
    int chunk, n_chunks, tiles_x, tiles_y;
    int pool;                           // items a warp reserves per atomic (megakernel)
    int tile_offset, tile_stride;       // this launch owns the tiles tile_offset + k * tile_stride (tile split; 0 / 1 = all)
    unsigned per_tile, m_per_tile, m_chunks, m_tiles_x; int items_fit32;   // decode_item: divisors' reciprocals (fastdiv)
    int min_steps, exit_num;            // BVH-loop policy, see pathtrace.cu
    int start_min;                      // megakernel: new paths start once this many lanes of the warp need one (or none has a path left)
    int max_hk, hk_min;                 // megakernel: housekeeping passes per trip; repeated while fewer than hk_min/32 of the live lanes have BVH work
    int defer;                          // megakernel: > 0 = deferred leaves, tested once 1/defer of the lanes inside hold one (bvh_trip_deferred)
    int item_order;                     // see decode_item (trt_shade.cuh)
    int top_nodes;                      // megakernel: this many wide nodes (the top levels) are staged into shared memory by TMA; 0 = off (TRT_TUNE_TOPK)
    // engine 4: the primary rays were traced as packets (primary_packet.cu); prehit[pixel * prehit_slab + sample] = (t, bits(prim), u, v)
    const float4* prehit; int prehit_slab;
    int exp_flags;                      // experiments (TRT_TUNE_EXP): 1 = L1 prefetch of a leaf's triangles when it is parked, 2 = of the next node after a visit
    // slot-pool kernel (pathtrace_pool.cu): n_items = paths of this launch, pool = paths a warp reserves per atomic
    int pp_slots;                       // path slots per warp (<= 64)
    int pp_shade_at;                    // a shading pass starts once this many slots are hit-ready
    int pp_refill_min;                  // lanes re-arm from the ray-ready ring once this many of the warp are idle
    int pp_exit_below;                  // ... or the warp shades what it has when fewer rays than this are in flight or waiting
    unsigned pp_per_tile, pp_m_per_tile, pp_m_spp; int pp_fit32;   // decode_path: 32 * spp and the reciprocals (fastdiv)
    int count_tests;
    long long n_items;
    unsigned long long* item_counter;   // zeroed before launch
    unsigned long long* counters;       // [0] rays [1] box tests [2] triangle tests [3] sphere tests
    long long* accum;                   // planar [9][W*H] int64, value * 2^28
};

struct ResolveArgs {
    int W, H, divisor, n_src;
    size_t p0, p1;                      // pixels [p0, p1) of the frame; the output pointers are relative to p0 (a device resolves its own band)
    const long long* src[TRT_MAX_RESOLVE_SOURCES];
    double *canva, *albedo, *normal;    // [W*H][3]
    unsigned char* canva8;              // when set, canva is written here as 3 bytes per pixel instead (exact: the values are integers 0..255)
};

cudaError_t trt_launch_pathtrace(const RenderArgs& A, int n_sms, int blocks_per_sm, bool brute, bool count_tests, cudaStream_t stream);
// slot-pool kernel for triangle meshes (pathtrace_pool.cu); blocks_per_sm 0 = default
cudaError_t trt_launch_pathtrace_pool(const RenderArgs& A, int n_sms, int blocks_per_sm, bool count_tests, cudaStream_t stream);
size_t trt_pool_smem_bytes(const RenderArgs& A);
// packet traversal of the primary rays into hits[pixel * slab + sample] (primary_packet.cu)
cudaError_t trt_launch_primary_packets(const RenderArgs& A, int n_sms, float4* hits, int slab, bool count_tests, cudaStream_t stream);
cudaError_t trt_launch_primary_fp32(const RenderArgs& A, int* id, float* t, cudaStream_t stream);     // K2's FP32 first hit, pixel centres
cudaError_t trt_launch_resolve(const ResolveArgs& R, int n_sms, cudaStream_t stream);
cudaError_t trt_launch_resolve_hdr(int npix, int divisor, const long long* acc, float* radiance, float* albedo, float* normal, int n_sms, cudaStream_t stream);

// K1: FP64 primary visibility, linear scan in the caller's order.
struct PrimaryArgs {
    DevScene scene;
    int W, H;
    double focus;
    double cam[12];                     // origin, horizontal, vertical, coin_bas_gauche
    double* t; int* id; double* normal; double* albedo;    // any may be null
};
cudaError_t trt_launch_primary(const PrimaryArgs& P, cudaStream_t stream);

// Scene conversion + BVH build (K3).  Raw = the caller's FP64 AoS arrays, already on the device.
struct BuildInputs {
    const double* spheres_raw;   int n_spheres;     // 14 doubles each
    const double* tris_raw;      int n_tris;        // 25 doubles each
    const float*  mats_packed;   int tex_w, tex_h, n_materials;   // 10 floats per texel: the caller's material structs, rounded to FP32 on the host
    const int*    quel_mat;                         // [n_tris]
    const double* overrides_raw; int n_overrides;   // trt_mat_override, 8 doubles each (first = ordinal|mask)
    void* scratch; size_t scratch_bytes;            // device scratch of at least trt_build_scratch_bytes(n_tris)
};
struct BuildOutputs {            // device buffers allocated by the caller with these element counts
    float4* sph;        // n_spheres
    float4* sph_mat;    // 3*n_spheres
    float4* tri;        // 4*n_tris (+ slack: the unified fetch may read 64 B past a record)
    float4* tri_uv;     // 2*n_tris
    int*    tri_orig;   // n_tris
    float4* nodes;      // 8*max(n_tris-1,1): 4-wide nodes, 128 B each (at most n_tris-1 of them)
    float4* texels;     // 3*n_materials*tex_w*tex_h
    // filled by the build
    int root;
    float tri_tmin;
    float3 root_lo, root_hi;   // padded bounds of all triangles
    int max_depth, n_twins, n_wide_nodes, wide_depth, binary_depth, n_groups;   // n_groups = leaves of the hierarchy (runs of equal boxes)
    double sah_cost;
};
cudaError_t trt_build_scene(const BuildInputs& in, BuildOutputs& out, cudaStream_t stream);
size_t trt_build_scratch_bytes(int n_tris);
// experiment: the built 128-byte nodes as 64-byte nodes with 8-bit quantised child boxes (n_nodes * 64 bytes at nodes64)
cudaError_t trt_compress_nodes(int n_nodes, const float4* nodes128, void* nodes64, cudaStream_t stream);

// Known-answer kernels (leaf functions of K2 run over arrays).
cudaError_t trt_kat_launch_sphere(int n, const float4* cr, const float* ray, int* hit, float* dst, float* point, float* normal, cudaStream_t s);
cudaError_t trt_kat_launch_triangle(int n, const float* abc, const float* ray, int* hit, float* dst_u_v, cudaStream_t s);
cudaError_t trt_kat_launch_refract(int n, const float* in8, float* out3, cudaStream_t s);
cudaError_t trt_kat_launch_philox(int n, const uint32_t* ctr_key, uint32_t* out4, cudaStream_t s);
cudaError_t trt_kat_launch_tonemap(int n, const long long* sums, int spp, double* out, cudaStream_t s);
