// pathtrace.cu — K2, the persistent path-tracing kernel, and K4, the resolve (tonemap) kernel.
//
// Replaces fill_canva + tracer + closest_hit + ambient_occlusion of the reference
// (main.c:245-284, 118-242, 52-92, 94-116) for a range of samples of every pixel.
//
// Execution model (sm_100a, 148 SMs):
//   * persistent CTAs, a multiple of the SM count; no per-pixel thread mapping;
//   * the unit of scheduling is an ITEM = (pixel, chunk of consecutive samples).  A warp reserves a block of consecutive
//     items from one global counter (one atomic per block) and hands them to its lanes as they run dry.  Items are numbered
//     pixel-fastest on sphere scenes (a block = one 8x4 tile at one chunk) and chunk-fastest on meshes (a block = the chunks
//     of two or three neighbouring pixels: near-identical primary rays);
//   * meshes: every lane re-arms on its own ("ray regeneration"), so a lane never waits for another lane's path; the BVH
//     loop is resumable and is left as soon as too few lanes are still inside, leaves are deferred and tested for the warp
//     together (trt_isect.cuh, bvh_trip_deferred);  sphere scenes: the lanes of a warp start their paths together and run
//     them in step (RenderArgs.start_min), every trip is one sphere scan for all lanes;
//   * a lane keeps the 9 sums of its item as int64 fixed point in shared memory (18 registers saved) and flushes them
//     with 9 RED.ADD.64 when the item ends: integer adds commute, so the image is bit-identical for any
//     schedule, any grid and any split of the samples or tiles over GPUs;
//   * spheres (and their materials) are staged once per CTA into shared memory; BVH nodes,
//     triangles and texels are read through the read-only path and live in L1/L2 (the shared-memory carve-out is kept
//     to what the resident CTAs need, so that L1 has 192 KB on meshes).
#include "trt_shade.cuh"

#define PT_THREADS 256
#ifndef PT_MIN_BLOCKS
#define PT_MIN_BLOCKS 4           // sphere-only kernel: 64 registers, 32 warps/SM (C1: 2.7 % faster than 3 CTAs at 80 registers)
#endif
#ifndef PT_MIN_BLOCKS_MESH
#define PT_MIN_BLOCKS_MESH 3      // mesh kernel: 80 registers, no spills, 24 warps/SM (C2 +0.6 %, C4 +1.6 % over 4 CTAs at 64 registers with 158 B of spills)
#endif
// Traversal-loop exit policy (RenderArgs.min_steps / exit_num): a warp always takes min_steps BVH steps,
// then leaves the loop as soon as fewer than exit_num/32 of its live lanes are still inside the BVH,
// so that the lanes that are done get shaded and re-armed instead of idling.

template <bool HAS_TRIS, bool BRUTE, bool COUNT, bool DEFER, bool TOP = false, bool PREHIT = false, bool N64 = false>
__global__ void __launch_bounds__(PT_THREADS, HAS_TRIS ? PT_MIN_BLOCKS_MESH : PT_MIN_BLOCKS)
pathtrace_kernel(const __grid_constant__ RenderArgs A)
{
    extern __shared__ float4 smem[];
    float4* s_sph = smem;                                   // [n_spheres]
    float4* s_mat = smem + A.scene.n_spheres;               // [3*n_spheres]
    long long* s_acc = (long long*)(smem + 4 * A.scene.n_spheres);   // [9][PT_THREADS] per-lane sums of the current item
    // TOP: the first A.top_nodes wide nodes of the BVH (its top levels) staged into shared memory with one TMA bulk copy
    // (cp.async.bulk global -> shared, completion on an mbarrier): north_star's "top levels staged through shared memory or TMA"
    float4* s_top = (float4*)(s_acc + 9 * PT_THREADS);
    if (TOP) {
        unsigned long long* bar = (unsigned long long*)(s_top + 8 * A.top_nodes);
        const unsigned bar_s = (unsigned)__cvta_generic_to_shared(bar), dst_s = (unsigned)__cvta_generic_to_shared(s_top);
        const unsigned bytes = (unsigned)A.top_nodes * 128u;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst_s), "l"(A.scene.nodes), "r"(bytes), "r"(bar_s) : "memory");
        }
        __syncthreads();
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_s) : "memory");
    }
    for (int k = threadIdx.x; k < A.scene.n_spheres; k += blockDim.x) s_sph[k] = A.scene.sph[k];
    for (int k = threadIdx.x; k < 3 * A.scene.n_spheres; k += blockDim.x) s_mat[k] = A.scene.sph_mat[k];
#pragma unroll
    for (int c = 0; c < 9; c++) s_acc[c * PT_THREADS + threadIdx.x] = 0;
    __syncthreads();

    TRT_DEBUG_INIT(A);
    const DevScene& sc = A.scene;
    const int lane = threadIdx.x & 31;
    const int W = A.W, H = A.H;
    const size_t npix = (size_t)W * (size_t)H;
    long long* my_acc = s_acc + threadIdx.x;

    // item bookkeeping (warp-uniform)
    long long pool_next = 0, pool_end = 0;
    bool global_dry = false;

    // lane state
    int phase = PH_NEW;
    int pixel = -1, px = 0, py = 0, s_cur = 0, s_end = 0;
    PathState P;
    P.o = P.d = P.d_ao = P.T = P.L = P.aov_albedo = P.aov_normal = f3(0, 0, 0);
    P.cur_index = 1.0f; P.bounce = 0; P.alpha_depth = 0; P.src_prim = TRT_NO_PRIM; P.src_twin = -1; P.is_alpha = false; P.ao_pending = false;
    const bool fuse_ao = !HAS_TRIS && A.use_ao;      // sphere-only scenes: AO ray + next bounce ray in one scan
    Trav tr; tr.node = TRT_TRAV_DONE; tr.sp = 0; tr.t = 0.0f; tr.prim = TRT_NO_PRIM; tr.u = tr.v = 0.0f;
    int stack[TRT_STACK_DEPTH];
    int pend = 0;                           // DEFER: the parked leaf (0 = none), see bvh_trip_deferred

    unsigned ray_count = 0, skipped = 0;
    TraceCounters cnt; cnt.box_tests = cnt.tri_tests = cnt.sphere_tests = 0;
    // schedule counters (COUNT builds only, warp-uniform): BVH-loop trips, lanes inside / at a node / at a leaf per trip,
    // loop sessions, housekeeping passes, lanes consuming a hit per pass
    unsigned long long sc_trips = 0, sc_inside = 0, sc_node = 0, sc_leaf = 0, sc_sessions = 0, sc_hk = 0, sc_consume = 0;

    for (;;) {
      // ---- A. housekeeping, repeated (up to max_hk passes) while too few lanes have BVH work: lanes whose
      //         queries need no traversal (e.g. rays that miss the mesh's bounding box) keep retiring paths
      //         until the warp has gathered enough rays that do ----------------------------------------
      unsigned live_mask = 0;
      for (int hk = 0;; hk++) {
        bool new_ray = false;
        if (COUNT) {
            sc_hk++;
            sc_consume += __popc(__ballot_sync(0xffffffffu, tr.node == TRT_TRAV_DONE && pend == 0 && (phase == PH_BOUNCE || phase == PH_AO)));
        }

        // ---- A1. lanes whose query is complete: consume the hit (one trip of main.c:132-239) -------
        if (tr.node == TRT_TRAV_DONE && pend == 0 && (phase == PH_BOUNCE || phase == PH_AO)) {
            PathResult res;
            if (consume_hit(A, s_sph, s_mat, pixel, s_cur, tr.t, tr.prim, tr.u, tr.v, P, phase, res, fuse_ao)) {
                my_acc[0 * PT_THREADS] += to_fixed(res.L.x, A.sample_clamp);      my_acc[1 * PT_THREADS] += to_fixed(res.L.y, A.sample_clamp);      my_acc[2 * PT_THREADS] += to_fixed(res.L.z, A.sample_clamp);
                my_acc[3 * PT_THREADS] += to_fixed(res.albedo.x, A.sample_clamp); my_acc[4 * PT_THREADS] += to_fixed(res.albedo.y, A.sample_clamp); my_acc[5 * PT_THREADS] += to_fixed(res.albedo.z, A.sample_clamp);
                my_acc[6 * PT_THREADS] += to_fixed(res.normal.x, A.sample_clamp); my_acc[7 * PT_THREADS] += to_fixed(res.normal.y, A.sample_clamp); my_acc[8 * PT_THREADS] += to_fixed(res.normal.z, A.sample_clamp);
                ray_count += res.dead_ao; skipped += res.dead_ao;
                s_cur++;
                phase = PH_NEW;
            } else {
                new_ray = true;
            }
        }

        // ---- A2. lanes without a path: flush a finished item, take the next sample or the next item ---
        if (phase == PH_NEW && s_cur >= s_end && pixel >= 0) {
#pragma unroll
            for (int c = 0; c < 9; c++) {
                const long long v = my_acc[c * PT_THREADS];
                if (v != 0) atomicAdd((unsigned long long*)(A.accum + (size_t)c * npix + pixel), (unsigned long long)v);
                my_acc[c * PT_THREADS] = 0;
            }
            pixel = -1;
        }
        {
            const bool need = (phase == PH_NEW) && (pixel < 0);
            const unsigned need_mask = __ballot_sync(0xffffffffu, need);
            if (need_mask != 0) {
                const int want = __popc(need_mask);
                const int rank = __popc(need_mask & lanemask_lt());
                const long long avail = pool_end - pool_next;
                long long fresh = 0;
                if (avail < want && !global_dry) {
                    if (lane == 0) fresh = (long long)atomicAdd(A.item_counter, (unsigned long long)A.pool);
                    fresh = __shfl_sync(0xffffffffu, fresh, 0);
                }
                long long item = -1;
                if (need) {
                    if (rank < avail) item = pool_next + rank;
                    else if (!global_dry) item = fresh + (rank - avail);
                }
                if (avail < want && !global_dry) {
                    pool_next = fresh + (want - avail);
                    pool_end = fresh + A.pool;
                    if (fresh >= A.n_items) global_dry = true;
                } else {
                    pool_next += min((long long)want, avail);
                }
                if (need) {
                    if (item >= 0 && item < A.n_items) {
                        // item -> (tile, chunk, pixel in tile)
                        int p, chunk, tile, tx, ty;
                        decode_item(A, item, p, chunk, tile, tx, ty);
                        px = tx * 8 + (p & 7); py = ty * 4 + (p >> 3);
                        if (px < W && py < H) {
                            pixel = py * W + px;
                            s_cur = chunk * A.chunk;
                            s_end = min(A.spp, s_cur + A.chunk);
                        }
                        // else: padding item of an edge tile — stays without work, asks again next trip
                    } else {
                        phase = PH_DONE;
                    }
                }
            }
        }
        live_mask = __ballot_sync(0xffffffffu, phase != PH_DONE);
        if (live_mask == 0) break;

        // ---- A3. start a path: fill_canva's per-sample set-up (main.c:265-271) --------------------
        if (phase == PH_NEW && pixel >= 0 && s_cur < s_end && A.bounces <= 0) s_cur = s_end;   // main.c:132 never iterates: all sums stay 0
        {
            // lanes that need a new path wait until start_min of them do, or no lane of the warp has a path left (api.cu: 32 on
            // sphere scenes = the warp's paths run in step, 1 on meshes = every lane re-arms on its own)
            const bool want_start = phase == PH_NEW && pixel >= 0 && s_cur < s_end;
            bool go = want_start;
            if (A.start_min > 1) {
                const unsigned ws = __ballot_sync(0xffffffffu, want_start);
                const unsigned busy = __ballot_sync(0xffffffffu, phase == PH_BOUNCE || phase == PH_AO);
                go = want_start && (__popc(ws) >= A.start_min || busy == 0);
            }
            if (go) {
                start_path(A, pixel, px, py, s_cur, P);
                phase = PH_BOUNCE;
                new_ray = true;
            }
        }

        // ---- A4. new rays: sphere scan, scene-box cull, traversal set-up ---------------------------
        if (new_ray) {
            const bool ao_ray = (phase == PH_AO);
            if (PREHIT && P.bounce == 0 && !ao_ray) {
                // a fresh path: its primary ray was traced by the packet kernel (primary_packet.cu), and counted there
                const float4 h = __ldg(A.prehit + (size_t)pixel * (size_t)A.prehit_slab + (size_t)s_cur);
                tr.t = h.x; tr.prim = __float_as_int(h.y); tr.u = h.z; tr.v = h.w;
                tr.node = TRT_TRAV_DONE; tr.sp = 0;
            } else if (!HAS_TRIS && fuse_ao) {
                const bool do_bounce = !ao_ray, do_ao = ao_ray || P.ao_pending;
                bool occluded;
                spheres_pair<COUNT>(sc, s_sph, P.o, P.d, P.d_ao, do_bounce, do_ao, P.src_prim, tr, occluded, cnt);
                ray_count += (do_bounce ? 1u : 0u) + (do_ao ? 1u : 0u);
                if (ao_ray) tr.prim = occluded ? 0 : TRT_NO_PRIM;                              // consumed by consume_hit's AO branch
                else if (do_ao) { P.T = P.T * (occluded ? A.inv_ao : 0.0f); P.ao_pending = false; }   // main.c:220-221, before the next shading
            } else {
                ray_begin<HAS_TRIS, BRUTE, COUNT>(sc, s_sph, P.o, ao_ray ? P.d_ao : P.d, P.src_prim, P.src_twin, ao_ray, tr, stack, cnt);
                ray_count++;
            }
        }
        if (!HAS_TRIS || BRUTE || hk + 1 >= A.max_hk) break;
        const int n_inside = __popc(__ballot_sync(0xffffffffu, tr.node != TRT_TRAV_DONE || pend != 0));
        const int n_ready = __popc(__ballot_sync(0xffffffffu, phase != PH_DONE && tr.node == TRT_TRAV_DONE && pend == 0));
        if (n_ready == 0 || n_inside >= ((A.hk_min * __popc(live_mask) + 31) >> 5)) break;
      }
      if (live_mask == 0) break;
      const bool ao = (phase == PH_AO);
      const float3 rd = ao ? P.d_ao : P.d;

        // ---- B. BVH traversal, left early when too few lanes are still inside ------------------------
        if (HAS_TRIS && !BRUTE) {
            const float3 idir = safe_inverse(rd);
            const float3 ood = f3(P.o.x * idir.x, P.o.y * idir.y, P.o.z * idir.z);
            const int src_leaf = P.src_prim - sc.n_spheres, twin_leaf = P.src_twin;
            const int exit_below = (A.exit_num * __popc(live_mask) + 31) >> 5;
            for (int step = 0;; step++) {
                const bool inside = tr.node != TRT_TRAV_DONE || pend != 0;
                const unsigned in_mask = __ballot_sync(0xffffffffu, inside);
                if (in_mask == 0) break;
                if (step >= A.min_steps && __popc(in_mask) < exit_below) break;
                if (COUNT) {
                    if (step == 0) sc_sessions++;
                    sc_trips++; sc_inside += __popc(in_mask);
                    sc_node += __popc(__ballot_sync(0xffffffffu, tr.node >= 0 && tr.node != TRT_TRAV_DONE));
                    sc_leaf += __popc(__ballot_sync(0xffffffffu, tr.node < 0));
                }
                // if-if: every trip, each lane does whatever it is at — a node visit or a leaf test (measured faster than
                // deferring, voting on or pairing the two on B200, profiles/README.md)
                if (DEFER) bvh_trip_deferred<COUNT, TOP, N64>(sc, P.o, rd, idir, ood, src_leaf, twin_leaf, ao, A.defer, tr, pend, stack, cnt, 0, s_top, A.top_nodes);
                else if (inside) bvh_step_unified<COUNT>(sc, P.o, rd, idir, ood, src_leaf, twin_leaf, ao, tr, stack, cnt);
            }
        }
    }

    // ---- counters: one atomic per warp ---------------------------------------------------------
    unsigned long long rc = ray_count;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) rc += __shfl_xor_sync(0xffffffffu, rc, off);
    if (lane == 0 && rc) atomicAdd(A.counters + 0, rc);
    if (A.use_ao) {
        unsigned long long sk = skipped;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sk += __shfl_xor_sync(0xffffffffu, sk, off);
        if (lane == 0 && sk) atomicAdd(A.counters + 11, sk);
    }
    if (COUNT) {
        unsigned long long b = cnt.box_tests, t = cnt.tri_tests, s = cnt.sphere_tests;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            b += __shfl_xor_sync(0xffffffffu, b, off);
            t += __shfl_xor_sync(0xffffffffu, t, off);
            s += __shfl_xor_sync(0xffffffffu, s, off);
        }
        if (lane == 0) {
            atomicAdd(A.counters + 1, b); atomicAdd(A.counters + 2, t); atomicAdd(A.counters + 3, s);
            atomicAdd(A.counters + 4, sc_trips); atomicAdd(A.counters + 5, sc_inside); atomicAdd(A.counters + 6, sc_node);
            atomicAdd(A.counters + 7, sc_leaf); atomicAdd(A.counters + 8, sc_sessions); atomicAdd(A.counters + 9, sc_hk);
            atomicAdd(A.counters + 10, sc_consume);
        }
    }
}

// ---- K4: resolve — write_color_canva (rtutility.h:56-71) + the AOV means (main.c:275-279) -------
// Sums up to n_src accumulation buffers (this GPU's and, over NVLink peer mappings, the other
// GPUs' of the ctx) while tonemapping, so a multi-GPU render needs no separate reduction pass.
__global__ void __launch_bounds__(256)
resolve_kernel(ResolveArgs R)
{
    const size_t npix = (size_t)R.W * (size_t)R.H;
    for (size_t p = R.p0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < R.p1; p += (size_t)gridDim.x * blockDim.x) {
        long long s[9];
#pragma unroll
        for (int c = 0; c < 9; c++) s[c] = 0;
        for (int k = 0; k < R.n_src; k++) {
            const long long* src = R.src[k];
#pragma unroll
            for (int c = 0; c < 9; c++) s[c] += src[(size_t)c * npix + p];
        }
        const size_t q = p - R.p0;
        const double rapport = 1.0 / (double)R.divisor;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double sum = (double)s[c] * TRT_ACCUM_INV_SCALE;
            float r = sqrtf((float)(rapport * sum));                 // sqrtf: FLOAT square root, like the reference
            double cl = (double)r;
            cl = cl < 0.0 ? 0.0 : (cl > 0.999 ? 0.999 : cl);         // NaN falls through both tests, as in clamp()
            if (R.canva8) R.canva8[3 * q + c] = (unsigned char)(int)(256.0 * cl);
            else R.canva[3 * q + c] = (double)(int)(256.0 * cl);
            R.albedo[3 * q + c] = ((double)s[3 + c] * TRT_ACCUM_INV_SCALE) / (double)R.divisor;
            R.normal[3 * q + c] = ((double)s[6 + c] * TRT_ACCUM_INV_SCALE) / (double)R.divisor;
        }
    }
}

// Un-quantised float3 means (linear radiance, albedo, normal) in OIDN_FORMAT_FLOAT3 layout: what a denoiser wants
// instead of the 8-bit round trip of denoiser.h:43-45.  The denoiser itself stays external.
__global__ void __launch_bounds__(256)
resolve_hdr_kernel(int npix, int divisor, const long long* __restrict__ acc, float* radiance, float* albedo, float* normal)
{
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < (size_t)npix; p += (size_t)gridDim.x * blockDim.x) {
        const double k = TRT_ACCUM_INV_SCALE / (double)divisor;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            radiance[3 * p + c] = (float)((double)acc[(size_t)c * npix + p] * k);
            albedo[3 * p + c] = (float)((double)acc[(size_t)(3 + c) * npix + p] * k);
            normal[3 * p + c] = (float)((double)acc[(size_t)(6 + c) * npix + p] * k);
        }
    }
}

cudaError_t trt_launch_resolve_hdr(int npix, int divisor, const long long* acc, float* radiance, float* albedo, float* normal, int n_sms, cudaStream_t stream)
{
    int grid = (npix + 255) / 256;
    if (grid > n_sms * 8) grid = n_sms * 8;
    resolve_hdr_kernel<<<grid, 256, 0, stream>>>(npix, divisor, acc, radiance, albedo, normal);
    return cudaGetLastError();
}

// ---- the path-tracing kernel's own first hit through PIXEL CENTRES -----------------------------------------------
// Parity test 1 compares the FP64 pass of primary.cu with the CPU path; this kernel answers the other half of the
// question: which primitive does the FP32 machinery K2 actually renders with (camera_ray without jitter, sphere_root,
// the BVH walk, triangle_test, first-wins ties) find for the same pixel-centre rays?  Ids in the caller's numbering.
template <bool HAS_TRIS>
__global__ void __launch_bounds__(128)
primary_fp32_kernel(const __grid_constant__ RenderArgs A, int* __restrict__ id_out, float* __restrict__ t_out)
{
    const DevScene& sc = A.scene;
    const int pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= A.W * A.H) return;
    const int px = pixel % A.W, py = pixel / A.W;
    const float u = (float)px * A.inv_wm1, v = (float)py * A.inv_hm1;                  // main.c:265-266 without the jitter
    const float3 direction = fma3(A.cam_hor, u, fma3(A.cam_ver, v, A.cam_corner));
    const float3 dest = fma3(direction, A.focus, A.cam_origin);
    const float3 o = A.cam_origin;
    const float3 d = normalize3(dest - o);
    Trav tr;
    int stack[TRT_STACK_DEPTH];
    TraceCounters cnt; cnt.box_tests = cnt.tri_tests = cnt.sphere_tests = 0;
    ray_begin<HAS_TRIS, false, false>(sc, sc.sph, o, d, TRT_NO_PRIM, -1, false, tr, stack, cnt);
    if (HAS_TRIS) {
        const float3 idir = safe_inverse(d);
        const float3 ood = f3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
        while (tr.node != TRT_TRAV_DONE) bvh_step_unified<false>(sc, o, d, idir, ood, -1, -1, false, tr, stack, cnt);
    }
    int id = -1;
    if (tr.prim != TRT_NO_PRIM) id = tr.prim < sc.n_spheres ? tr.prim : sc.n_spheres + __ldg(sc.tri_orig + (tr.prim - sc.n_spheres));
    if (id_out) id_out[pixel] = id;
    if (t_out) t_out[pixel] = tr.prim != TRT_NO_PRIM ? tr.t : 0.0f;
}

cudaError_t trt_launch_primary_fp32(const RenderArgs& A, int* id, float* t, cudaStream_t stream)
{
    const int n = A.W * A.H, grid = (n + 127) / 128;
    if (A.scene.n_tris > 0) primary_fp32_kernel<true><<<grid, 128, 0, stream>>>(A, id, t);
    else primary_fp32_kernel<false><<<grid, 128, 0, stream>>>(A, id, t);
    return cudaGetLastError();
}

// ---- launchers -------------------------------------------------------------------------------------
template <bool HAS_TRIS, bool BRUTE, bool COUNT, bool DEFER = false, bool TOP = false, bool PREHIT = false, bool N64 = false>
static cudaError_t launch_variant(const RenderArgs& A, int grid, size_t smem, cudaStream_t stream)
{
    {
        // Ask for the smallest shared-memory carve-out that still holds the resident CTAs, so that the rest of the 256 KB
        // stays L1 for BVH nodes and triangles: the driver's default picked a larger one (C2: 33.5 -> 32.5 ms with the hint).
        const int ctas = HAS_TRIS ? PT_MIN_BLOCKS_MESH : PT_MIN_BLOCKS;
        int pct = (int)((ctas * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
        const char* v = getenv("TRT_TUNE_CARVEOUT");
        if (v && *v) pct = atoi(v);
        if (pct > 100) pct = 100;
        cudaFuncSetAttribute(pathtrace_kernel<HAS_TRIS, BRUTE, COUNT, DEFER, TOP, PREHIT, N64>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        // more than ~460 spheres: the sphere table + the per-lane sums pass the 48 KB a kernel gets without asking
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(pathtrace_kernel<HAS_TRIS, BRUTE, COUNT, DEFER, TOP, PREHIT, N64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
    }
    pathtrace_kernel<HAS_TRIS, BRUTE, COUNT, DEFER, TOP, PREHIT, N64><<<grid, PT_THREADS, smem, stream>>>(A);
    return cudaGetLastError();
}

cudaError_t trt_launch_pathtrace(const RenderArgs& A, int n_sms, int blocks_per_sm, bool brute, bool count_tests, cudaStream_t stream)
{
    const bool top = A.top_nodes > 0 && A.scene.n_tris > 0 && !brute && A.defer > 0 && !count_tests;
    const size_t smem = (size_t)A.scene.n_spheres * 4 * sizeof(float4) + sizeof(long long) * 9 * PT_THREADS + (top ? (size_t)A.top_nodes * 128 + 16 : 0);
    const int grid = n_sms * (blocks_per_sm > 0 ? blocks_per_sm : (A.scene.n_tris > 0 ? PT_MIN_BLOCKS_MESH : PT_MIN_BLOCKS));
    const bool has_tris = A.scene.n_tris > 0;
    if (!has_tris) return count_tests ? launch_variant<false, false, true>(A, grid, smem, stream)
                                      : launch_variant<false, false, false>(A, grid, smem, stream);
    if (brute) return count_tests ? launch_variant<true, true, true>(A, grid, smem, stream)
                                  : launch_variant<true, true, false>(A, grid, smem, stream);
    if (A.scene.nodes64 && A.defer > 0 && !top && !A.prehit) return count_tests ? launch_variant<true, false, true, true, false, false, true>(A, grid, smem, stream)
                                                                             : launch_variant<true, false, false, true, false, false, true>(A, grid, smem, stream);
    if (top) return launch_variant<true, false, false, true, true>(A, grid, smem, stream);
    if (A.prehit && A.defer > 0) return count_tests ? launch_variant<true, false, true, true, false, true>(A, grid, smem, stream)
                                                     : launch_variant<true, false, false, true, false, true>(A, grid, smem, stream);
    if (A.defer > 0) return count_tests ? launch_variant<true, false, true, true>(A, grid, smem, stream)
                                        : launch_variant<true, false, false, true>(A, grid, smem, stream);
    return count_tests ? launch_variant<true, false, true>(A, grid, smem, stream)
                       : launch_variant<true, false, false>(A, grid, smem, stream);
}

cudaError_t trt_launch_resolve(const ResolveArgs& R, int n_sms, cudaStream_t stream)
{
    const size_t npix = R.p1 - R.p0;
    if (npix == 0) return cudaSuccess;
    int grid = (int)((npix + 255) / 256);
    if (grid > n_sms * 8) grid = n_sms * 8;
    resolve_kernel<<<grid, 256, 0, stream>>>(R);
    return cudaGetLastError();
}
