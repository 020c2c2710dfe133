// pathtrace_pool.cu — K2p, the path-tracing kernel for triangle meshes: a persistent megakernel whose warps keep a
// POOL OF PATH SLOTS in shared memory and compact rays at warp level.
//
// Same per-path arithmetic as pathtrace.cu (trt_shade.cuh / trt_isect.cuh: fill_canva + tracer + closest_hit +
// ambient_occlusion, main.c:245-284, 118-242, 52-92, 94-116), different execution model.  In the register-resident
// kernel a lane owns one path; when its ray ends it idles until the whole warp leaves the BVH loop for a housekeeping
// pass in which every lane is somewhere else in tracer() (ncu, C2: 17 of 32 threads per instruction).  Here:
//
//   * a warp owns N path SLOTS (default 64) whose state lives in shared memory (96 B per slot, SoA float4 so that
//     lanes reading different slots do not collide on banks) and two ring buffers of slot numbers: RAY-READY (the
//     slot holds a ray that has to walk the BVH) and HIT-READY (its closest_hit query is complete);
//   * TRAVERSAL: the 32 lanes walk the BVH for 32 slots.  A lane whose ray ends writes the hit into its slot (4
//     stores), pushes it on the hit-ready ring and takes the next ray-ready slot (3 LDS.128): ~40 instructions to
//     re-arm instead of a ~900-instruction housekeeping pass, so the loop stays full.  The loop only keeps what a
//     query needs in registers (ray, inverse direction, Trav); the path state is not live in it;
//   * SHADING: once 32 slots are hit-ready (or nothing else can run) the warp runs ONE trip of tracer()'s bounce loop
//     for 32 slots together — every lane is at the same place of main.c:132-239 — starts new paths in the slots
//     whose path ended, scans the spheres and culls against the scene box for the new rays, and sorts the slots into
//     the two rings again;
//   * work = one global counter over paths numbered (tile, pixel in tile, sample): a warp reserves a few hundred
//     consecutive paths per atomic, so its slots work on the samples of one or two pixels (coherent primary rays);
//   * a finished path adds its radiance to the pixel's int64 sums with RED.ADD.64; the albedo / normal AOVs are
//     added at the hit that fixes them (main.c:137-150), so they are not carried in the slot.  Integer adds commute:
//     the image is bit-identical to pathtrace.cu's, for any schedule, grid, sample split or GPU count.
#include "trt_shade.cuh"

#define PP_THREADS 128
#define PP_WARPS (PP_THREADS / 32)
#define PP_RING 64            // ring capacity = maximum number of slots per warp
#define PP_RING_QUADS 9       // 2 rings x 64 bytes + 16 bytes of per-warp scalars, in float4 units

__device__ __forceinline__ unsigned long long pp_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Slot record, 6 (7 with ambient occlusion) float4 per slot, stored as q[quad][slot]:
//   q0 (ray origin, hit distance)              q1 (direction of the ray to trace, bits: hit primitive)
//   q2 bits(src_prim, src_twin, any-hit, pixel) q3 (rayColor T, current refractive index)
//   q4 (incomingLight L, bits: meta)            q5 (hit u, hit v, bits: sample, -)
//   q6 (bounce direction parked during an ambient-occlusion query)
// meta = phase | is_alpha << 2 | bounce << 3 | alpha_depth << 16
template <bool COUNT, bool AO_PARK, int MIN_BLOCKS>
__global__ void __launch_bounds__(PP_THREADS, MIN_BLOCKS)
pathtrace_pool_kernel(const __grid_constant__ RenderArgs A)
{
    extern __shared__ float4 smem[];
    TRT_DEBUG_INIT(A);
    const DevScene& sc = A.scene;
    const int ns = sc.n_spheres;
    float4* s_sph = smem;                                   // [n_spheres]
    float4* s_mat = smem + ns;                              // [3*n_spheres]
    for (int k = threadIdx.x; k < ns; k += blockDim.x) s_sph[k] = sc.sph[k];
    for (int k = threadIdx.x; k < 3 * ns; k += blockDim.x) s_mat[k] = sc.sph_mat[k];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = A.pp_slots;
    constexpr int NQ = AO_PARK ? 7 : 6;
    float4* wq = smem + 4 * ns + warp * (NQ * N + PP_RING_QUADS);
    unsigned char* rr_ring = (unsigned char*)(wq + NQ * N);
    unsigned char* hr_ring = rr_ring + PP_RING;
    const size_t npix = (size_t)A.W * (size_t)A.H;

    // every slot starts hit-ready without a path
    for (int s = lane; s < N; s += 32) {
        wq[4 * N + s] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(PH_NEW));
        hr_ring[s] = (unsigned char)s;
    }
    __syncthreads();

    // warp-uniform bookkeeping
    int rr_head = 0, rr_count = 0, hr_head = 0, hr_count = N;
    long long pool_next = 0, pool_end = 0;
    bool global_dry = false;

    // lane state of the traversal loop
    int slot = -1;
    float3 ro = f3(0, 0, 0), rd = f3(0, 0, 0), idir = f3(0, 0, 0), ood = f3(0, 0, 0);
    int src_leaf = -1, twin_leaf = -1;
    bool ao = false;
    Trav tr; tr.node = TRT_TRAV_DONE; tr.sp = 0; tr.t = 0.0f; tr.prim = TRT_NO_PRIM; tr.u = tr.v = 0.0f;
    int pend = 0;
    int stack[TRT_STACK_DEPTH];

    unsigned ray_count = 0, skipped = 0;
    TraceCounters cnt; cnt.box_tests = cnt.tri_tests = cnt.sphere_tests = 0;
    unsigned long long sc_trips = 0, sc_inside = 0, sc_node = 0, sc_leaf = 0, sc_sessions = 0, sc_hk = 0, sc_consume = 0;
    unsigned long long* t_start = (unsigned long long*)(hr_ring + PP_RING);
    if (lane == 0) *t_start = pp_globaltimer();
    unsigned wd = 0;

// the ray of slot S -> the lane's traversal registers (not Trav: t / prim of a query under way live in registers)
#define PP_LOAD_RAY(S)                                                                                           \
    {                                                                                                            \
        const float4 l0 = wq[S], l1 = wq[N + (S)], l2 = wq[2 * N + (S)];                                         \
        ro = xyz(l0); rd = xyz(l1);                                                                              \
        src_leaf = __float_as_int(l2.x) - ns; twin_leaf = __float_as_int(l2.y); ao = __float_as_int(l2.z) != 0;  \
        idir = safe_inverse(rd);                                                                                 \
        ood = f3(ro.x * idir.x, ro.y * idir.y, ro.z * idir.z);                                                   \
    }
// lanes whose query is complete: hit -> slot, slot -> hit-ready ring
#define PP_FLUSH_DONE()                                                                                          \
    {                                                                                                            \
        const bool fin = slot >= 0 && tr.node == TRT_TRAV_DONE && pend == 0;                                     \
        const unsigned m_fin = __ballot_sync(0xffffffffu, fin);                                                  \
        if (fin) {                                                                                               \
            ((float*)(wq + slot))[3] = tr.t;                                                                     \
            ((int*)(wq + N + slot))[3] = tr.prim;                                                                \
            *(float2*)(wq + 5 * N + slot) = make_float2(tr.u, tr.v);                                             \
            hr_ring[(hr_head + hr_count + __popc(m_fin & lanemask_lt())) & (PP_RING - 1)] = (unsigned char)slot; \
            slot = -1;                                                                                           \
        }                                                                                                        \
        hr_count += __popc(m_fin);                                                                               \
    }

    for (;;) {
        // ---- A. shading: one trip of the bounce loop for up to 32 hit-ready slots ----------------------------
        {
            const int n_inside = __popc(__ballot_sync(0xffffffffu, slot >= 0));
            while (hr_count > 0 && (hr_count >= A.pp_shade_at || rr_count + n_inside < A.pp_exit_below)) {
                __syncwarp();
                const int n = min(32, hr_count);
                const bool mine = lane < n;
                const int s = mine ? (int)hr_ring[(hr_head + lane) & (PP_RING - 1)] : 0;
                TRT_CHECK(s >= 0 && s < N && hr_count >= 0 && rr_count >= 0 && hr_count + rr_count <= N, 201);
                hr_head = (hr_head + n) & (PP_RING - 1); hr_count -= n;

                PathState P;
                P.aov_albedo = P.aov_normal = f3(0, 0, 0);
                P.ao_pending = false;
                int phase = PH_DONE, pixel = 0, sample = 0, hit_prim = TRT_NO_PRIM;
                float hit_t = 0.0f, hit_u = 0.0f, hit_v = 0.0f;
                if (mine) {
                    const float4 q0 = wq[s], q1 = wq[N + s], q2 = wq[2 * N + s], q3 = wq[3 * N + s], q4 = wq[4 * N + s], q5 = wq[5 * N + s];
                    const int meta = __float_as_int(q4.w);
                    phase = meta & 3;
                    P.is_alpha = ((meta >> 2) & 1) != 0; P.bounce = (meta >> 3) & 0x1fff; P.alpha_depth = (meta >> 16) & 0xffff;
                    P.o = xyz(q0); hit_t = q0.w;
                    P.d = xyz(q1); hit_prim = __float_as_int(q1.w);
                    P.d_ao = P.d;
                    if (AO_PARK) { if (phase == PH_AO) P.d = xyz(wq[6 * N + s]); }
                    P.src_prim = __float_as_int(q2.x); P.src_twin = __float_as_int(q2.y); pixel = __float_as_int(q2.w);
                    P.T = xyz(q3); P.cur_index = q3.w;
                    P.L = xyz(q4);
                    hit_u = q5.x; hit_v = q5.y; sample = __float_as_int(q5.z);
                }
                if (COUNT) { sc_hk++; sc_consume += __popc(__ballot_sync(0xffffffffu, phase == PH_BOUNCE || phase == PH_AO)); }

                // A1. consume the hit
                bool new_ray = false;
                if (phase == PH_BOUNCE || phase == PH_AO) {
                    // the AOVs are set at bounce 0 and again at the first hit after a run of texture holes (main.c:137-150);
                    // a value set now is final unless this hit is itself a hole and the path goes on (then the next hit sets
                    // it again), so it is added to the sums here instead of being carried in the slot
                    const bool aov_set_now = phase == PH_BOUNCE && (P.bounce == 0 || (P.bounce == P.alpha_depth && P.is_alpha));
                    PathResult res;
                    const bool ended = consume_hit(A, s_sph, s_mat, pixel, sample, hit_t, hit_prim, hit_u, hit_v, P, phase, res, false);
                    const bool aov_final = aov_set_now && (ended || !P.is_alpha);
                    if (ended || aov_final) {
                        const float3 eL = ended ? res.L : f3(0, 0, 0);
                        const float3 ea = ended ? res.albedo : P.aov_albedo, en = ended ? res.normal : P.aov_normal;
                        const float v[9] = {eL.x, eL.y, eL.z, ea.x, ea.y, ea.z, en.x, en.y, en.z};
#pragma unroll
                        for (int c = 0; c < 9; c++) {
                            const long long f = to_fixed(v[c], A.sample_clamp);
                            if (f != 0 && !(A.exp_flags & 4)) atomicAdd((unsigned long long*)(A.accum + (size_t)c * npix + pixel), (unsigned long long)f);
                        }
                    }
                    if (ended) { phase = PH_NEW; ray_count += res.dead_ao; skipped += res.dead_ao; } else new_ray = true;
                }

                // A2. slots without a path take the next one of the warp's reservation
                {
                    const bool need = mine && phase == PH_NEW;
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
                        long long g = -1;
                        if (need) {
                            if (rank < avail) g = pool_next + rank;
                            else if (!global_dry) g = fresh + (rank - avail);
                        }
                        if (avail < want && !global_dry) {
                            pool_next = fresh + (want - avail);
                            pool_end = fresh + A.pool;
                            if (fresh >= A.n_items) global_dry = true;
                        } else {
                            pool_next += min((long long)want, avail);
                        }
                        if (need) {
                            if (g >= 0 && g < A.n_items) {
                                int px, py;
                                decode_path(A, g, px, py, sample);
                                if (px < A.W && py < A.H) {
                                    pixel = py * A.W + px;
                                    start_path(A, pixel, px, py, sample, P);
                                    phase = PH_BOUNCE;
                                    new_ray = true;
                                }
                                // else: padding pixel of an edge tile, the slot asks again in the next pass
                            } else {
                                phase = PH_DONE;                                   // no work left: the slot retires
                            }
                        }
                    }
                }

                // A3. new rays: sphere scan + scene-box cull; only rays that have to walk the BVH go to the ray-ready ring
                bool needs_bvh = false;
                Trav tn; tn.t = 0.0f; tn.prim = TRT_NO_PRIM; tn.u = tn.v = 0.0f;
                const bool ao_ray = phase == PH_AO;
                if (new_ray) {
                    needs_bvh = ray_scan<true, false, COUNT>(sc, s_sph, P.o, ao_ray ? P.d_ao : P.d, P.src_prim, P.src_twin, ao_ray, tn, cnt);
                    ray_count++;
                }

                // A4. state back to the slot, slot into its ring
                const bool alive = mine && phase != PH_DONE;
                if (alive) {
                    const float3 dir = ao_ray ? P.d_ao : P.d;
                    const int meta = phase | (P.is_alpha ? 4 : 0) | (P.bounce << 3) | (P.alpha_depth << 16);
                    wq[s] = make_float4(P.o.x, P.o.y, P.o.z, tn.t);
                    wq[N + s] = make_float4(dir.x, dir.y, dir.z, __int_as_float(tn.prim));
                    wq[2 * N + s] = make_float4(__int_as_float(P.src_prim), __int_as_float(P.src_twin), __int_as_float(ao_ray ? 1 : 0), __int_as_float(pixel));
                    wq[3 * N + s] = make_float4(P.T.x, P.T.y, P.T.z, P.cur_index);
                    wq[4 * N + s] = make_float4(P.L.x, P.L.y, P.L.z, __int_as_float(meta));
                    wq[5 * N + s] = make_float4(0.0f, 0.0f, __int_as_float(sample), 0.0f);
                    if (AO_PARK) { if (ao_ray) wq[6 * N + s] = make_float4(P.d.x, P.d.y, P.d.z, 0.0f); }
                }
                const bool to_rr = alive && needs_bvh, to_hr = alive && !needs_bvh;
                const unsigned m_rr = __ballot_sync(0xffffffffu, to_rr), m_hr = __ballot_sync(0xffffffffu, to_hr);
                if (to_rr) rr_ring[(rr_head + rr_count + __popc(m_rr & lanemask_lt())) & (PP_RING - 1)] = (unsigned char)s;
                if (to_hr) hr_ring[(hr_head + hr_count + __popc(m_hr & lanemask_lt())) & (PP_RING - 1)] = (unsigned char)s;
                rr_count += __popc(m_rr); hr_count += __popc(m_hr);
                __syncwarp();
            }
            if (rr_count == 0 && n_inside == 0) break;          // hr_count is 0 here: every slot has retired
        }

        // ---- B. traversal: lanes re-arm from the ray-ready ring as their rays end --------------------------------
        // Between sessions a lane in the middle of a query keeps only (slot, Trav, parked leaf) in registers; the ray itself is
        // re-read from the slot, so that the shading passes do not carry it (registers: 6 CTAs of 128 threads per SM).
        if (COUNT) sc_sessions++;
        __syncwarp();
        if (slot >= 0) PP_LOAD_RAY(slot)
        int n_empty = 32 - __popc(__ballot_sync(0xffffffffu, slot >= 0));
        for (int step = 0;; step++) {
            // busy = in the middle of a query; a lane whose query is complete WAITS with its result until enough lanes are
            // free for one service pass (results to the slots, slots to the hit-ready ring, next rays from the ray-ready ring)
            int n_busy = __popc(__ballot_sync(0xffffffffu, slot >= 0 && !(tr.node == TRT_TRAV_DONE && pend == 0)));
            int n_wait = 32 - n_busy - n_empty;
            if ((32 - n_busy >= A.pp_refill_min || n_busy == 0 || step == 0) && (n_wait > 0 || (rr_count > 0 && n_empty > 0))) {
                PP_FLUSH_DONE()
                const unsigned need_mask = __ballot_sync(0xffffffffu, slot < 0);
                n_empty = __popc(need_mask);
                if (rr_count > 0) {
                    __syncwarp();
                    const int rank = __popc(need_mask & lanemask_lt());
                    if (slot < 0 && rank < rr_count) {
                        slot = (int)rr_ring[(rr_head + rank) & (PP_RING - 1)];
                        TRT_CHECK(slot >= 0 && slot < N, 202);
                        PP_LOAD_RAY(slot)
                        const float4 q0 = wq[slot], q1 = wq[N + slot];
                        tr.t = q0.w; tr.prim = __float_as_int(q1.w); tr.u = tr.v = 0.0f;
                        stack[0] = TRT_TRAV_DONE; tr.sp = 1; tr.node = sc.root;
                        pend = 0;
                    }
                    const int k = min(n_empty, rr_count);
                    rr_head = (rr_head + k) & (PP_RING - 1); rr_count -= k; n_empty -= k;
                }
                n_busy = 32 - n_empty; n_wait = 0;
            }
            if (n_busy == 0) break;
            if (hr_count + n_wait >= A.pp_shade_at) break;                                   // a full shading pass is waiting
            if (rr_count == 0 && hr_count + n_wait > 0 && step >= A.min_steps && n_busy < A.pp_exit_below) break;
            if (COUNT) {
                sc_trips++; sc_inside += n_busy;
                sc_node += __popc(__ballot_sync(0xffffffffu, tr.node >= 0 && tr.node != TRT_TRAV_DONE));
                sc_leaf += __popc(__ballot_sync(0xffffffffu, tr.node < 0));
            }
            bvh_trip_deferred<COUNT>(sc, ro, rd, idir, ood, src_leaf, twin_leaf, ao, A.defer, tr, pend, stack, cnt, A.exp_flags);
            if ((++wd & 0xfff) == 0 && pp_globaltimer() - *t_start > 20000000000ULL) {      // 20 s: a hung schedule, not a long frame
                if (lane == 0) atomicExch(A.counters + 13, 1ULL);
                hr_count = rr_count = 0; slot = -1; global_dry = true;
                break;
            }
        }
        PP_FLUSH_DONE()
    }
#undef PP_LOAD_RAY
#undef PP_FLUSH_DONE

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

// ---- launcher ------------------------------------------------------------------------------------------
size_t trt_pool_smem_bytes(const RenderArgs& A)
{
    const int nq = A.use_ao ? 7 : 6;
    return sizeof(float4) * ((size_t)4 * A.scene.n_spheres + (size_t)PP_WARPS * ((size_t)nq * A.pp_slots + PP_RING_QUADS));
}

template <bool COUNT, bool AO_PARK, int MIN_BLOCKS>
static cudaError_t launch_pool(const RenderArgs& A, int n_sms, size_t smem, cudaStream_t stream)
{
    auto kernel = pathtrace_pool_kernel<COUNT, AO_PARK, MIN_BLOCKS>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // the smallest carve-out that holds the resident CTAs: the rest of the 256 KB stays L1 for BVH nodes and triangles
    int pct = (int)((MIN_BLOCKS * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    const char* v = getenv("TRT_TUNE_CARVEOUT");
    if (v && *v) pct = atoi(v);
    if (pct > 100) pct = 100;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    kernel<<<n_sms * MIN_BLOCKS, PP_THREADS, smem, stream>>>(A);
    return cudaGetLastError();
}

// blocks_per_sm: 4, 5 (default) or 6 CTAs of 128 threads per SM (128 / 96 / 80 registers per thread)
cudaError_t trt_launch_pathtrace_pool(const RenderArgs& A, int n_sms, int blocks_per_sm, bool count_tests, cudaStream_t stream)
{
    const size_t smem = trt_pool_smem_bytes(A);
    const bool ao = A.use_ao != 0;
#define PP_DISPATCH(MB)                                                                                      \
    return count_tests ? (ao ? launch_pool<true, true, MB>(A, n_sms, smem, stream) : launch_pool<true, false, MB>(A, n_sms, smem, stream)) \
                       : (ao ? launch_pool<false, true, MB>(A, n_sms, smem, stream) : launch_pool<false, false, MB>(A, n_sms, smem, stream));
    if (blocks_per_sm == 4) { PP_DISPATCH(4) }
    if (blocks_per_sm == 6) { PP_DISPATCH(6) }
    PP_DISPATCH(5)
#undef PP_DISPATCH
}
