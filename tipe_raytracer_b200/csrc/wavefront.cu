// wavefront.cu — the wavefront engine of K2: the same path tracer as pathtrace.cu, split into two kernels
// that alternate ("waves") over a pool of SLOTS, one in-flight path per slot, state in global memory.
//
// Why: with a BVH the register-resident megakernel is latency-bound — ~120 registers/thread allow 16 warps
// per SM, and lanes that finish a traversal idle or drag the warp through shading with a handful of lanes
// active (ncu: 14 of 32 threads per instruction, profiles/).  Splitting gives
//   wf_trace_kernel : ONLY the closest-hit / any-hit query.  ~50 registers => 40+ warps per SM to hide the
//                     node-fetch latency; persistent warps pull slots in batches of 32 from one counter and
//                     re-arm lanes from the batch as they finish, so node visits run with most lanes active;
//   wf_shade_kernel : one thread per slot, everything else: consume the hit (tracer, main.c:132-239), finish
//                     paths into the slot's int64 sums, start the next sample (fill_canva, main.c:264-272),
//                     flush finished items with RED.ADD.64, take new items from the global item counter.
// Slots keep their identity across waves (no compaction): a slot is "dead" only when the items run out.
// Results do not depend on the schedule: Philox counters are (pixel, sample, bounce) and sums are integers.
#include "trt_shade.cuh"

#define WF_THREADS 256

namespace {

__device__ __forceinline__ int pack_meta(int phase, bool is_alpha, int bounce, int alpha_depth)
{
    return phase | ((is_alpha ? 1 : 0) << 2) | (bounce << 3) | (alpha_depth << 16);
}

__global__ void __launch_bounds__(WF_THREADS)
wf_init_kernel(WfState S)
{
    const int slot = blockIdx.x * WF_THREADS + threadIdx.x;
    if (slot >= S.n_slots) return;
    S.path_L[slot] = make_float4(0, 0, 0, __int_as_float(pack_meta(PH_NEW, false, 0, 0)));
    S.aov_a[slot] = make_float4(0, 0, 0, __int_as_float(-1));
    S.aov_n[slot] = make_float4(0, 0, 0, __int_as_float(0));
    S.dnext[slot] = make_float4(0, 0, 0, __int_as_float(0));
    S.rflag[slot] = 0;
#pragma unroll
    for (int c = 0; c < 9; c++) S.acc[(size_t)c * S.n_slots + slot] = 0;
    if (slot == 0) { *S.live = S.n_slots; *S.cursor = 0ULL; }
}

__global__ void __launch_bounds__(WF_THREADS)
wf_shade_kernel(const __grid_constant__ RenderArgs A, const __grid_constant__ WfState S)
{
    const int slot = blockIdx.x * WF_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int n = S.n_slots;
    if (slot == 0) *S.cursor = 0ULL;                       // the trace kernel of this wave starts from slot 0
    const DevScene& sc = A.scene;
    const size_t npix = (size_t)A.W * (size_t)A.H;

    const float4 pl = S.path_L[slot];
    const int meta = __float_as_int(pl.w);
    int phase = meta & 3;
    const bool alive = phase != PH_DONE;                   // n is a multiple of WF_THREADS: every thread has a slot

    PathState P;
    P.is_alpha = (meta >> 2) & 1; P.bounce = (meta >> 3) & 0x1fff; P.alpha_depth = (meta >> 16) & 0x7fff;
    P.L = xyz(pl);
    int pixel = -1, s_cur = 0, s_end = 0;
    float4 aa = make_float4(0, 0, 0, 0), an = aa, dn = aa;
    if (alive) {
        aa = S.aov_a[slot]; an = S.aov_n[slot]; dn = S.dnext[slot];
        pixel = __float_as_int(aa.w); s_cur = __float_as_int(an.w); s_end = __float_as_int(dn.w);
        P.aov_albedo = xyz(aa); P.aov_normal = xyz(an);
    }

    // ---- consume the query traced in the previous wave --------------------------------------------
    if (alive && (phase == PH_BOUNCE || phase == PH_AO)) {
        const float4 ro = S.ray_o[slot], rd = S.ray_d[slot], h = S.hit[slot], pt = S.path_T[slot];
        P.o = xyz(ro); P.src_prim = __float_as_int(ro.w); P.src_twin = __float_as_int(rd.w);
        P.T = xyz(pt); P.cur_index = pt.w;
        if (phase == PH_AO) { P.d_ao = xyz(rd); P.d = xyz(dn); } else { P.d = xyz(rd); P.d_ao = f3(0, 0, 0); }
        PathResult res;
        if (consume_hit(A, sc.sph, sc.sph_mat, pixel, s_cur, h.x, __float_as_int(h.y), h.z, h.w, P, phase, res)) {
            long long* a = S.acc + slot;
            a[0 * (size_t)n] += to_fixed(res.L.x, A.sample_clamp);      a[1 * (size_t)n] += to_fixed(res.L.y, A.sample_clamp);      a[2 * (size_t)n] += to_fixed(res.L.z, A.sample_clamp);
            a[3 * (size_t)n] += to_fixed(res.albedo.x, A.sample_clamp); a[4 * (size_t)n] += to_fixed(res.albedo.y, A.sample_clamp); a[5 * (size_t)n] += to_fixed(res.albedo.z, A.sample_clamp);
            a[6 * (size_t)n] += to_fixed(res.normal.x, A.sample_clamp); a[7 * (size_t)n] += to_fixed(res.normal.y, A.sample_clamp); a[8 * (size_t)n] += to_fixed(res.normal.z, A.sample_clamp);
            s_cur++;
            phase = PH_NEW;
        }
    }

    // ---- item finished: flush its sums, take the next item -----------------------------------------
    if (alive && phase == PH_NEW && s_cur >= s_end && pixel >= 0) {
#pragma unroll
        for (int c = 0; c < 9; c++) {
            long long* a = S.acc + (size_t)c * n + slot;
            const long long v = *a;
            if (v != 0) { atomicAdd((unsigned long long*)(A.accum + (size_t)c * npix + pixel), (unsigned long long)v); *a = 0; }
        }
        pixel = -1;
    }
    {
        const bool need = alive && phase == PH_NEW && pixel < 0;
        const unsigned need_mask = __ballot_sync(0xffffffffu, need);
        if (need_mask != 0) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(A.item_counter, (unsigned long long)__popc(need_mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (need) {
                const long long item = (long long)base + __popc(need_mask & lanemask_lt());
                if (item < A.n_items) {
                    int p, chunk, tile, tx, ty;
                    decode_item(A, item, p, chunk, tile, tx, ty);
                    const int px = tx * 8 + (p & 7), py = ty * 4 + (p >> 3);
                    if (px < A.W && py < A.H) {
                        pixel = py * A.W + px;
                        s_cur = chunk * A.chunk;
                        s_end = min(A.spp, s_cur + A.chunk);
                    }
                } else {
                    phase = PH_DONE;
                }
            }
            const unsigned died = __ballot_sync(0xffffffffu, need && phase == PH_DONE);
            if (lane == 0 && died) atomicSub(S.live, __popc(died));
        }
    }

    // ---- next sample of the item: fill_canva's set-up (main.c:265-271) -----------------------------
    if (alive && phase == PH_NEW && pixel >= 0 && s_cur < s_end) {
        if (A.bounces <= 0) s_cur = s_end;                  // main.c:132 never iterates: the sums stay 0
        else {
            start_path(A, pixel, pixel % A.W, pixel / A.W, s_cur, P);
            phase = PH_BOUNCE;
        }
    }

    // ---- write the slot back ------------------------------------------------------------------------
    if (alive) {
        const bool ao = phase == PH_AO, has_ray = ao || phase == PH_BOUNCE;
        S.rflag[slot] = has_ray ? (ao ? 2 : 1) : 0;
        S.path_L[slot] = make_float4(P.L.x, P.L.y, P.L.z, __int_as_float(pack_meta(phase, P.is_alpha, P.bounce, P.alpha_depth)));
        S.aov_a[slot] = make_float4(P.aov_albedo.x, P.aov_albedo.y, P.aov_albedo.z, __int_as_float(pixel));
        S.aov_n[slot] = make_float4(P.aov_normal.x, P.aov_normal.y, P.aov_normal.z, __int_as_float(s_cur));
        if (has_ray) {
            const float3 d = ao ? P.d_ao : P.d;
            S.ray_o[slot] = make_float4(P.o.x, P.o.y, P.o.z, __int_as_float(P.src_prim));
            S.ray_d[slot] = make_float4(d.x, d.y, d.z, __int_as_float(P.src_twin));
            S.path_T[slot] = make_float4(P.T.x, P.T.y, P.T.z, P.cur_index);
            S.dnext[slot] = make_float4(P.d.x, P.d.y, P.d.z, __int_as_float(s_end));
        } else {
            S.dnext[slot] = make_float4(0, 0, 0, __int_as_float(s_end));
        }
    }
}

// The query kernel.  Persistent: grid = SMs x resident CTAs; a warp reserves 32 consecutive slots at a time.
template <bool HAS_TRIS, bool COUNT>
__global__ void __launch_bounds__(WF_THREADS, 4)
wf_trace_kernel(const __grid_constant__ RenderArgs A, const __grid_constant__ WfState S)
{
    extern __shared__ float4 smem[];
    float4* s_sph = smem;
    for (int k = threadIdx.x; k < A.scene.n_spheres; k += blockDim.x) s_sph[k] = A.scene.sph[k];
    __syncthreads();

    const DevScene& sc = A.scene;
    const int lane = threadIdx.x & 31;
    const int n = S.n_slots;
    int pool_next = 0, pool_end = 0;        // warp-uniform
    bool dry = false;                       // warp-uniform: the cursor ran past the last slot
    int slot = -1;
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1);
    int src_prim = TRT_NO_PRIM, src_twin = -1;
    bool any_hit = false;
    Trav tr; tr.node = TRT_TRAV_DONE; tr.sp = 0; tr.t = 0.0f; tr.prim = TRT_NO_PRIM; tr.u = tr.v = 0.0f;
    int stack[TRT_STACK_DEPTH];
    unsigned ray_count = 0;
    TraceCounters cnt; cnt.box_tests = cnt.tri_tests = cnt.sphere_tests = 0;

    for (;;) {
        // ---- finished queries out, next slots in -------------------------------------------------
        if (slot >= 0 && tr.node == TRT_TRAV_DONE) {
            S.hit[slot] = make_float4(tr.t, __int_as_float(tr.prim), tr.u, tr.v);
            slot = -1;
        }
        const bool need = slot < 0;
        const unsigned need_mask = __ballot_sync(0xffffffffu, need);
        if (need_mask == 0xffffffffu && dry && pool_next >= pool_end) break;     // nothing in flight, nothing left
        if (need_mask != 0 && !(dry && pool_next >= pool_end)) {
            const int want = __popc(need_mask);
            const int rank = __popc(need_mask & lanemask_lt());
            const int avail = pool_end - pool_next;
            int fresh = 0;
            const bool refill = avail < want && !dry;
            if (refill) {
                unsigned long long f = 0;
                if (lane == 0) f = atomicAdd(S.cursor, 32ULL);
                f = __shfl_sync(0xffffffffu, f, 0);
                fresh = f >= (unsigned long long)n ? n : (int)f;
            }
            int idx = -1;
            if (need) {
                if (rank < avail) idx = pool_next + rank;
                else if (refill && fresh < n) idx = fresh + (rank - avail);
            }
            if (refill) {
                if (fresh >= n) { dry = true; pool_next = pool_end; }
                else { pool_next = fresh + (want - avail); pool_end = fresh + 32; }
            } else {
                pool_next += min(want, avail);
            }
            if (idx >= 0 && idx < n) {
                const int flag = S.rflag[idx];
                if (flag != 0) {
                    const float4 ro = S.ray_o[idx], rd = S.ray_d[idx];
                    o = xyz(ro); d = xyz(rd);
                    src_prim = __float_as_int(ro.w); src_twin = __float_as_int(rd.w);
                    any_hit = flag == 2;
                    slot = idx;
                    ray_begin<HAS_TRIS, false, COUNT>(sc, s_sph, o, d, src_prim, src_twin, any_hit, tr, stack, cnt);
                    ray_count++;
                }
            }
        }

        // ---- BVH traversal; left early (to re-arm finished lanes) while slots remain ------------------
        if (HAS_TRIS) {
            const float3 idir = safe_inverse(d);
            const float3 ood = f3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
            const int src_leaf = src_prim - sc.n_spheres;
            const bool more = !(dry && pool_next >= pool_end);
            for (int step = 0;; step++) {
                const bool inside = slot >= 0 && (tr.node != TRT_TRAV_DONE);
                const unsigned in_mask = __ballot_sync(0xffffffffu, inside);
                if (in_mask == 0) break;
                if (more && step >= A.min_steps && __popc(in_mask) < A.exit_num) break;
                if (inside) bvh_step_unified<COUNT>(sc, o, d, idir, ood, src_leaf, src_twin, any_hit, tr, stack, cnt);
            }
        }
    }

    unsigned long long rc = ray_count;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) rc += __shfl_xor_sync(0xffffffffu, rc, off);
    if (lane == 0 && rc) atomicAdd(A.counters + 0, rc);
    if (COUNT) {
        unsigned long long b = cnt.box_tests, t = cnt.tri_tests, s = cnt.sphere_tests;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            b += __shfl_xor_sync(0xffffffffu, b, off);
            t += __shfl_xor_sync(0xffffffffu, t, off);
            s += __shfl_xor_sync(0xffffffffu, s, off);
        }
        if (lane == 0) { atomicAdd(A.counters + 1, b); atomicAdd(A.counters + 2, t); atomicAdd(A.counters + 3, s); }
    }
}

inline size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

size_t trt_wavefront_bytes(int n_slots)
{
    const size_t n = (size_t)n_slots;
    return 8 * a256(sizeof(float4) * n) + a256(sizeof(long long) * 9 * n) + a256(n) + 256;
}

void trt_wavefront_carve(void* base, int n_slots, WfState& S)
{
    char* p = (char*)base;
    const size_t n = (size_t)n_slots, f4 = a256(sizeof(float4) * n);
    S.n_slots = n_slots;
    S.ray_o = (float4*)p; p += f4;  S.ray_d = (float4*)p; p += f4;  S.hit = (float4*)p; p += f4;   S.path_T = (float4*)p; p += f4;
    S.path_L = (float4*)p; p += f4; S.aov_a = (float4*)p; p += f4;  S.aov_n = (float4*)p; p += f4; S.dnext = (float4*)p; p += f4;
    S.acc = (long long*)p; p += a256(sizeof(long long) * 9 * n);
    S.rflag = (unsigned char*)p; p += a256(n);
    S.live = (int*)p; S.cursor = (unsigned long long*)(p + 64);
}

cudaError_t trt_launch_wavefront(const RenderArgs& A, WfState& S, int n_sms, int* h_live_pinned, int* launches, cudaStream_t stream)
{
    const int blocks = S.n_slots / WF_THREADS;
    const bool count = A.count_tests != 0, has_tris = A.scene.n_tris > 0;
    const size_t smem = sizeof(float4) * (size_t)A.scene.n_spheres;
    const int trace_grid = n_sms * (A.wf_blocks > 0 ? A.wf_blocks : 4);
    int n_launch = 0;
    wf_init_kernel<<<blocks, WF_THREADS, 0, stream>>>(S);
    n_launch++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int waves_per_check = A.wf_check > 0 ? A.wf_check : 8;
    for (long long wave = 0;; wave++) {
        wf_shade_kernel<<<blocks, WF_THREADS, 0, stream>>>(A, S);
        if (has_tris) {
            if (count) wf_trace_kernel<true, true><<<trace_grid, WF_THREADS, smem, stream>>>(A, S);
            else wf_trace_kernel<true, false><<<trace_grid, WF_THREADS, smem, stream>>>(A, S);
        } else {
            if (count) wf_trace_kernel<false, true><<<trace_grid, WF_THREADS, smem, stream>>>(A, S);
            else wf_trace_kernel<false, false><<<trace_grid, WF_THREADS, smem, stream>>>(A, S);
        }
        n_launch += 2;
        if ((wave + 1) % waves_per_check == 0) {
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = cudaMemcpyAsync(h_live_pinned, S.live, sizeof(int), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
            if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
            if (*h_live_pinned <= 0) break;
        }
        if (wave > (1LL << 40)) return cudaErrorUnknown;
    }
    if (launches) *launches = n_launch;
    return cudaSuccess;
}
