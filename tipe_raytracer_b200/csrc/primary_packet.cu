// primary_packet.cu — K2a: the primary rays of a frame traced as PACKETS, ahead of the path-tracing kernel (engine 4).
//
// The samples of a pixel differ only by the sub-pixel jitter (main.c:265-266) and the aperture offset (main.c:268-269):
// 32 consecutive samples of one pixel are 32 nearly identical rays.  In the path-tracing kernel each lane walks the BVH with
// its own stack; bounce rays need that, primary rays do not.  Here a warp takes 32 consecutive paths (pixel-major, sample-
// fastest: one pixel, or the end of one and the start of its neighbour), generates their camera rays exactly like K2
// (same Philox draw, same arithmetic), scans the spheres per lane, and walks the BVH ONCE for the whole warp:
//   * one stack per warp in shared memory; every node is fetched with a warp-uniform address (one L1 wavefront instead
//     of up to 32) and tested by all lanes against their own ray and their own best distance;
//   * a child is visited if ANY lane's slab test passes (ballot); children are ordered by the smallest entry distance in
//     the warp (redux.min on the float bits); a leaf's triangles are tested by all lanes;
//   * every lane keeps its own closest hit, with the same acceptance rule and tie-break as K2, so the result per ray is
//     the one K2 would have found (order of traversal does not matter for a closest-hit query).
// The hits (t, primitive, u, v) go to a buffer indexed (pixel, sample); K2 then starts each path from its stored hit
// instead of tracing the primary ray (pathtrace.cu, PREHIT).  Rays are counted here; K2 does not count them again.
#include "trt_shade.cuh"

#define PK_THREADS 256
#define PK_WARPS (PK_THREADS / 32)
#define PK_STACK 96

template <bool COUNT>
__global__ void __launch_bounds__(PK_THREADS, 4)
primary_packet_kernel(const __grid_constant__ RenderArgs A, float4* __restrict__ hits, int slab)
{
    extern __shared__ float4 smem[];
    TRT_DEBUG_INIT(A);
    const DevScene& sc = A.scene;
    const int ns = sc.n_spheres;
    float4* s_sph = smem;                                                        // [n_spheres]
    int* s_stack = (int*)(smem + ns) + (threadIdx.x >> 5) * PK_STACK;            // [PK_WARPS][PK_STACK]
    for (int k = threadIdx.x; k < ns; k += blockDim.x) s_sph[k] = sc.sph[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const float INF = __int_as_float(0x7f800000);
    unsigned ray_count = 0;
    TraceCounters cnt; cnt.box_tests = cnt.tri_tests = cnt.sphere_tests = 0;
    const long long n_packets = (A.n_items + 31) >> 5;

    for (;;) {
        long long first = 0;
        if (lane == 0) first = (long long)atomicAdd(A.item_counter, (unsigned long long)A.pool);
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= n_packets) break;
        const long long last = min(first + A.pool, n_packets);
        for (long long pk = first; pk < last; pk++) {
            const long long g = pk * 32 + lane;
            int px = 0, py = 0, sample = 0;
            bool active = g < A.n_items;
            if (active) { decode_path(A, g, px, py, sample); active = px < A.W && py < A.H; }
            const int pixel = py * A.W + px;
            float3 o = A.cam_origin, d = f3(0.0f, 0.0f, 1.0f);
            Trav tr; tr.t = INF; tr.prim = TRT_NO_PRIM; tr.u = tr.v = 0.0f; tr.node = TRT_TRAV_DONE; tr.sp = 0;
            bool in = false;                                                    // this lane's ray has to walk the BVH
            if (active) {
                camera_ray(A, pixel, px, py, sample, o, d);
                in = ray_scan<true, false, COUNT>(sc, s_sph, o, d, TRT_NO_PRIM, -1, false, tr, cnt);
                ray_count++;
            }
            if (__ballot_sync(0xffffffffu, in) != 0) {
                const float3 idir = safe_inverse(d);
                const float3 ood = f3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
                int sp = 0;                                                     // warp-uniform
                int node = sc.root;
                for (;;) {
                    TRT_CHECK(sp >= 0 && sp + 3 <= PK_STACK && (node < 0 || node < sc.n_nodes), 301);
                    if (node >= 0) {
                        // ---- inner node: uniform fetch, four slab tests per lane ---------------------------------
                        const float4* rec = sc.nodes + 8 * (size_t)node;
                        const int rot = node & 3;
                        float4 f0a, f0b, f1a, f1b, f2a, f2b, f3a, f3b;
                        ldg256(rec + 2 * ((0 + rot) & 3), f0a, f0b); ldg256(rec + 2 * ((1 + rot) & 3), f1a, f1b);
                        ldg256(rec + 2 * ((2 + rot) & 3), f2a, f2b); ldg256(rec + 2 * ((3 + rot) & 3), f3a, f3b);
#define PK_SLAB(LX, HX, LY, HY, LZ, HZ, REF, KEY)                                                         \
    {                                                                                                      \
        const float ax = fmaf(LX, idir.x, -ood.x), bx = fmaf(HX, idir.x, -ood.x);                          \
        const float ay = fmaf(LY, idir.y, -ood.y), by = fmaf(HY, idir.y, -ood.y);                          \
        const float az = fmaf(LZ, idir.z, -ood.z), bz = fmaf(HZ, idir.z, -ood.z);                          \
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), sc.box_tmin));    \
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tr.t));           \
        KEY = (in && tf >= tn && REF != TRT_CHILD_EMPTY) ? tn : INF;                                       \
        if (COUNT && in && REF != TRT_CHILD_EMPTY) cnt.box_tests++;                                        \
    }
                        float k0, k1, k2, k3;
                        int r0 = __float_as_int(f3a.x), r1 = __float_as_int(f3a.y), r2 = __float_as_int(f3a.z), r3 = __float_as_int(f3a.w);
                        PK_SLAB(f0a.x, f0b.x, f1a.x, f1b.x, f2a.x, f2b.x, r0, k0)
                        PK_SLAB(f0a.y, f0b.y, f1a.y, f1b.y, f2a.y, f2b.y, r1, k1)
                        PK_SLAB(f0a.z, f0b.z, f1a.z, f1b.z, f2a.z, f2b.z, r2, k2)
                        PK_SLAB(f0a.w, f0b.w, f1a.w, f1b.w, f2a.w, f2b.w, r3, k3)
#undef PK_SLAB
                        // entry distances are >= box_tmin >= 0: their bit patterns order like the floats; INF = nobody hits the child
                        float m0 = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(k0)));
                        float m1 = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(k1)));
                        float m2 = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(k2)));
                        float m3 = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(k3)));
                        order2(m0, r0, m1, r1); order2(m2, r2, m3, r3);
                        order2(m0, r0, m2, r2); order2(m1, r1, m3, r3);
                        order2(m1, r1, m2, r2);
                        if (lane == 0) {
                            int q = sp;
                            if (m3 < INF) s_stack[q++] = r3;
                            if (m2 < INF) s_stack[q++] = r2;
                            if (m1 < INF) s_stack[q++] = r1;
                        }
                        sp += (m3 < INF ? 1 : 0) + (m2 < INF ? 1 : 0) + (m1 < INF ? 1 : 0);
                        __syncwarp();
                        if (m0 < INF) node = r0;
                        else { if (sp == 0) break; node = s_stack[--sp]; }
                    } else {
                        // ---- leaf: the run's triangles, uniform fetches, every lane tests its own ray -----------------
                        int p = ~node & TRT_LEAF_POS_MASK;
                        const int run_len = ((~node) >> TRT_LEAF_POS_BITS) + 1;
                        TRT_CHECK(p >= 0 && p + run_len <= sc.n_tris, 302);
                        for (int k = 0; k < run_len; k++, p++) {
                            const float4* rec = sc.tri + 4 * (size_t)p;
                            float4 f0a, f0b, f1a, f1b;
                            ldg256(rec, f0a, f0b); ldg256(rec + 2, f1a, f1b);
                            float t, u, v;
                            if (COUNT && in) cnt.tri_tests++;
                            if (in && triangle_test(f0a, f0b, f1a, o, d, sc.tri_tmin, t, u, v)) {
                                bool take = t < tr.t;
                                if (t == tr.t && tr.prim >= ns) take = __float_as_int(f1b.x) < __ldg(sc.tri_orig + (tr.prim - ns));
                                if (take) { tr.t = t; tr.prim = ns + p; tr.u = u; tr.v = v; }
                            }
                        }
                        if (sp == 0) break;
                        node = s_stack[--sp];
                    }
                }
                __syncwarp();
            }
            if (active) hits[(size_t)pixel * (size_t)slab + (size_t)sample] = make_float4(tr.t, __int_as_float(tr.prim), tr.u, tr.v);
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

// A.n_items = paths (tiles * 32 * spp, padding included), A.pool = packets a warp reserves per atomic, A.item_counter zeroed.
// hits[pixel * slab + sample], sample in [0, A.spp), slab >= A.spp.
cudaError_t trt_launch_primary_packets(const RenderArgs& A, int n_sms, float4* hits, int slab, bool count_tests, cudaStream_t stream)
{
    const size_t smem = sizeof(float4) * (size_t)A.scene.n_spheres + sizeof(int) * PK_WARPS * PK_STACK;
    if (smem > 48 * 1024) return cudaErrorInvalidValue;              // more than ~2 800 spheres: the caller falls back to engine 1
    const int grid = n_sms * 4;
    if (count_tests) primary_packet_kernel<true><<<grid, PK_THREADS, smem, stream>>>(A, hits, slab);
    else primary_packet_kernel<false><<<grid, PK_THREADS, smem, stream>>>(A, hits, slab);
    return cudaGetLastError();
}
