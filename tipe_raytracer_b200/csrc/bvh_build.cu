// bvh_build.cu — K3: scene conversion (FP64 AoS -> device layout) and the GPU LBVH build.
//
// The reference has no acceleration structure (closest_hit scans every triangle, main.c:81-90);
// this is new.  Input = the caller's arrays exactly as main.c holds them (sphere[] 112 B,
// triangle[] 200 B, material[] 80 B per texel), already copied to the device.  Output = the
// DevScene layout of trt_device.cuh.
//
//   1. tri_bounds      per-triangle FP32 box (rounded outward) + centroid; scene bounds by atomics
//   2. morton_keys     36-bit Morton code of the centroid << 28 | triangle index  (unique 64-bit keys)
//   3. radix sort      cub::DeviceRadixSort (library; the build is not on the per-ray path)
//   4. emit_leaves     triangles re-laid out in sorted (leaf) order: A, AB, AC, N + ids as 4 float4 (64 B)
//   4b. group_marks/heads  triangles with IDENTICAL boxes (the two halves of a quad, the back faces of a double-faced mesh, crossed
//                      plant quads: 2, 4 or 8 triangles per box in the reference's Minecraft exports) are neighbours after
//                      the sort and become ONE leaf: the hierarchy is built over the distinct boxes, and a leaf step tests
//                      the whole run (count in the head record).  Without this the tree has to separate equal boxes with
//                      inner nodes every ray through the box visits in full.
//   5. find_twins      coincident opposite-facing triangles ("doubleface" meshes) share a Morton code and
//                      are neighbours after the sort; each records the other's leaf position
//   6. hierarchy       PLOC (Meister & Bittner 2018), the default: clusters in Morton order repeatedly merge with
//                      their mutual nearest neighbour (surface area of the union, search radius 64) — close to
//                      SAH-sweep quality, ~30 small launches.  TRT_BVH=lbvh selects the Karras 2012 LBVH instead
//                      (one thread per internal node, longest-common-prefix ranges, then a bottom-up refit)
//   8. sah_cost        surface-area cost of the binary tree (reported through trt_bvh_info)
//   9. collapse_level  binary -> 4-wide: each surviving node adopts its grandchildren, largest box first,
//                      until it has 4 children; one launch per level of the wide tree.  128-byte nodes
//                      (one L1 line): 4 child boxes as 6 float4 (SoA) + int4 child references
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cstdlib>

#include "trt_kernels.h"

namespace {

constexpr int IDX_BITS = 28;                        // up to 2^28 triangles
constexpr unsigned long long IDX_MASK = (1ULL << IDX_BITS) - 1ULL;

struct MatOverrideDev { int ordinal; unsigned mask; double ec[3]; double es, rs, alpha, index; };
static_assert(sizeof(MatOverrideDev) == 64, "mirror of trt_mat_override");

// ---- ordered-int encoding so float min/max can use integer atomics ------------------------------
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void convert_spheres(const double* __restrict__ raw, int n, float4* sph, float4* mat)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double* s = raw + 14 * (size_t)k;      // center[3], radius, diffuse[3], emission[3], es, rs, alpha, index
    sph[k] = make_float4((float)s[0], (float)s[1], (float)s[2], (float)s[3]);
    mat[3 * k + 0] = make_float4((float)s[4], (float)s[5], (float)s[6], (float)s[12]);
    mat[3 * k + 1] = make_float4((float)s[7], (float)s[8], (float)s[9], (float)s[10]);
    mat[3 * k + 2] = make_float4((float)s[11], (float)s[13], 0.0f, 0.0f);
}

// material[] slab (packed to 10 floats per texel on the host) -> texel records, with the per-ordinal overrides of texture.h:71-87 baked in.
__global__ void convert_texels(const float* __restrict__ raw, size_t n_texels, int texels_per_mat,
                               const MatOverrideDev* __restrict__ ov, int n_ov, float4* out)
{
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_texels; t += (size_t)gridDim.x * blockDim.x) {
        const float* m = raw + 10 * t;
        float dr = m[0], dg = m[1], db = m[2], er = m[3], eg = m[4], eb = m[5], es = m[6], rs = m[7], al = m[8], ix = m[9];
        const int ordinal = (int)(t / (size_t)texels_per_mat);
        for (int k = 0; k < n_ov; k++) {
            if (ov[k].ordinal != ordinal) continue;
            const unsigned mask = ov[k].mask;
            if (mask & 1u) { er = (float)ov[k].ec[0]; eg = (float)ov[k].ec[1]; eb = (float)ov[k].ec[2]; }
            if (mask & 2u) es = (float)ov[k].es;
            if (mask & 4u) rs = (float)ov[k].rs;
            if (mask & 8u) al = (float)ov[k].alpha;
            if (mask & 16u) ix = (float)ov[k].index;
        }
        out[3 * t + 0] = make_float4(dr, dg, db, al);
        out[3 * t + 1] = make_float4(er, eg, eb, es);
        out[3 * t + 2] = make_float4(rs, ix, 0.0f, 0.0f);
    }
}

// bounds[0..2] = ordered-int min, [3..5] = ordered-int max, [6] = ordered-int max |coord|
__global__ void tri_bounds(const double* __restrict__ raw, int n, float4* box_lo, float4* box_hi, int* bounds)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    float cen[3] = {0, 0, 0};
    if (k < n) {
        const double* t = raw + 25 * (size_t)k;
        for (int a = 0; a < 3; a++) {
            const double mn = fmin(t[a], fmin(t[3 + a], t[6 + a])), mx = fmax(t[a], fmax(t[3 + a], t[6 + a]));
            lo[a] = __double2float_rd(mn); hi[a] = __double2float_ru(mx);
            cen[a] = (float)((t[a] + t[3 + a] + t[6 + a]) * (1.0 / 3.0));
        }
        box_lo[k] = make_float4(lo[0], lo[1], lo[2], cen[0]);
        box_hi[k] = make_float4(hi[0], hi[1], hi[2], cen[1]);
    }
    // warp-reduce then one atomic per warp
    float amax = fmaxf(fmaxf(fmaxf(fabsf(lo[0]), fabsf(hi[0])), fmaxf(fabsf(lo[1]), fabsf(hi[1]))), fmaxf(fabsf(lo[2]), fabsf(hi[2])));
    if (k >= n) amax = 0.0f;
    for (int off = 16; off > 0; off >>= 1) {
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
        }
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
    }
    if ((threadIdx.x & 31) == 0) {
        for (int a = 0; a < 3; a++) { atomicMin(bounds + a, f2ord(lo[a])); atomicMax(bounds + 3 + a, f2ord(hi[a])); }
        atomicMax(bounds + 6, f2ord(amax));
    }
}

__device__ __forceinline__ unsigned long long spread3(unsigned v)     // bit k of v -> bit 3k
{
    unsigned long long x = v & 0x1fffffu;
    x = (x | (x << 32)) & 0x001f00000000ffffULL;
    x = (x | (x << 16)) & 0x001f0000ff0000ffULL;
    x = (x | (x << 8)) & 0x100f00f00f00f00fULL;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ULL;
    x = (x | (x << 2)) & 0x1249249249249249ULL;
    return x;
}

__global__ void morton_keys(int n, const float4* __restrict__ box_lo, const float4* __restrict__ box_hi,
                            const int* __restrict__ bounds, unsigned long long* keys)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 lo = box_lo[k], hi = box_hi[k];
    const float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);   // box centre: equal boxes get equal codes
    float q[3];
    const float c[3] = {cx, cy, cz};
    for (int a = 0; a < 3; a++) {
        const float mn = ord2f(bounds[a]), mx = ord2f(bounds[3 + a]);
        const float ext = fmaxf(mx - mn, 1e-30f);
        q[a] = fminf(fmaxf((c[a] - mn) / ext, 0.0f), 1.0f) * 4095.0f;
    }
    const unsigned long long code = (spread3((unsigned)q[0]) << 2) | (spread3((unsigned)q[1]) << 1) | spread3((unsigned)q[2]);
    keys[k] = (code << IDX_BITS) | (unsigned long long)k;
}

// Triangles in leaf order.  N = AB x AC is formed in FP64 and rounded once.
__global__ void emit_leaves(int n, const unsigned long long* __restrict__ keys, const double* __restrict__ raw,
                            const int* __restrict__ quel_mat, const float4* __restrict__ box_lo, const float4* __restrict__ box_hi,
                            const int* __restrict__ bounds, float4* tri, float4* tri_uv, int* tri_orig, float4* leaf_lo, float4* leaf_hi)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float pad = __fadd_rn(__fmul_rn(ord2f(bounds[6]), 4.76837158e-7f), 1e-30f);          // 2^-21 of the largest |coordinate| (host: trt_build_scene)
    const int k = (int)(keys[p] & IDX_MASK);
    const double* t = raw + 25 * (size_t)k;
    const double ab[3] = {t[3] - t[0], t[4] - t[1], t[5] - t[2]}, ac[3] = {t[6] - t[0], t[7] - t[1], t[8] - t[2]};
    const double nx = ab[1] * ac[2] - ab[2] * ac[1], ny = ab[2] * ac[0] - ab[0] * ac[2], nz = ab[0] * ac[1] - ab[1] * ac[0];
    tri[4 * p + 0] = make_float4((float)t[0], (float)t[1], (float)t[2], (float)ab[0]);
    tri[4 * p + 1] = make_float4((float)ab[1], (float)ab[2], (float)ac[0], (float)ac[1]);
    tri[4 * p + 2] = make_float4((float)ac[2], (float)nx, (float)ny, (float)nz);
    tri[4 * p + 3] = make_float4(__int_as_float(k), __int_as_float(-1), __int_as_float(quel_mat[k]), 0.0f);   // caller index, twin, ordinal
    tri_uv[2 * p + 0] = make_float4((float)t[19], (float)t[20], (float)t[21], (float)t[22]);
    tri_uv[2 * p + 1] = make_float4((float)t[23], (float)t[24], __int_as_float(quel_mat[k]), __int_as_float(-1));
    tri_orig[p] = k;
    const float4 lo = box_lo[k], hi = box_hi[k];
    leaf_lo[p] = make_float4(lo.x - pad, lo.y - pad, lo.z - pad, 0.0f);
    leaf_hi[p] = make_float4(hi.x + pad, hi.y + pad, hi.z + pad, 0.0f);
}

// Runs of equal boxes, cut every `cap` triangles.  run_mark[p] = p when the box at sorted position p differs from its
// predecessor's (or p == 0, or grouping is off), else 0; an inclusive max-scan turns that into the start of p's run, and
// head[p] = 1 where (p - start) is a multiple of cap: a run of 40 identical boxes becomes leaves of 16 + 16 + 8.
__global__ void group_marks(int n, const unsigned long long* __restrict__ keys, const float4* __restrict__ box_lo,
                            const float4* __restrict__ box_hi, int group, int* run_mark)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int differs = 1;
    if (p > 0 && group) {
        const int k = (int)(keys[p] & IDX_MASK), k2 = (int)(keys[p - 1] & IDX_MASK);
        const float4 a = box_lo[k], b = box_hi[k], c = box_lo[k2], d = box_hi[k2];
        differs = !(a.x == c.x && a.y == c.y && a.z == c.z && b.x == d.x && b.y == d.y && b.z == d.z);
    }
    run_mark[p] = differs ? p : 0;
}

__global__ void group_heads(int n, const int* __restrict__ run_start, int cap, int* head)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) head[p] = ((p - run_start[p]) % cap == 0) ? 1 : 0;
}

__global__ void group_first(int n, const int* __restrict__ head, const int* __restrict__ gpos, int* gfirst)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n && head[p]) gfirst[gpos[p]] = p;
}

// the head record of every leaf carries the number of triangles of the run
__global__ void group_counts(int n, const int* __restrict__ ghead, const int* __restrict__ gpos, const int* __restrict__ gfirst, float4* tri)
{
    const int n_groups = gpos[n - 1] + ghead[n - 1];
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int first = gfirst[g], next = (g + 1 < n_groups) ? gfirst[g + 1] : n;
    tri[4 * first + 3].w = __int_as_float(next - first);
}

__device__ __forceinline__ bool same_point(const double* a, const double* b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }

// A "twin" has the same three vertices and the opposite winding.  In exact arithmetic a ray leaving
// triangle S from its own surface meets the twin at distance 0 < 1e-7 and the reference rejects it
// (mesh.h:88); in FP32 that distance is noise, so the twin is skipped by id like S itself.
__global__ void find_twins(int n, const unsigned long long* __restrict__ keys, const double* __restrict__ raw,
                           float4* tri_uv, float4* tri, int* n_twins)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const unsigned long long code = keys[p] >> IDX_BITS;
    const double* a = raw + 25 * (size_t)(keys[p] & IDX_MASK);
    int twin = -1;
    for (int dir = -1; dir <= 1 && twin < 0; dir += 2) {
        for (int s = 1; s <= 16; s++) {
            const int q = p + dir * s;
            if (q < 0 || q >= n || (keys[q] >> IDX_BITS) != code) break;
            const double* b = raw + 25 * (size_t)(keys[q] & IDX_MASK);
            // same vertex set, any rotation of the reversed order
            const bool rev = (same_point(a, b) && same_point(a + 3, b + 6) && same_point(a + 6, b + 3)) ||
                             (same_point(a, b + 3) && same_point(a + 3, b) && same_point(a + 6, b + 6)) ||
                             (same_point(a, b + 6) && same_point(a + 3, b + 3) && same_point(a + 6, b));
            if (rev) { twin = q; break; }
        }
    }
    if (twin >= 0) {
        tri_uv[2 * p + 1].w = __int_as_float(twin);
        tri[4 * p + 3].y = __int_as_float(twin);
        atomicAdd(n_twins, 1);
    }
}

// ---- Karras 2012 -------------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const unsigned long long* keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    return __clzll((long long)(keys[i] ^ keys[j]));     // keys are unique, so the xor is never 0
}

// children / parents: internal node c >= 0, leaf p as ~p
__global__ void build_hierarchy(int n, const unsigned long long* __restrict__ keys, int2* children, int* parent_node, int* parent_leaf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = (lo == gamma) ? ~gamma : gamma;
    const int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    if (left < 0) parent_leaf[~left] = i; else parent_node[left] = i;
    if (right < 0) parent_leaf[~right] = i; else parent_node[right] = i;
    if (i == 0) parent_node[0] = -1;
}

__global__ void refit(int n, const int2* __restrict__ children, const int* __restrict__ parent_node, const int* __restrict__ parent_leaf,
                      const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi,
                      float4* node_lo, float4* node_hi, int* flags, int* max_depth)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int node = parent_leaf[p];
    int depth = 1;
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(flags + node, 1) == 0) break;          // first arriver stops; the second one has both children ready
        const int2 ch = children[node];
        // sibling boxes were written by another thread: read them from L2, not a stale L1 line
        const float4 alo = ch.x < 0 ? leaf_lo[~ch.x] : __ldcg(node_lo + ch.x);
        const float4 ahi = ch.x < 0 ? leaf_hi[~ch.x] : __ldcg(node_hi + ch.x);
        const float4 blo = ch.y < 0 ? leaf_lo[~ch.y] : __ldcg(node_lo + ch.y);
        const float4 bhi = ch.y < 0 ? leaf_hi[~ch.y] : __ldcg(node_hi + ch.y);
        node_lo[node] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
        node_hi[node] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
        node = parent_node[node];
        depth++;
    }
    if (node < 0) atomicMax(max_depth, depth);               // this thread reached the root
}

__device__ __forceinline__ float half_area(float4 lo, float4 hi)
{
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// ---- PLOC ----------------------------------------------------------------------------------------------
// Clusters live in Morton order in (cid, clo, chi): cid < 0 is leaf ~p, cid >= 0 an internal node.
// The cluster count lives on the device (d_c): several passes are launched back to back without a host round trip,
// each with a grid for the last count the host saw; threads beyond the current count leave at once.
__global__ void ploc_nearest(const int* __restrict__ d_c, const float4* __restrict__ clo, const float4* __restrict__ chi, int radius, int* nn)
{
    const int c = *d_c;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const float4 lo = clo[i], hi = chi[i];
    float best = 3.0e38f; int bj = -1;
    const int j0 = max(0, i - radius), j1 = min(c - 1, i + radius);
    for (int j = j0; j <= j1; j++) {
        if (j == i) continue;
        const float4 l2 = clo[j], h2 = chi[j];
        const float dx = fmaxf(hi.x, h2.x) - fminf(lo.x, l2.x), dy = fmaxf(hi.y, h2.y) - fminf(lo.y, l2.y), dz = fmaxf(hi.z, h2.z) - fminf(lo.z, l2.z);
        const float a = dx * dy + dy * dz + dz * dx;
        if (a < best) { best = a; bj = j; }          // ties: lowest j, so "mutual" is symmetric and deterministic
    }
    nn[i] = bj;
}

// Mutual nearest neighbours merge into a new internal node (owned by the lower index); flag[i] = cluster i survives.
__global__ void ploc_merge(const int* __restrict__ d_c, int c_bound, int* cid, float4* clo, float4* chi, const int* __restrict__ nn, int* node_counter,
                           int2* children, float4* node_lo, float4* node_hi, int* flag)
{
    const int c = *d_c;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) { if (i < c_bound) flag[i] = 0; return; }       // the scan runs over c_bound entries
    const int j = nn[i];
    int keep = 1;
    if (j >= 0 && nn[j] == i) {
        if (i < j) {
            const int node = atomicAdd(node_counter, 1);
            const float4 lo = clo[i], hi = chi[i], l2 = clo[j], h2 = chi[j];
            const float4 mlo = make_float4(fminf(lo.x, l2.x), fminf(lo.y, l2.y), fminf(lo.z, l2.z), 0.0f);
            const float4 mhi = make_float4(fmaxf(hi.x, h2.x), fmaxf(hi.y, h2.y), fmaxf(hi.z, h2.z), 0.0f);
            children[node] = make_int2(cid[i], cid[j]);
            node_lo[node] = mlo; node_hi[node] = mhi;
            cid[i] = node; clo[i] = mlo; chi[i] = mhi;      // nobody else reads slot i's box in this kernel's remainder: j only checks nn
        } else keep = 0;
    }
    flag[i] = keep;
}

__global__ void ploc_compact(const int* __restrict__ d_c, int* d_c_next, const int* __restrict__ cid, const float4* __restrict__ clo, const float4* __restrict__ chi,
                             const int* __restrict__ flag, const int* __restrict__ pos, int* cid2, float4* clo2, float4* chi2)
{
    const int c = *d_c;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    if (i == c - 1) *d_c_next = pos[i] + flag[i];              // survivors = count of the next pass
    if (!flag[i]) return;
    const int q = pos[i];
    cid2[q] = cid[i]; clo2[q] = clo[i]; chi2[q] = chi[i];
}

__global__ void ploc_init(int n_groups, const int* __restrict__ gfirst, const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi,
                          int* cid, float4* clo, float4* chi)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int p = gfirst[g];                 // a leaf is referenced by the position of its first triangle
    cid[g] = ~p; clo[g] = leaf_lo[p]; chi[g] = leaf_hi[p];
}

// the same with the group count read on the device (small scenes: the host does not know it yet); d_cnt[0] = clusters, node counter = 0
__global__ void ploc_init_dev(int n, const int* __restrict__ ghead, const int* __restrict__ gpos, const int* __restrict__ gfirst,
                              const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi, int* cid, float4* clo, float4* chi,
                              int* d_cnt, int* node_counter)
{
    const int n_groups = gpos[n - 1] + ghead[n - 1];
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) { d_cnt[0] = n_groups; d_cnt[1] = n_groups; *node_counter = 0; }
    if (g >= n_groups) return;
    const int p = gfirst[g];
    cid[g] = ~p; clo[g] = leaf_lo[p]; chi[g] = leaf_hi[p];
}

__global__ void sah_cost_kernel(int n, const int2* __restrict__ children, const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi,
                                const float4* __restrict__ node_lo, const float4* __restrict__ node_hi, int root, double* sah)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double cost = 0.0;
    if (i < n - 1) {
        const int2 ch = children[i];
        // SAH with C_box = 1 per child box test and C_tri = 1 per leaf, relative to the root area
        const float root_area = fmaxf(half_area(node_lo[root], node_hi[root]), 1e-30f);
        cost = (double)(half_area(node_lo[i], node_hi[i]) / root_area) * 2.0;
        if (ch.x < 0) cost += (double)(half_area(leaf_lo[~ch.x], leaf_hi[~ch.x]) / root_area);
        if (ch.y < 0) cost += (double)(half_area(leaf_lo[~ch.y], leaf_hi[~ch.y]) / root_area);
    }
    for (int off = 16; off > 0; off >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, off);
    if ((threadIdx.x & 31) == 0 && cost != 0.0) atomicAdd(sah, cost);
}

// One entry of the binary -> 4-wide collapse: e = (binary node, wide node index to fill).  The node adopts its grandchildren,
// largest box first, until it has 4 children.  counters[0] = wide nodes allocated so far, counters[1] = entries written to q_out
// (global or shared memory).  children / node_lo / node_hi carry no __restrict__: build_small_kernel writes them earlier in the same
// kernel, and a non-coherent (LDG.CONSTANT) read would be allowed to miss those writes.
__device__ __forceinline__ void collapse_entry(int2 e, const int2* children, const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi,
                                               const float4* node_lo, const float4* node_hi, int2* q_out, int* counters, float4* nodes4,
                                               const float4* __restrict__ tri)
{
    int c[4]; int nc = 2;
    { const int2 ch = children[e.x]; c[0] = ch.x; c[1] = ch.y; c[2] = c[3] = 0; }
    while (nc < 4) {
        int best = -1; float best_area = -1.0f;
        for (int j = 0; j < nc; j++)
            if (c[j] >= 0) { const float a = half_area(node_lo[c[j]], node_hi[c[j]]); if (a > best_area) { best_area = a; best = j; } }
        if (best < 0) break;
        const int2 ch = children[c[best]];
        c[best] = ch.x; c[nc++] = ch.y;
    }
    float lo[3][4], hi[3][4]; int ref[4];
    for (int j = 0; j < 4; j++) {
        if (j < nc) {
            const float4 l = c[j] < 0 ? leaf_lo[~c[j]] : node_lo[c[j]], h = c[j] < 0 ? leaf_hi[~c[j]] : node_hi[c[j]];
            lo[0][j] = l.x; lo[1][j] = l.y; lo[2][j] = l.z; hi[0][j] = h.x; hi[1][j] = h.y; hi[2][j] = h.z;
            if (c[j] < 0) ref[j] = TRT_LEAF_REF(~c[j], __float_as_int(tri[4 * (size_t)(~c[j]) + 3].w));      // run length from the head record (group_counts)
            else {
                const int idx = atomicAdd(counters + 0, 1);
                ref[j] = idx;
                q_out[atomicAdd(counters + 1, 1)] = make_int2(c[j], idx);
            }
        } else {
            for (int a = 0; a < 3; a++) { lo[a][j] = 0.0f; hi[a][j] = 0.0f; }
            ref[j] = 0x7ffffffe;                                  // TRT_CHILD_EMPTY
        }
    }
    // The four 32-byte fields of a node are stored ROTATED by (node index & 3): lanes of a warp fetch the same
    // field of different nodes at the same time, and without the rotation they would all hit the same L1 data
    // banks (tools/microbench/gather.cu: 72 -> 212 G records/s with it).
    float4* out = nodes4 + 8 * (size_t)e.y;
    const int rot = e.y & 3;
    float4* f0 = out + 2 * ((0 + rot) & 3); float4* f1 = out + 2 * ((1 + rot) & 3);
    float4* f2 = out + 2 * ((2 + rot) & 3); float4* f3 = out + 2 * ((3 + rot) & 3);
    f0[0] = make_float4(lo[0][0], lo[0][1], lo[0][2], lo[0][3]); f0[1] = make_float4(hi[0][0], hi[0][1], hi[0][2], hi[0][3]);
    f1[0] = make_float4(lo[1][0], lo[1][1], lo[1][2], lo[1][3]); f1[1] = make_float4(hi[1][0], hi[1][1], hi[1][2], hi[1][3]);
    f2[0] = make_float4(lo[2][0], lo[2][1], lo[2][2], lo[2][3]); f2[1] = make_float4(hi[2][0], hi[2][1], hi[2][2], hi[2][3]);
    f3[0] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), __int_as_float(ref[2]), __int_as_float(ref[3]));
    f3[1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// One level of the collapse, one launch per level (large scenes).
__global__ void collapse_level(const int2* __restrict__ children, const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi,
                               const float4* __restrict__ node_lo, const float4* __restrict__ node_hi,
                               const int2* __restrict__ q_in, int n_in, int2* q_out, int* counters, float4* nodes4, const float4* __restrict__ tri)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_in) return;
    collapse_entry(q_in[k], children, leaf_lo, leaf_hi, node_lo, node_hi, q_out, counters, nodes4, tri);
}

// ---- small scenes: hierarchy, SAH report and collapse in ONE CTA, no host round trip ---------------------------------
// Up to SMALL_MAX_GROUPS leaves (every mesh of the reference: 2 035 on C2, 895 on C4).  PLOC exactly as above (same nearest /
// merge / compact passes, same tie rules), but the clusters live in shared memory, the passes are separated by __syncthreads()
// instead of kernel boundaries, and the loop ends on the device.  The multi-kernel path needs ~100 launches and ~10 host
// read-backs for C2 (1.5 ms of a 1.7 ms trt_scene_set); this one is a single launch.
constexpr int SMALL_MAX_GROUPS = 4096;
constexpr int SMALL_THREADS = 1024;
struct SmallResult { int root, n_wide, wide_depth, n_groups; double sah; };

__global__ void __launch_bounds__(SMALL_THREADS, 1)
build_small_kernel(int n, const int* __restrict__ ghead, const int* __restrict__ gpos,
                   const int* g_cid, const float4* g_clo, const float4* g_chi, const int* d_count, const int* node_counter,
                   const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi, int radius,
                   int2* children, float4* node_lo, float4* node_hi, int2* queue_a, int2* queue_b, float4* nodes4, const float4* __restrict__ tri,
                   SmallResult* result)
{
    extern __shared__ float4 sm4[];
    const int n_groups = gpos[n - 1] + ghead[n - 1];
    const int c0 = *d_count;                                // clusters left by the multi-block passes that ran before this kernel
    if (c0 > SMALL_MAX_GROUPS) { if (threadIdx.x == 0) result->n_groups = -c0; return; }     // host falls back to the multi-kernel path
    float4* clo = sm4;                                      // [SMALL_MAX_GROUPS] cluster boxes, in Morton order
    float4* chi = clo + SMALL_MAX_GROUPS;
    int* cid = (int*)(chi + SMALL_MAX_GROUPS);              // leaf ~p or internal node
    int* nn = cid + SMALL_MAX_GROUPS;
    __shared__ int s_warp[32];
    __shared__ int s_count, s_nodes, s_counters[2];
    __shared__ double s_sah[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int g = tid; g < c0; g += SMALL_THREADS) { cid[g] = g_cid[g]; clo[g] = g_clo[g]; chi[g] = g_chi[g]; }
    if (tid == 0) { s_count = c0; s_nodes = *node_counter; }
    __syncthreads();
    int root = ~0;
    if (n_groups > 1) {
        int c = c0;
        while (c > 1) {
            // nearest neighbour within the search radius (ties: lowest index)
            for (int i = tid; i < c; i += SMALL_THREADS) {
                const float4 lo = clo[i], hi = chi[i];
                float best = 3.0e38f; int bj = -1;
                const int j0 = max(0, i - radius), j1 = min(c - 1, i + radius);
                for (int j = j0; j <= j1; j++) {
                    if (j == i) continue;
                    const float4 l2 = clo[j], h2 = chi[j];
                    const float dx = fmaxf(hi.x, h2.x) - fminf(lo.x, l2.x), dy = fmaxf(hi.y, h2.y) - fminf(lo.y, l2.y), dz = fmaxf(hi.z, h2.z) - fminf(lo.z, l2.z);
                    const float a = dx * dy + dy * dz + dz * dx;
                    if (a < best) { best = a; bj = j; }
                }
                nn[i] = bj;
            }
            __syncthreads();
            // merge mutual pairs (the lower index keeps the merged cluster), then compact: every thread owns a contiguous chunk
            const int per = (c + SMALL_THREADS - 1) / SMALL_THREADS;          // <= 4
            const int first = tid * per;
            int k_id[4]; float4 k_lo[4], k_hi[4]; int kept = 0;
            for (int q = 0; q < per; q++) {
                const int i = first + q;
                if (i >= c) break;
                const int j = nn[i];
                int id = cid[i]; float4 lo = clo[i], hi = chi[i];
                bool keep = true;
                if (j >= 0 && nn[j] == i) {
                    if (i < j) {
                        const int node = atomicAdd(&s_nodes, 1);
                        const float4 l2 = clo[j], h2 = chi[j];
                        lo = make_float4(fminf(lo.x, l2.x), fminf(lo.y, l2.y), fminf(lo.z, l2.z), 0.0f);
                        hi = make_float4(fmaxf(hi.x, h2.x), fmaxf(hi.y, h2.y), fmaxf(hi.z, h2.z), 0.0f);
                        children[node] = make_int2(id, cid[j]);
                        node_lo[node] = lo; node_hi[node] = hi;
                        id = node;
                    } else keep = false;
                }
                if (keep) { k_id[kept] = id; k_lo[kept] = lo; k_hi[kept] = hi; kept++; }
            }
            // exclusive scan of `kept` over the block
            int incl = kept;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += y; }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();                                  // also: every thread has read its chunk (and its partners' boxes) by now
            if (warp == 0) {
                int w = s_warp[lane];
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, off); if (lane >= off) w += y; }
                s_warp[lane] = w;
                if (lane == 31) s_count = w;
            }
            __syncthreads();
            const int base = (warp > 0 ? s_warp[warp - 1] : 0) + incl - kept;
            for (int q = 0; q < kept; q++) { cid[base + q] = k_id[q]; clo[base + q] = k_lo[q]; chi[base + q] = k_hi[q]; }
            c = s_count;
            __syncthreads();
        }
        root = cid[0];
    }
    __syncthreads();
    const int n_inner = n_groups - 1;

    // SAH report (same formula as sah_cost_kernel)
    double cost = 0.0;
    if (n_inner > 0) {
        const float root_area = fmaxf(half_area(node_lo[root], node_hi[root]), 1e-30f);
        for (int i = tid; i < n_inner; i += SMALL_THREADS) {
            const int2 ch = children[i];
            cost += (double)(half_area(node_lo[i], node_hi[i]) / root_area) * 2.0;
            if (ch.x < 0) cost += (double)(half_area(leaf_lo[~ch.x], leaf_hi[~ch.x]) / root_area);
            if (ch.y < 0) cost += (double)(half_area(leaf_lo[~ch.y], leaf_hi[~ch.y]) / root_area);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, off);
    if (lane == 0) s_sah[warp] = cost;

    // binary -> 4-wide, level by level; the queues ping-pong in global scratch
    int n_in = 0, level = 0;
    if (tid == 0) { s_counters[0] = 1; s_counters[1] = 0; if (n_inner > 0) queue_a[0] = make_int2(root, 0); }
    __syncthreads();
    n_in = n_inner > 0 ? 1 : 0;
    while (n_in > 0) {
        const int2* q_in = (level & 1) ? queue_b : queue_a;
        int2* q_out = (level & 1) ? queue_a : queue_b;
        for (int k = tid; k < n_in; k += SMALL_THREADS) collapse_entry(q_in[k], children, leaf_lo, leaf_hi, node_lo, node_hi, q_out, s_counters, nodes4, tri);
        __syncthreads();
        n_in = s_counters[1];
        __syncthreads();
        if (tid == 0) s_counters[1] = 0;
        __syncthreads();
        level++;
    }
    if (tid == 0) {
        double total = 0.0;
        for (int w = 0; w < SMALL_THREADS / 32; w++) total += s_sah[w];
        result->root = n_inner > 0 ? 0 : TRT_LEAF_REF(0, n);               // a single run: the root is that leaf
        result->n_wide = n_inner > 0 ? s_counters[0] : 0;
        result->wide_depth = level;
        result->n_groups = n_groups;
        result->sah = total;
    }
}

// ---- 64-byte nodes (experiment, TRT_TUNE_NODE64): the four child boxes quantised to 8 bits per plane relative to the node's own box
//   word 0-3 : origin.xyz (float), exponent bytes ex | ey << 8 | ez << 16 (scale_a = 2^(e_a - 127))
//   word 4-7 : bytes lo.x[4] | hi.x[4] | lo.y[4] | hi.y[4]        word 8-9 : lo.z[4] | hi.z[4]        word 12-15 : child references
// Planes are rounded outward by one extra step (the traversal folds the byte -> float conversion into the slab FMA, which costs
// up to half a step of accuracy), so the quantised box always contains the exact one: it can only add node visits, never lose a hit.
__global__ void compress_nodes(int n_nodes, const float4* __restrict__ nodes128, uint4* __restrict__ nodes64)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const float4* in = nodes128 + 8 * (size_t)i;
    const int rot = i & 3;
    float lo[3][4], hi[3][4]; int ref[4];
    for (int a = 0; a < 3; a++) {
        const float4 l = in[2 * ((a + rot) & 3)], h = in[2 * ((a + rot) & 3) + 1];
        lo[a][0] = l.x; lo[a][1] = l.y; lo[a][2] = l.z; lo[a][3] = l.w; hi[a][0] = h.x; hi[a][1] = h.y; hi[a][2] = h.z; hi[a][3] = h.w;
    }
    { const float4 r = in[2 * ((3 + rot) & 3)]; ref[0] = __float_as_int(r.x); ref[1] = __float_as_int(r.y); ref[2] = __float_as_int(r.z); ref[3] = __float_as_int(r.w); }
    float org[3]; unsigned ex[3]; unsigned qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0};
    for (int a = 0; a < 3; a++) {
        float mn = 3.0e38f, mx = -3.0e38f;
        for (int j = 0; j < 4; j++) if (ref[j] != 0x7ffffffe) { mn = fminf(mn, lo[a][j]); mx = fmaxf(mx, hi[a][j]); }
        org[a] = mn;
        // smallest power of two with (mx - mn) / scale <= 253: steps 1..254 hold the exact range, 0 and 255 the outward step
        const float need = fmaxf((mx - mn) * (1.0f / 253.0f), 1e-30f);
        int e = ((__float_as_int(need) >> 23) & 0xff) + 1;                   // 2^(e-127) > need
        if (e < 1) e = 1; if (e > 254) e = 254;
        ex[a] = (unsigned)e;
        const float inv = __int_as_float((254 - e) << 23);                  // 2^-(e-127)
        for (int j = 0; j < 4; j++) {
            int l = 255, h = 0;                                              // unused slot (never read: its reference is EMPTY)
            if (ref[j] != 0x7ffffffe) {
                l = (int)floorf((lo[a][j] - mn) * inv) + 1 - 1;              // exact range in steps [0, 253]
                h = (int)ceilf((hi[a][j] - mn) * inv) + 1 + 1;               // shifted by one step (origin moved down below) and rounded outward by one more
                l = max(l, 0); h = min(h, 255);
            }
            qlo[a] |= (unsigned)l << (8 * j); qhi[a] |= (unsigned)h << (8 * j);
        }
        org[a] = mn - __int_as_float(e << 23);                              // the origin sits one step below the smallest plane: step 0 = mn - scale
    }
    uint4* out = nodes64 + 4 * (size_t)i;
    out[0] = make_uint4(__float_as_uint(org[0]), __float_as_uint(org[1]), __float_as_uint(org[2]), ex[0] | (ex[1] << 8) | (ex[2] << 16));
    out[1] = make_uint4(qlo[0], qhi[0], qlo[1], qhi[1]);
    out[2] = make_uint4(qlo[2], qhi[2], 0u, 0u);
    out[3] = make_uint4((unsigned)ref[0], (unsigned)ref[1], (unsigned)ref[2], (unsigned)ref[3]);
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

#define BT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t trt_compress_nodes(int n_nodes, const float4* nodes128, void* nodes64, cudaStream_t stream)
{
    if (n_nodes <= 0) return cudaSuccess;
    compress_nodes<<<(n_nodes + 255) / 256, 256, 0, stream>>>(n_nodes, nodes128, (uint4*)nodes64);
    return cudaGetLastError();
}

size_t trt_build_scratch_bytes(int n_tris)
{
    if (n_tris <= 0) return 256;
    size_t cub_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, n_tris, 0, 64, 0);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int*)nullptr, (int*)nullptr, n_tris, 0);
    if (scan_bytes > cub_bytes) cub_bytes = scan_bytes;
    cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, (const int*)nullptr, (int*)nullptr, cub::Max(), n_tris, 0);
    if (scan_bytes > cub_bytes) cub_bytes = scan_bytes;
    const size_t n = (size_t)n_tris;
    return 8 * align256(sizeof(float4) * n) + 2 * align256(8 * n) + 3 * align256(sizeof(int2) * n) + 6 * align256(sizeof(int) * n) + 256 + align256(cub_bytes) + 4096;
}

cudaError_t trt_build_scene(const BuildInputs& in, BuildOutputs& out, cudaStream_t stream)
{
    out.root = 0; out.max_depth = 0; out.n_twins = 0; out.sah_cost = 0.0; out.tri_tmin = 1e-7f;
    out.n_wide_nodes = 0; out.wide_depth = 0; out.binary_depth = 0; out.n_groups = 0;
    out.root_lo = make_float3(0, 0, 0); out.root_hi = make_float3(0, 0, 0);
    const int TB = 256;

    if (in.n_spheres > 0) {
        convert_spheres<<<(in.n_spheres + TB - 1) / TB, TB, 0, stream>>>(in.spheres_raw, in.n_spheres, out.sph, out.sph_mat);
        BT(cudaGetLastError());
    }
    const size_t n_texels = (size_t)in.n_materials * (size_t)in.tex_w * (size_t)in.tex_h;
    if (n_texels > 0) {
        int grid = (int)((n_texels + TB - 1) / TB); if (grid > 148 * 16) grid = 148 * 16;
        convert_texels<<<grid, TB, 0, stream>>>(in.mats_packed, n_texels, in.tex_w * in.tex_h,
                                                (const MatOverrideDev*)in.overrides_raw, in.n_overrides, out.texels);
        BT(cudaGetLastError());
    }
    const int n = in.n_tris;
    if (n == 0) return cudaSuccess;
    if (n > (1 << TRT_LEAF_POS_BITS)) return cudaErrorInvalidValue;

    // ---- scratch layout (the caller owns the buffer: trt_build_scratch_bytes) ------------------------
    size_t cub_bytes = 0, scan_bytes = 0;
    BT(cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, n, 0, 64, stream));
    BT(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int*)nullptr, (int*)nullptr, n, stream));
    if (scan_bytes > cub_bytes) cub_bytes = scan_bytes;
    BT(cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, (const int*)nullptr, (int*)nullptr, cub::Max(), n, stream));
    if (scan_bytes > cub_bytes) cub_bytes = scan_bytes;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_box_lo = take(sizeof(float4) * n), o_box_hi = take(sizeof(float4) * n);
    const size_t o_leaf_lo = take(sizeof(float4) * n), o_leaf_hi = take(sizeof(float4) * n);
    const size_t o_node_lo = take(sizeof(float4) * n), o_node_hi = take(sizeof(float4) * n);
    const size_t o_keys_a = take(sizeof(unsigned long long) * n), o_keys_b = take(sizeof(unsigned long long) * n);
    const size_t o_children = take(sizeof(int2) * n), o_parent_node = take(sizeof(int) * n), o_parent_leaf = take(sizeof(int) * n);
    const size_t o_flags = take(sizeof(int) * n);
    const size_t o_ghead = take(sizeof(int) * n), o_gpos = take(sizeof(int) * n), o_gfirst = take(sizeof(int) * n);
    const size_t o_queue_a = take(sizeof(int2) * n), o_queue_b = take(sizeof(int2) * n);
    const size_t o_c1lo = take(sizeof(float4) * n), o_c1hi = take(sizeof(float4) * n);     // PLOC ping-pong cluster boxes
    const size_t o_small = take(256);                 // bounds[7], max_depth, n_twins, sah (double at +64), collapse counters, PLOC node counter
    const size_t o_cub = take(cub_bytes);
    if (!in.scratch || in.scratch_bytes < off) return cudaErrorInvalidValue;
    char* base = (char*)in.scratch;
    float4 *box_lo = (float4*)(base + o_box_lo), *box_hi = (float4*)(base + o_box_hi);
    float4 *leaf_lo = (float4*)(base + o_leaf_lo), *leaf_hi = (float4*)(base + o_leaf_hi);
    float4 *node_lo = (float4*)(base + o_node_lo), *node_hi = (float4*)(base + o_node_hi);
    unsigned long long *keys_a = (unsigned long long*)(base + o_keys_a), *keys_b = (unsigned long long*)(base + o_keys_b);
    int2* children = (int2*)(base + o_children);
    int *parent_node = (int*)(base + o_parent_node), *parent_leaf = (int*)(base + o_parent_leaf), *flags = (int*)(base + o_flags);
    int* small = (int*)(base + o_small);
    double* sah = (double*)(base + o_small + 64);
    int *ghead = (int*)(base + o_ghead), *gpos = (int*)(base + o_gpos), *gfirst = (int*)(base + o_gfirst);
    const char* bvh_mode = getenv("TRT_BVH");
    const bool lbvh = bvh_mode && strcmp(bvh_mode, "lbvh") == 0;
    const char* grp = getenv("TRT_TUNE_GROUP");
    const int group = (!lbvh && !(grp && *grp && atoi(grp) == 0)) ? 1 : 0;       // equal-box runs become one leaf (PLOC only)

    // bounds init: min = +big, max = -big in ordered-int space
    int h_small[64];
    for (int k = 0; k < 64; k++) h_small[k] = 0;
    h_small[0] = h_small[1] = h_small[2] = 0x7f7fffff;                 // ord(+FLT_MAX)
    h_small[3] = h_small[4] = h_small[5] = (int)0x80800000;            // ord(-FLT_MAX) = 0xff7fffff ^ 0x7fffffff
    BT(cudaMemcpyAsync(small, h_small, sizeof h_small, cudaMemcpyHostToDevice, stream));
    BT(cudaMemsetAsync(flags, 0, sizeof(int) * n, stream));

    const int grid_n = (n + TB - 1) / TB;
    tri_bounds<<<grid_n, TB, 0, stream>>>(in.tris_raw, n, box_lo, box_hi, small);
    morton_keys<<<grid_n, TB, 0, stream>>>(n, box_lo, box_hi, small, keys_a);
    BT(cudaGetLastError());
    BT(cub::DeviceRadixSort::SortKeys(base + o_cub, cub_bytes, keys_a, keys_b, n, 0, 64, stream));

    const char* capv = getenv("TRT_TUNE_RUNCAP");
    const int run_cap = (capv && *capv && atoi(capv) > 0) ? atoi(capv) : 16;
    {
        int* run_mark = parent_leaf;            // free until the LBVH hierarchy pass, which comes after the grouping
        group_marks<<<grid_n, TB, 0, stream>>>(n, keys_b, box_lo, box_hi, group, run_mark);
        size_t tmp = cub_bytes;
        BT(cub::DeviceScan::InclusiveScan(base + o_cub, tmp, run_mark, gpos, cub::Max(), n, stream));      // gpos = start of the run, for now
        group_heads<<<grid_n, TB, 0, stream>>>(n, gpos, run_cap, ghead);
    }
    { size_t tmp = cub_bytes; BT(cub::DeviceScan::ExclusiveSum(base + o_cub, tmp, ghead, gpos, n, stream)); }
    group_first<<<grid_n, TB, 0, stream>>>(n, ghead, gpos, gfirst);
    BT(cudaGetLastError());

    emit_leaves<<<grid_n, TB, 0, stream>>>(n, keys_b, in.tris_raw, in.quel_mat, box_lo, box_hi, small,
                                          out.tri, out.tri_uv, out.tri_orig, leaf_lo, leaf_hi);
    find_twins<<<grid_n, TB, 0, stream>>>(n, keys_b, in.tris_raw, out.tri_uv, out.tri, small + 9);
    group_counts<<<grid_n, TB, 0, stream>>>(n, ghead, gpos, gfirst, out.tri);
    BT(cudaGetLastError());

    auto ord2f_h = [](int i) { int b = i >= 0 ? i : i ^ 0x7fffffff; float f; memcpy(&f, &b, 4); return f; };
    auto scene_scale = [&](const int* h_bounds) {       // box padding and the FP32 distance threshold from the scene's size
        const float amax = ord2f_h(h_bounds[6]);
        const float pad = amax * 4.76837158e-7f + 1e-30f;                    // 2^-21 relative
        out.tri_tmin = fmaxf(1e-7f, amax * 9.53674316e-7f);                  // 2^-20 relative
        out.root_lo = make_float3(ord2f_h(h_bounds[0]) - pad, ord2f_h(h_bounds[1]) - pad, ord2f_h(h_bounds[2]) - pad);
        out.root_hi = make_float3(ord2f_h(h_bounds[3]) + pad, ord2f_h(h_bounds[4]) + pad, ord2f_h(h_bounds[5]) + pad);
    };

    // ---- small scenes: the rest of the build is ONE launch, and the host reads everything back once -------------------
    const char* small_env = getenv("TRT_TUNE_SMALL_BUILD");
    if (!lbvh && n <= 4 * SMALL_MAX_GROUPS && !(small_env && *small_env && atoi(small_env) == 0)) {
        const char* rad = getenv("TRT_TUNE_PLOC_RADIUS");
        const int radius = rad && *rad ? atoi(rad) : 64;
        SmallResult* d_res = (SmallResult*)(small + 32);
        const size_t smem = (size_t)SMALL_MAX_GROUPS * (2 * sizeof(float4) + 2 * sizeof(int));
        BT(cudaFuncSetAttribute(build_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      // per device: set on every call (cheap), not once per process
        // optionally (TRT_TUNE_SMALL_PRE) the first passes, while there are thousands of clusters, run on the whole GPU (the neighbour
        // search of one pass over 2 035 clusters is 260 k box unions: 87 % of the single-CTA kernel's instructions); the tail of the loop,
        // the SAH report and the collapse run in one CTA.  Nothing comes back to the host in between: grids are sized for n, the count
        // lives on the device.
        int* cidp[2] = {(int*)(base + o_queue_a), (int*)(base + o_queue_b)};
        int* nn = cidp[0] + n; int* flag = cidp[1] + n; int* pos = parent_node;
        float4* clop[2] = {box_lo, (float4*)(base + o_c1lo)};
        float4* chip[2] = {box_hi, (float4*)(base + o_c1hi)};
        int* node_counter = small + 26;
        int* d_cnt = small + 28;
        ploc_init_dev<<<grid_n, TB, 0, stream>>>(n, ghead, gpos, gfirst, leaf_lo, leaf_hi, cidp[0], clop[0], chip[0], d_cnt, node_counter);
        const char* pre_env = getenv("TRT_TUNE_SMALL_PRE");
        const int pre = pre_env && *pre_env ? atoi(pre_env) : 0;         // measured on C2: 0 / 4 / 6 / 8 passes -> trt_scene_set 0.75 / 0.75 / 0.78 / 0.79 ms: the launches cost what the kernel saves
        int cur = 0;
        for (int pass = 0; pass < pre; pass++) {
            const int* d_c = d_cnt + (pass & 1); int* d_c_next = d_cnt + ((pass + 1) & 1);
            ploc_nearest<<<grid_n, TB, 0, stream>>>(d_c, clop[cur], chip[cur], radius, nn);
            ploc_merge<<<grid_n, TB, 0, stream>>>(d_c, n, cidp[cur], clop[cur], chip[cur], nn, node_counter, children, node_lo, node_hi, flag);
            { size_t tmp = cub_bytes; BT(cub::DeviceScan::ExclusiveSum(base + o_cub, tmp, flag, pos, n, stream)); }
            ploc_compact<<<grid_n, TB, 0, stream>>>(d_c, d_c_next, cidp[cur], clop[cur], chip[cur], flag, pos, cidp[cur ^ 1], clop[cur ^ 1], chip[cur ^ 1]);
            cur ^= 1;
        }
        build_small_kernel<<<1, SMALL_THREADS, smem, stream>>>(n, ghead, gpos, cidp[cur], clop[cur], chip[cur], d_cnt + (pre & 1), node_counter,
                                                                leaf_lo, leaf_hi, radius, children, node_lo, node_hi,
                                                                (int2*)(base + o_queue_a), (int2*)(base + o_queue_b), out.nodes, out.tri, d_res);      // the cluster arrays in the queues' memory are dead once the kernel has loaded them
        BT(cudaGetLastError());
        int h_small2[64];
        BT(cudaMemcpyAsync(h_small2, small, sizeof h_small2, cudaMemcpyDeviceToHost, stream));
        BT(cudaStreamSynchronize(stream));
        SmallResult res;
        memcpy(&res, h_small2 + 32, sizeof res);
        if (res.n_groups >= 1 && res.n_groups <= SMALL_MAX_GROUPS) {
            scene_scale(h_small2);
            out.n_groups = res.n_groups; out.root = res.root; out.n_wide_nodes = res.n_wide; out.wide_depth = res.wide_depth;
            out.max_depth = res.n_groups > 1 ? res.wide_depth + 1 : 1;
            out.n_twins = h_small2[9]; out.sah_cost = res.sah;
            return cudaSuccess;
        }
        // more distinct boxes than the single CTA holds: the kernel left at once, the multi-kernel path below takes over
    }

    int h_bounds[8], h_grp[2];
    BT(cudaMemcpyAsync(h_bounds, small, sizeof(int) * 8, cudaMemcpyDeviceToHost, stream));
    BT(cudaMemcpyAsync(h_grp, gpos + n - 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
    BT(cudaMemcpyAsync(h_grp + 1, ghead + n - 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
    BT(cudaStreamSynchronize(stream));
    const int n_groups = h_grp[0] + h_grp[1];
    if (n_groups < 1 || n_groups > n) return cudaErrorUnknown;
    scene_scale(h_bounds);
    out.n_groups = n_groups;
    if (n_groups == 1) {
        out.root = TRT_LEAF_REF(0, n); out.max_depth = 1; out.n_wide_nodes = 0; out.wide_depth = 0;       // one run of n <= 16 triangles
    } else {
        int root = 0;
        if (lbvh) {
            build_hierarchy<<<grid_n, TB, 0, stream>>>(n, keys_b, children, parent_node, parent_leaf);
            refit<<<grid_n, TB, 0, stream>>>(n, children, parent_node, parent_leaf, leaf_lo, leaf_hi, node_lo, node_hi, flags, small + 8);
        } else {
            // PLOC: clusters ping-pong between (queue_a | box_lo | box_hi) and (queue_b | c1lo | c1hi)
            int* cid[2] = {(int*)(base + o_queue_a), (int*)(base + o_queue_b)};
            int* nn = cid[0] + n; int* flag = cid[1] + n; int* pos = parent_node;
            float4* clo[2] = {box_lo, (float4*)(base + o_c1lo)};
            float4* chi[2] = {box_hi, (float4*)(base + o_c1hi)};
            int* node_counter = small + 26;
            const char* rad = getenv("TRT_TUNE_PLOC_RADIUS");
            const int radius = rad && *rad ? atoi(rad) : 64;
            BT(cudaMemsetAsync(node_counter, 0, sizeof(int), stream));
            ploc_init<<<(n_groups + TB - 1) / TB, TB, 0, stream>>>(n_groups, gfirst, leaf_lo, leaf_hi, cid[0], clo[0], chi[0]);
            // counts ping-pong between small[28] and small[29]; PLOC_BATCH passes per host read-back (a pass on a single
            // cluster copies it across unchanged, so overshooting the end is harmless)
            int* d_cnt = small + 28;
            int c = n_groups, cur = 0, passes = 0;
            BT(cudaMemcpyAsync(d_cnt, &c, sizeof(int), cudaMemcpyHostToDevice, stream));
            const int PLOC_BATCH = 8;
            while (c > 1) {
                const int c_bound = c;
                const int g = (c_bound + TB - 1) / TB;
                for (int b = 0; b < PLOC_BATCH; b++, passes++) {
                    const int* d_c = d_cnt + (passes & 1); int* d_c_next = d_cnt + ((passes + 1) & 1);
                    ploc_nearest<<<g, TB, 0, stream>>>(d_c, clo[cur], chi[cur], radius, nn);
                    ploc_merge<<<g, TB, 0, stream>>>(d_c, c_bound, cid[cur], clo[cur], chi[cur], nn, node_counter, children, node_lo, node_hi, flag);
                    { size_t tmp = cub_bytes; BT(cub::DeviceScan::ExclusiveSum(base + o_cub, tmp, flag, pos, c_bound, stream)); }
                    ploc_compact<<<g, TB, 0, stream>>>(d_c, d_c_next, cid[cur], clo[cur], chi[cur], flag, pos, cid[cur ^ 1], clo[cur ^ 1], chi[cur ^ 1]);
                    cur ^= 1;
                }
                BT(cudaGetLastError());
                int c_new = 0;
                BT(cudaMemcpyAsync(&c_new, d_cnt + (passes & 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
                BT(cudaStreamSynchronize(stream));
                if (c_new >= c || c_new < 1 || passes > 8192) return cudaErrorUnknown;     // every pass on > 1 clusters merges at least one pair
                c = c_new;
            }
            BT(cudaMemcpyAsync(&root, cid[cur], sizeof(int), cudaMemcpyDeviceToHost, stream));
            BT(cudaStreamSynchronize(stream));
            if (root < 0 || root >= n_groups - 1) return cudaErrorUnknown;
        }
        sah_cost_kernel<<<grid_n, TB, 0, stream>>>(lbvh ? n : n_groups, children, leaf_lo, leaf_hi, node_lo, node_hi, root, sah);
        BT(cudaGetLastError());
        // binary -> 4-wide, one launch per level (the queue sizes come back to the host: ~depth/2 round trips)
        int2* q[2] = {(int2*)(base + o_queue_a), (int2*)(base + o_queue_b)};
        int* counters = small + 24;
        const int2 root_entry = make_int2(root, 0);
        int h_counters[2] = {1, 0};
        BT(cudaMemcpyAsync(q[0], &root_entry, sizeof(int2), cudaMemcpyHostToDevice, stream));
        BT(cudaMemcpyAsync(counters, h_counters, sizeof h_counters, cudaMemcpyHostToDevice, stream));
        int n_in = 1, level = 0;
        while (n_in > 0) {
            collapse_level<<<(n_in + TB - 1) / TB, TB, 0, stream>>>(children, leaf_lo, leaf_hi, node_lo, node_hi,
                                                                   q[level & 1], n_in, q[(level + 1) & 1], counters, out.nodes, out.tri);
            BT(cudaGetLastError());
            BT(cudaMemcpyAsync(h_counters, counters, sizeof h_counters, cudaMemcpyDeviceToHost, stream));
            BT(cudaStreamSynchronize(stream));
            n_in = h_counters[1];
            h_counters[1] = 0;
            BT(cudaMemcpyAsync(counters + 1, h_counters + 1, sizeof(int), cudaMemcpyHostToDevice, stream));
            level++;
        }
        out.n_wide_nodes = h_counters[0];
        out.wide_depth = level;
        out.root = 0;
    }
    int h_tail[2]; double h_sah = 0.0;
    BT(cudaMemcpyAsync(h_tail, small + 8, sizeof(int) * 2, cudaMemcpyDeviceToHost, stream));
    BT(cudaMemcpyAsync(&h_sah, sah, sizeof(double), cudaMemcpyDeviceToHost, stream));
    BT(cudaStreamSynchronize(stream));
    if (n_groups > 1) out.binary_depth = h_tail[0];
    out.max_depth = n_groups > 1 ? out.wide_depth + 1 : 1;
    out.n_twins = h_tail[1];
    out.sah_cost = h_sah;
    return cudaSuccess;
}
