// primary.cu — K1: primary-visibility / AOV pass in FP64.
//
// One thread per pixel, ray through the pixel centre, every sphere then every triangle in the
// caller's order with strict '<' (closest_hit, main.c:52-92).  All arithmetic uses the
// round-to-nearest intrinsics (__dadd_rn, __dmul_rn, __ddiv_rn, __dsqrt_rn) in the reference's
// operation order, so nvcc cannot contract a*b+c into an FMA and the hit distance, id, normal and
// texel index are the CPU path's, bit for bit.  The albedo comes from the FP32 texel table.
// This is the parity pass (north_star test 1); K2 does the FP32 path tracing.
#include "trt_kernels.h"

namespace {

struct d3 { double x, y, z; };

__device__ __forceinline__ double ADD(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double SUB(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double MUL(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double DIV(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ d3 mk(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 add(d3 a, d3 b) { return mk(ADD(a.x, b.x), ADD(a.y, b.y), ADD(a.z, b.z)); }
__device__ __forceinline__ d3 sub(d3 a, d3 b) { return mk(SUB(a.x, b.x), SUB(a.y, b.y), SUB(a.z, b.z)); }
__device__ __forceinline__ d3 scale(d3 a, double t) { return mk(MUL(a.x, t), MUL(a.y, t), MUL(a.z, t)); }
__device__ __forceinline__ d3 divs(d3 a, double t) { return mk(DIV(a.x, t), DIV(a.y, t), DIV(a.z, t)); }
__device__ __forceinline__ double dot(d3 a, d3 b) { return ADD(ADD(MUL(a.x, b.x), MUL(a.y, b.y)), MUL(a.z, b.z)); }   // vec3.h:120-122
__device__ __forceinline__ d3 cross(d3 a, d3 b)                                                                           // vec3.h:124-130
{
    return mk(SUB(MUL(a.y, b.z), MUL(a.z, b.y)), SUB(MUL(a.z, b.x), MUL(a.x, b.z)), SUB(MUL(a.x, b.y), MUL(a.y, b.x)));
}
__device__ __forceinline__ d3 normalize(d3 a) { return divs(a, __dsqrt_rn(dot(a, a))); }                                  // vec3.h:136-142
__device__ __forceinline__ d3 ld3(const double* p) { return mk(p[0], p[1], p[2]); }

__global__ void __launch_bounds__(128)
primary_kernel(const __grid_constant__ PrimaryArgs P)
{
    const int W = P.W, H = P.H;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= W || j >= H) return;
    const DevScene& sc = P.scene;

    // get_ray (camera.h:42-54) with u = i/(W-1), v = j/(H-1), dx = dy = 0
    const double u = DIV((double)i, (double)(W - 1)), v = DIV((double)j, (double)(H - 1));
    const d3 org = ld3(P.cam), hor = ld3(P.cam + 3), ver = ld3(P.cam + 6), cbg = ld3(P.cam + 9);
    const d3 direction = add(cbg, add(scale(hor, u), sub(scale(ver, v), org)));
    const d3 dest = add(org, scale(direction, P.focus));
    const d3 o = add(org, mk(0.0, 0.0, 0.0));
    const d3 d = normalize(sub(dest, o));

    double best = __longlong_as_double(0x7ff0000000000000LL);
    int id = -1;
    d3 nrm = mk(0, 0, 0), hp = mk(0, 0, 0);

    // spheres — hit_sphere, sphere.h:13-47
    for (int k = 0; k < sc.n_spheres; k++) {
        const double* s = sc.sph64 + 14 * (size_t)k;
        const d3 c = ld3(s); const double radius = s[3];
        const d3 oc = sub(o, c);
        const double a = dot(d, d);
        const double b = MUL(2.0, dot(oc, d));
        const double cc = SUB(dot(oc, oc), MUL(radius, radius));
        const double disc = SUB(MUL(b, b), MUL(MUL(4.0, a), cc));
        if (disc > 0) {
            double t = -1.0; bool ok = false;
            const double t1 = DIV(SUB(-b, __dsqrt_rn(disc)), MUL(2.0, a));
            if (t1 >= 0.0001) { t = t1; ok = true; }
            else {
                const double t2 = DIV(ADD(-b, __dsqrt_rn(disc)), MUL(2.0, a));
                if (t2 >= 0.0001) { t = t2; ok = true; }
            }
            if (ok && t < best) {
                best = t; id = k;
                hp = add(o, scale(d, t));
                nrm = normalize(sub(hp, c));
            }
        }
    }

    // triangles — hit_triangle, mesh.h:70-94
    for (int k = 0; k < sc.n_tris; k++) {
        const double* tr = sc.tri64 + 25 * (size_t)k;
        const d3 A = ld3(tr), B = ld3(tr + 3), C = ld3(tr + 6);
        const d3 ab = sub(B, A), ac = sub(C, A);
        const d3 n = cross(ab, ac);
        const d3 ao = sub(o, A);
        const d3 dao = cross(ao, d);
        const double det = -dot(d, n);
        const double inv = DIV(1.0, det);
        const double dst = MUL(dot(ao, n), inv);
        const double bu = MUL(dot(ac, dao), inv);
        const double bv = MUL(-dot(ab, dao), inv);
        const double bw = SUB(SUB(1.0, bu), bv);
        const bool hit = det >= 1E-6 && dst >= 0.0000001 && bu >= 0.0000001 && bv >= 0.0000001 && bw >= 0.0000001;
        if (hit && dst < best) {
            best = dst; id = sc.n_spheres + k;
            hp = add(o, scale(d, dst));
            nrm = normalize(n);
        }
    }

    const size_t p = (size_t)j * W + i;
    if (P.t) P.t[p] = best;
    if (P.id) P.id[p] = id;
    if (P.normal) { P.normal[3 * p] = nrm.x; P.normal[3 * p + 1] = nrm.y; P.normal[3 * p + 2] = nrm.z; }
    if (P.albedo) {
        double r = 0, g = 0, bl = 0;
        if (id >= 0 && id < sc.n_spheres) {
            const double* s = sc.sph64 + 14 * (size_t)id;
            r = s[4]; g = s[5]; bl = s[6];                       // sphere.mat.diffuseColor
        } else if (id >= sc.n_spheres) {
            // tri_uvmapping, texture.h:16-27 + 44-68, in FP64 up to the texel index
            const int k = id - sc.n_spheres;
            const double* tr = sc.tri64 + 25 * (size_t)k;
            const d3 A = ld3(tr), B = ld3(tr + 3), C = ld3(tr + 6);
            const double areaABC = dot(nrm, cross(sub(B, A), sub(C, A)));
            const double areaPBC = dot(nrm, cross(sub(B, hp), sub(C, hp)));
            const double areaPCA = dot(nrm, cross(sub(C, hp), sub(A, hp)));
            const double b0 = DIV(areaPBC, areaABC), b1 = DIV(areaPCA, areaABC);
            const double b2 = SUB(SUB(1.0, b0), b1);
            const double* uv = tr + 19;                         // uvA, uvB, uvC
            double tu = ADD(ADD(MUL(b0, uv[0]), MUL(b1, uv[2])), MUL(b2, uv[4]));
            double tv = ADD(ADD(MUL(b0, uv[1]), MUL(b1, uv[3])), MUL(b2, uv[5]));
            tu = fmod(tu, 1.0); tv = fmod(tv, 1.0);
            if (tu < 0) tu = ADD(tu, 1.0);
            if (tv < 0) tv = ADD(tv, 1.0);
            // a tiny negative uv + 1.0 rounds to exactly 1.0: x == tex_w reads past the slab (the reference has the same
            // hazard on the host, texture.h:61-62); clamped like the FP32 path (trt_isect.cuh, triangle_material)
            const int x = min((int)MUL(tu, (double)sc.tex_w), sc.tex_w - 1), y = min((int)MUL(tv, (double)sc.tex_h), sc.tex_h - 1);
            const int ordinal = sc.quel_mat[k];
            const size_t texel = (size_t)(y * sc.tex_w + x) + (size_t)sc.tex_h * (size_t)sc.tex_w * (size_t)ordinal;
            const float4 m0 = sc.texels[3 * texel];
            r = m0.x; g = m0.y; bl = m0.z;
        }
        P.albedo[3 * p] = r; P.albedo[3 * p + 1] = g; P.albedo[3 * p + 2] = bl;
    }
}

}  // namespace

cudaError_t trt_launch_primary(const PrimaryArgs& P, cudaStream_t stream)
{
    dim3 block(128, 1), grid((P.W + 127) / 128, P.H);
    primary_kernel<<<grid, block, 0, stream>>>(P);
    return cudaGetLastError();
}
