// kat.cu — known-answer kernels: the FP32 leaf functions of the path-tracing kernel applied to
// arrays of independent inputs, so tests can compare each of them with the FP64 oracle.
#include "trt_isect.cuh"
#include "trt_kernels.h"

namespace {

__global__ void kat_sphere(int n, const float4* cr, const float* ray, int* hit, float* dst, float* point, float* normal)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float3 o = f3(ray[6 * k], ray[6 * k + 1], ray[6 * k + 2]), d = f3(ray[6 * k + 3], ray[6 * k + 4], ray[6 * k + 5]);
    const float a = dot3(d, d);
    const float t = sphere_root(cr[k], o, d, a, __frcp_rn(a), false);
    hit[k] = t >= 0.0f;
    dst[k] = t;
    const float3 p = fma3(d, t, o);
    const float3 nn = normalize3(p - xyz(cr[k]));
    point[3 * k] = p.x; point[3 * k + 1] = p.y; point[3 * k + 2] = p.z;
    normal[3 * k] = nn.x; normal[3 * k + 1] = nn.y; normal[3 * k + 2] = nn.z;
}

__global__ void kat_triangle(int n, const float* abc, const float* ray, int* hit, float* dst_u_v)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float* t = abc + 9 * k;
    const float3 A = f3(t[0], t[1], t[2]), AB = f3(t[3] - t[0], t[4] - t[1], t[5] - t[2]), AC = f3(t[6] - t[0], t[7] - t[1], t[8] - t[2]);
    const float3 N = cross3(AB, AC);
    const float4 r0 = make_float4(A.x, A.y, A.z, AB.x), r1 = make_float4(AB.y, AB.z, AC.x, AC.y), r2 = make_float4(AC.z, N.x, N.y, N.z);
    const float3 o = f3(ray[6 * k], ray[6 * k + 1], ray[6 * k + 2]), d = f3(ray[6 * k + 3], ray[6 * k + 4], ray[6 * k + 5]);
    float tt, u, v;
    hit[k] = triangle_test(r0, r1, r2, o, d, 1e-7f, tt, u, v);
    dst_u_v[3 * k] = tt; dst_u_v[3 * k + 1] = u; dst_u_v[3 * k + 2] = v;
}

__global__ void kat_refract(int n, const float* in8, float* out3)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float* q = in8 + 8 * k;
    const float3 r = refract3(f3(q[0], q[1], q[2]), f3(q[3], q[4], q[5]), q[6], q[7]);
    out3[3 * k] = r.x; out3[3 * k + 1] = r.y; out3[3 * k + 2] = r.z;
}

__global__ void kat_philox(int n, const uint32_t* ck, uint32_t* out4)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t* q = ck + 6 * k;
    const uint4 r = philox4x32_10(q[0], q[1], q[2], q[3], q[4], q[5]);
    out4[4 * k] = r.x; out4[4 * k + 1] = r.y; out4[4 * k + 2] = r.z; out4[4 * k + 3] = r.w;
}

// one channel of write_color_canva (rtutility.h:56-71) on fixed-point sums, same code as resolve_kernel
__global__ void kat_tonemap(int n, const long long* sums, int spp, double* out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double rapport = 1.0 / (double)spp;
    const double sum = (double)sums[k] * TRT_ACCUM_INV_SCALE;
    float r = sqrtf((float)(rapport * sum));
    double cl = (double)r;
    cl = cl < 0.0 ? 0.0 : (cl > 0.999 ? 0.999 : cl);
    out[k] = (double)(int)(256.0 * cl);
}

inline int blocks(int n) { return (n + 127) / 128; }

}  // namespace

cudaError_t trt_kat_launch_sphere(int n, const float4* cr, const float* ray, int* hit, float* dst, float* point, float* normal, cudaStream_t s)
{ kat_sphere<<<blocks(n), 128, 0, s>>>(n, cr, ray, hit, dst, point, normal); return cudaGetLastError(); }
cudaError_t trt_kat_launch_triangle(int n, const float* abc, const float* ray, int* hit, float* dst_u_v, cudaStream_t s)
{ kat_triangle<<<blocks(n), 128, 0, s>>>(n, abc, ray, hit, dst_u_v); return cudaGetLastError(); }
cudaError_t trt_kat_launch_refract(int n, const float* in8, float* out3, cudaStream_t s)
{ kat_refract<<<blocks(n), 128, 0, s>>>(n, in8, out3); return cudaGetLastError(); }
cudaError_t trt_kat_launch_philox(int n, const uint32_t* ck, uint32_t* out4, cudaStream_t s)
{ kat_philox<<<blocks(n), 128, 0, s>>>(n, ck, out4); return cudaGetLastError(); }
cudaError_t trt_kat_launch_tonemap(int n, const long long* sums, int spp, double* out, cudaStream_t s)
{ kat_tonemap<<<blocks(n), 128, 0, s>>>(n, sums, spp, out); return cudaGetLastError(); }
