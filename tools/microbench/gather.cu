// Microbenchmark: cost of a divergent "node fetch" on sm_100a — every lane reads its own 64/128-byte record.
// Variants: 128-bit global loads, 256-bit global loads, shared-memory loads; record sizes 64 B and 128 B.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a gather.cu -o gather
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned next_idx(unsigned x, unsigned mask) { return (x * 1664525u + 1013904223u) & mask; }

template <int BYTES, int MODE>   // MODE 0: LDG.128, 1: LDG.256, 2: LDS.128 (table copied to smem)
__global__ void __launch_bounds__(256) gather(const float4* __restrict__ table, unsigned mask, int iters, float* out)
{
    extern __shared__ float4 s_tab[];
    if (MODE == 2) {
        for (unsigned k = threadIdx.x; k < (mask + 1) * (BYTES / 16); k += blockDim.x) s_tab[k] = table[k];
        __syncthreads();
    }
    unsigned idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u & mask;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
            const float4* p = table + (size_t)idx * (BYTES / 16);
#pragma unroll
            for (int k = 0; k < BYTES / 16; k++) { float4 v = __ldg(p + k); acc += v.x + v.y + v.z + v.w; }
        } else if (MODE == 1) {
            const float4* p = table + (size_t)idx * (BYTES / 16);
#pragma unroll
            for (int k = 0; k < BYTES / 32; k++) {
                float a, b, c, d, e, f, g, h;
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(h) : "l"(p + 2 * k));
                acc += a + b + c + d + e + f + g + h;
            }
        } else if (MODE == 3) {   // 256-bit loads, fields rotated by the record index: lanes spread over the 32-byte columns of a line
            const float4* p = table + (size_t)idx * (BYTES / 16);
#pragma unroll
            for (int k = 0; k < BYTES / 32; k++) {
                float a, b, c, d, e, f, g, h;
                const int kk = (k + idx) & (BYTES / 32 - 1);
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(h) : "l"(p + 2 * kk));
                acc += a + b + c + d + e + f + g + h;
            }
        } else if (MODE == 4) {   // 128-bit loads, rotated
            const float4* p = table + (size_t)idx * (BYTES / 16);
#pragma unroll
            for (int k = 0; k < BYTES / 16; k++) { float4 v = __ldg(p + ((k + idx) & (BYTES / 16 - 1))); acc += v.x + v.y + v.z + v.w; }
        } else {
            const float4* p = s_tab + (size_t)idx * (BYTES / 16);
#pragma unroll
            for (int k = 0; k < BYTES / 16; k++) { float4 v = p[k]; acc += v.x + v.y + v.z + v.w; }
        }
        idx = next_idx(idx + __float_as_uint(acc) % 3u, mask);     // data-dependent: next fetch waits for this one
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int BYTES, int MODE>
void run(const char* name, const float4* d_tab, unsigned n_rec, float* d_out, int blocks_per_sm)
{
    int dev = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * blocks_per_sm, iters = 2000;
    size_t smem = MODE == 2 ? (size_t)n_rec * BYTES : 0;
    if (smem) cudaFuncSetAttribute(gather<BYTES, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather<BYTES, MODE><<<grid, 256, smem>>>(d_tab, n_rec - 1, 10, d_out);
    cudaEventRecord(a);
    gather<BYTES, MODE><<<grid, 256, smem>>>(d_tab, n_rec - 1, iters, d_out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    const double fetches = (double)grid * 256 * iters;
    printf("%-28s rec=%3dB n=%6u blocks/SM=%d: %7.2f ms  %7.1f Gfetch/s  %6.2f TB/s  %s\n", name, BYTES, n_rec, blocks_per_sm, ms,
           fetches / ms / 1e6, fetches * BYTES / ms / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    const unsigned n_big = 1u << 20;
    float4* d_tab; float* d_out;
    cudaMalloc(&d_tab, (size_t)n_big * 128); cudaMemset(d_tab, 0, (size_t)n_big * 128);
    cudaMalloc(&d_out, 148 * 8 * 256 * sizeof(float));
    for (int bps : {4}) {
        // table sizes: 1024 records (L1-resident), 4096 (L1/L2), 1M (L2/HBM)
        for (unsigned n : {1024u, 4096u, 65536u}) {
            run<128, 0>("LDG.128 x8", d_tab, n, d_out, bps);
            run<128, 1>("LDG.256 x4", d_tab, n, d_out, bps);
            run<64, 0>("LDG.128 x4", d_tab, n, d_out, bps);
            run<64, 1>("LDG.256 x2", d_tab, n, d_out, bps);
            run<128, 3>("LDG.256 x4 rotated", d_tab, n, d_out, bps);
            run<64, 3>("LDG.256 x2 rotated", d_tab, n, d_out, bps);
            run<128, 4>("LDG.128 x8 rotated", d_tab, n, d_out, bps);
            run<64, 4>("LDG.128 x4 rotated", d_tab, n, d_out, bps);
        }
        if (bps <= 2) { run<128, 2>("LDS.128 x8 (smem table)", d_tab, 512, d_out, bps); run<64, 2>("LDS.128 x4 (smem table)", d_tab, 1024, d_out, bps); }
    }
    run<64, 2>("LDS.128 x4 (smem table)", d_tab, 2048, d_out, 1);
    run<128, 2>("LDS.128 x8 (smem table)", d_tab, 1024, d_out, 1);
    return 0;
}
