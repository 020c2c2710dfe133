// api.cu — the C ABI of libtipe_rt.so (include/tipe_rt.h): context, scene upload, the render call
// that replaces main.c:402-453, and the device-resident pieces it is made of.
//
// Host logic only; every kernel lives in pathtrace.cu / primary.cu / bvh_build.cu / kat.cu.
// There is no CPU rendering path in this library: no device => trt_init fails.
#include <chrono>
#include <new>
#include <exception>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tipe_rt.h"
#include "trt_kernels.h"

namespace {

thread_local std::string g_init_error;

struct Buf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes == 0) bytes = 256;
        if (cap >= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes + bytes / 8);      // a little headroom for slightly larger scenes
        if (e == cudaSuccess) cap = bytes + bytes / 8;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DeviceCtx {
    int dev = 0;
    int n_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;
    // scene: grow-only device buffers, reused by every trt_scene_set (no cudaMalloc/cudaFree churn per frame)
    Buf sph_raw, tri_raw, quel_mat, sph, sph_mat, tri, tri_uv, tri_orig, nodes, texels, mats_raw, ov_raw, scratch, sky;
    Buf prehit;                                // engine 4: primary hits of the frame, 16 B per path
    Buf nodes64;                               // experiment (TRT_TUNE_NODE64): 64-byte quantised nodes
    DevScene scene{};
    // work buffers
    unsigned long long* ctrs = nullptr;        // [0] item counter, [1..4] rays / box / tri / sphere tests
    unsigned long long* h_ctrs = nullptr;      // pinned mirror
    long long* accum = nullptr; size_t accum_px = 0;
    int last_launches = 0;
    double* out = nullptr; size_t out_px = 0;  // canva | albedo | normal, each [npix][3]
    // device -> caller copy: pinned staging + one event per chunk, so that host threads can move finished chunks into the
    // caller's (pageable) arrays while the DMA engine brings the next ones
    char* h_stage = nullptr; size_t stage_bytes = 0;
    std::vector<cudaEvent_t> chunk_ev;
};

}  // namespace

// Pinned ring for large host -> device uploads (texture slabs, big meshes): host threads pack / copy chunks into the ring
// while the DMA engine ships the previous ones, instead of one pageable cudaMemcpy the driver stages at a few GB/s.
struct UploadRing {
    char* h = nullptr;
    size_t chunk = (size_t)1 << 20;
    int slots = 8;
    cudaEvent_t ev[8] = {};
    cudaEvent_t done = nullptr;
    cudaStream_t stream = nullptr;
};

struct trt_ctx {
    std::vector<DeviceCtx> devs;
    std::string err;
    bool has_scene = false;
    trt_bvh_info info{};
    UploadRing ring;            // on devs[0]
};

namespace {

int fail(trt_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (ctx) ctx->err = buf; else g_init_error = buf;
    return code;
}

#define CU(ctx, call)                                                                               \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? TRT_ERR_OOM : TRT_ERR_CUDA,          \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

void free_scene(DeviceCtx& d)
{
    cudaSetDevice(d.dev);
    Buf* bufs[] = {&d.sph_raw, &d.tri_raw, &d.quel_mat, &d.sph, &d.sph_mat, &d.tri, &d.tri_uv, &d.tri_orig, &d.nodes, &d.texels, &d.mats_raw, &d.ov_raw, &d.scratch};
    for (Buf* b : bufs) b->release();
    d.scene = DevScene{};
}

int ensure_accum(trt_ctx* ctx, DeviceCtx& d, size_t npix)
{
    if (d.accum_px >= npix) return TRT_OK;
    if (d.accum) cudaFree(d.accum);
    d.accum = nullptr; d.accum_px = 0;
    CU(ctx, cudaMalloc(&d.accum, sizeof(long long) * 9 * npix));
    d.accum_px = npix;
    return TRT_OK;
}

int ensure_out(trt_ctx* ctx, DeviceCtx& d, size_t npix)
{
    if (d.out_px >= npix) return TRT_OK;
    if (d.out) cudaFree(d.out);
    d.out = nullptr; d.out_px = 0;
    CU(ctx, cudaMalloc(&d.out, sizeof(double) * 9 * npix + 512));      // 9 doubles per pixel, or the compact layout of copy_out (256-byte padding)
    d.out_px = npix;
    return TRT_OK;
}

const size_t COPY_CHUNK = (size_t)2 << 20;

// Grows the pinned staging buffer; called while the trace kernel runs, so the page-locking is hidden behind it.
void ensure_stage(DeviceCtx& d, size_t bytes)
{
    const char* v = getenv("TRT_TUNE_STAGED");
    if (v && *v && atoi(v) == 0) return;
    if (bytes < COPY_CHUNK) return;                 // small frames: plain copies
    if (d.stage_bytes >= bytes) return;
    if (d.h_stage) cudaFreeHost(d.h_stage);
    d.h_stage = nullptr; d.stage_bytes = 0;
    if (cudaHostAlloc((void**)&d.h_stage, bytes, cudaHostAllocDefault) == cudaSuccess) d.stage_bytes = bytes;
    else { d.h_stage = nullptr; cudaGetLastError(); }
    const size_t n_chunks = (bytes + COPY_CHUNK - 1) / COPY_CHUNK + 1;      // + the chunk of canva bytes
    while (d.chunk_ev.size() < n_chunks) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); break; }
        d.chunk_ev.push_back(e);
    }
}

// Layout of d.out when the three arrays go to the HOST: canva as 3 BYTES per pixel (the values are integers 0..255: exact, and
// 21 of the 72 bytes a pixel costs over PCIe stay home), padded to 256 bytes, then albedo and normal as 3 doubles per pixel.
size_t canva8_bytes(size_t npix) { return (3 * npix + 255) & ~(size_t)255; }

// d.out -> the caller's three arrays (canva widened to doubles on the way), then waits for the stream.
// Two halves, so that several devices can have their copies in flight before the host starts draining the first one:
// copy_out_issue enqueues the device -> pinned-staging copies (one event per chunk) and returns whether the staged path is used;
// copy_out_drain moves what arrives into the caller's arrays (canva widened to doubles) and waits for the stream.
bool copy_out_issue(trt_ctx* ctx, DeviceCtx& d, size_t npix, int* rc_out)
{
    *rc_out = TRT_OK;
    const size_t c8 = canva8_bytes(npix), part = sizeof(double) * 3 * npix, total = c8 + 2 * part;
    const size_t n_chunks = 1 + (2 * part + COPY_CHUNK - 1) / COPY_CHUNK;
    const bool staged = d.h_stage && d.stage_bytes >= total && d.chunk_ev.size() >= n_chunks && total >= COPY_CHUNK;
    if (!staged) return false;
    const char* src = (const char*)d.out;
    auto step = [&](cudaError_t e) { if (e != cudaSuccess && *rc_out == TRT_OK) *rc_out = fail(ctx, TRT_ERR_CUDA, "device -> host copy failed: %s", cudaGetErrorString(e)); };
    // chunk 0 = the canva bytes (they arrive first and every thread widens a slice of them), chunks 1.. = 2 MB pieces of the two AOV tables
    const size_t n_aov = (2 * part + COPY_CHUNK - 1) / COPY_CHUNK;
    step(cudaMemcpyAsync(d.h_stage, src, c8, cudaMemcpyDeviceToHost, d.stream));
    step(cudaEventRecord(d.chunk_ev[0], d.stream));
    for (size_t c = 0; c < n_aov; c++) {
        const size_t off = c8 + c * COPY_CHUNK, len = (c + 1) * COPY_CHUNK <= 2 * part ? COPY_CHUNK : 2 * part - c * COPY_CHUNK;
        step(cudaMemcpyAsync(d.h_stage + off, src + off, len, cudaMemcpyDeviceToHost, d.stream));
        step(cudaEventRecord(d.chunk_ev[1 + c], d.stream));
    }
    return true;
}

// One band of the image on one device, on its way to the caller's arrays (pointers already offset to the band).
struct OutJob {
    DeviceCtx* d; size_t nb; int mode;                     // mode: 0 plain copies, 1 staged (copy_out_issue), 2 direct into page-locked arrays
    trt_vec3 *canva, *albedo_tab, *normal_tab;
};

// Finishes all jobs: ONE team of host threads drains the staged bands of all devices (a thread per band and device would spend
// more time being created than copying), then every stream is waited for.
int drain_jobs(trt_ctx* ctx, const OutJob* jobs, int n_jobs)
{
    auto widen = [](double* dst, const unsigned char* s8, size_t count) { for (size_t k = 0; k < count; k++) dst[k] = (double)s8[k]; };
    bool any_staged = false;
    for (int j = 0; j < n_jobs; j++) {
        const OutJob& J = jobs[j];
        if (J.nb == 0) continue;
        if (J.mode == 1) { any_staged = true; continue; }
        CU(ctx, cudaSetDevice(J.d->dev));
        if (J.mode == 0) {
            const size_t c8 = canva8_bytes(J.nb), part = sizeof(double) * 3 * J.nb;
            const char* src = (const char*)J.d->out;
            std::vector<unsigned char> tmp(J.canva ? 3 * J.nb : 0);
            if (J.canva) CU(ctx, cudaMemcpyAsync(tmp.data(), src, 3 * J.nb, cudaMemcpyDeviceToHost, J.d->stream));
            if (J.albedo_tab) CU(ctx, cudaMemcpyAsync(J.albedo_tab, src + c8, part, cudaMemcpyDeviceToHost, J.d->stream));
            if (J.normal_tab) CU(ctx, cudaMemcpyAsync(J.normal_tab, src + c8 + part, part, cudaMemcpyDeviceToHost, J.d->stream));
            CU(ctx, cudaStreamSynchronize(J.d->stream));
            if (J.canva) widen((double*)J.canva, tmp.data(), 3 * J.nb);
        }
    }
    if (any_staged) {
        // the host side (bytes -> doubles, staging -> the caller's pageable arrays) is the slow half, not PCIe; measured on the B200 box's
        // 16-core host for C2's 750 000 pixels: 1 thread 5.6 ms, 2: 3.0, 4: 2.0, 8: 1.3, 16: 1.15
        int n_threads = 16;
        { const unsigned hc = std::thread::hardware_concurrency(); if (hc > 0 && (int)hc < n_threads) n_threads = (int)hc; }
        { const char* v = getenv("TRT_TUNE_COPY_THREADS"); if (v && *v) n_threads = atoi(v); }
        if (n_threads < 1) n_threads = 1;
        if (n_threads > 16) n_threads = 16;
        cudaError_t errs[16];
        auto worker = [&](int t) {
            errs[t] = cudaSuccess;
            for (int j = 0; j < n_jobs && errs[t] == cudaSuccess; j++) {
                const OutJob& J = jobs[j];
                if (J.nb == 0 || J.mode != 1) continue;
                DeviceCtx& d = *J.d;
                const size_t c8 = canva8_bytes(J.nb), part = sizeof(double) * 3 * J.nb, n_aov = (2 * part + COPY_CHUNK - 1) / COPY_CHUNK;
                char* aov_dst[2] = {(char*)J.albedo_tab, (char*)J.normal_tab};
                errs[t] = cudaSetDevice(d.dev);
                if (errs[t] == cudaSuccess) errs[t] = cudaEventSynchronize(d.chunk_ev[0]);
                if (errs[t] == cudaSuccess && J.canva) {
                    const size_t n8 = 3 * J.nb, lo = n8 * (size_t)t / (size_t)n_threads, hi = n8 * (size_t)(t + 1) / (size_t)n_threads;
                    widen((double*)J.canva + lo, (const unsigned char*)d.h_stage + lo, hi - lo);
                }
                for (size_t c = (size_t)t; c < n_aov && errs[t] == cudaSuccess; c += (size_t)n_threads) {
                    errs[t] = cudaEventSynchronize(d.chunk_ev[1 + c]);
                    size_t off = c * COPY_CHUNK;
                    const size_t end = (c + 1) * COPY_CHUNK <= 2 * part ? (c + 1) * COPY_CHUNK : 2 * part;
                    while (off < end && errs[t] == cudaSuccess) {          // a chunk may straddle the two tables
                        const size_t a = off / part, in = off - a * part, len = (end - off < part - in) ? end - off : part - in;
                        if (aov_dst[a]) memcpy(aov_dst[a] + in, d.h_stage + c8 + off, len);
                        off += len;
                    }
                }
            }
        };
        std::thread th[15];
        int started = 1;
        try {
            for (int t = 1; t < n_threads; t++) { th[t - 1] = std::thread(worker, t); started = t + 1; }
        } catch (...) {                                     // thread creation failed: never let an exception cross the C ABI; the missing shares are done here
        }
        worker(0);
        for (int t = 1; t < started; t++) th[t - 1].join();
        for (int t = started; t < n_threads; t++) worker(t);
        for (int t = 0; t < n_threads; t++) CU(ctx, errs[t]);
    }
    for (int j = 0; j < n_jobs; j++) {
        if (jobs[j].nb == 0) continue;
        CU(ctx, cudaSetDevice(jobs[j].d->dev));
        CU(ctx, cudaStreamSynchronize(jobs[j].d->stream));
    }
    return TRT_OK;
}

cudaError_t ensure_ring(trt_ctx* ctx)
{
    UploadRing& r = ctx->ring;
    if (r.h) return cudaSuccess;
    cudaError_t e = cudaHostAlloc((void**)&r.h, r.chunk * r.slots, cudaHostAllocDefault);
    if (e != cudaSuccess) { r.h = nullptr; return e; }
    if ((e = cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
    for (int k = 0; k < r.slots; k++) if ((e = cudaEventCreateWithFlags(&r.ev[k], cudaEventDisableTiming)) != cudaSuccess) return e;
    return cudaEventCreateWithFlags(&r.done, cudaEventDisableTiming);
}

// n items of in_bytes each at `src` (host, pageable) -> n items of out_bytes each at `dst` on devs[0], through pack(dst_chunk,
// src_chunk, count) run by host threads into the pinned ring.  d.stream waits for the last chunk; the call returns when every
// chunk has been packed (the caller's array may be freed), not when the copies have landed.
template <class Pack>
int staged_upload(trt_ctx* ctx, char* dst, const char* src, size_t n, size_t in_bytes, size_t out_bytes, Pack pack)
{
    DeviceCtx& d = ctx->devs[0];
    CU(ctx, ensure_ring(ctx));
    UploadRing& r = ctx->ring;
    const size_t per_chunk = r.chunk / out_bytes, n_chunks = (n + per_chunk - 1) / per_chunk;
    int n_threads = 8;                                   // a power of two that divides r.slots: chunk c and chunk c - slots belong to the same thread
    { const unsigned hc = std::thread::hardware_concurrency(); while (n_threads > 1 && (hc > 0 && (unsigned)n_threads > hc)) n_threads >>= 1; }
    while (n_threads > 1 && (size_t)n_threads > n_chunks) n_threads >>= 1;
    cudaError_t errs[8];
    auto worker = [&](int t) {
        errs[t] = cudaSetDevice(d.dev);
        for (size_t c = (size_t)t; c < n_chunks && errs[t] == cudaSuccess; c += (size_t)n_threads) {
            const int slot = (int)(c % (size_t)r.slots);
            if (c >= (size_t)r.slots) errs[t] = cudaEventSynchronize(r.ev[slot]);           // the slot's previous copy has left the ring
            if (errs[t] != cudaSuccess) break;
            const size_t first = c * per_chunk, count = (first + per_chunk <= n) ? per_chunk : n - first;
            char* h = r.h + (size_t)slot * r.chunk;
            pack(h, src + first * in_bytes, count);
            errs[t] = cudaMemcpyAsync(dst + first * out_bytes, h, count * out_bytes, cudaMemcpyHostToDevice, r.stream);
            if (errs[t] == cudaSuccess) errs[t] = cudaEventRecord(r.ev[slot], r.stream);
        }
    };
    std::thread th[7];
    int started = 1;
    try {
        for (int t = 1; t < n_threads; t++) { th[t - 1] = std::thread(worker, t); started = t + 1; }
    } catch (...) {                                     // no more threads to be had: the ones that started finish, the rest of the chunks are done here
    }
    worker(0);
    for (int t = 1; t < started; t++) th[t - 1].join();
    for (int t = started; t < n_threads; t++) worker(t);
    for (int t = 0; t < n_threads; t++) CU(ctx, errs[t]);
    CU(ctx, cudaEventRecord(r.done, r.stream));
    CU(ctx, cudaStreamWaitEvent(d.stream, r.done, 0));
    // the ring is reused by the next upload: its chunks must have left before that one overwrites them
    CU(ctx, cudaEventSynchronize(r.done));
    return TRT_OK;
}

const size_t STAGED_MIN_BYTES = (size_t)4 << 20;        // below this a plain cudaMemcpyAsync is as fast

int check_params(trt_ctx* ctx, const trt_params* p)
{
    if (!ctx) return TRT_ERR_INVALID;
    if (!p) return fail(ctx, TRT_ERR_INVALID, "params is NULL");
    if (p->largeur_image < 2 || p->hauteur_image < 2) return fail(ctx, TRT_ERR_INVALID, "image must be at least 2x2 (u = i/(W-1), main.c:265)");
    if ((long long)p->largeur_image * p->hauteur_image > (1LL << 30)) return fail(ctx, TRT_ERR_INVALID, "image too large");
    if (p->nbRayonParPixel < 1) return fail(ctx, TRT_ERR_INVALID, "nbRayonParPixel must be >= 1");
    if (p->nbRebondMax < 0) return fail(ctx, TRT_ERR_INVALID, "nbRebondMax must be >= 0");
    if (p->sample_offset < 0 || p->sample_total < 0) return fail(ctx, TRT_ERR_INVALID, "negative sample_offset / sample_total");
    // the ambient-occlusion term is 1/AO_intensity (main.c:115): 0 (what ThreadData's int truncation makes of 0.x, main.c:42)
    // is a division by zero there and would saturate every sample here
    if (p->useAO && !(p->AO_intensity > 0.0)) return fail(ctx, TRT_ERR_INVALID, "useAO needs AO_intensity > 0 (got %g)", p->AO_intensity);
    // int64 sums: a sample is clamped to 2^20 at 2^28 fixed point = 2^48, so 2^14 saturated samples still fit with a sign bit to spare
    {
        const long long total = p->sample_total > 0 ? p->sample_total : p->nbRayonParPixel;
        if (total > (1LL << 30) || (long long)p->sample_offset + p->nbRayonParPixel > (1LL << 30))
            return fail(ctx, TRT_ERR_INVALID, "more than 2^30 samples per pixel");
    }
    if (p->tile_part != 0) {
        const int part = p->tile_part & 0xffff, parts = (p->tile_part >> 16) & 0xffff;
        if (parts < 1 || part >= parts) return fail(ctx, TRT_ERR_INVALID, "tile_part = %d: part %d of %d", p->tile_part, part, parts);
    }
    if (!ctx->has_scene) return fail(ctx, TRT_ERR_NO_SCENE, "trt_scene_set has not been called");
    return TRT_OK;
}

// Launches K2 for `spp` samples starting at global sample `offset` into `accum` on device d.
// scene, image size, camera and the per-frame constants of RenderArgs (shared by the path tracer and trt_primary_fp32)
void fill_camera_args(RenderArgs& A, DeviceCtx& d, const trt_params* p)
{
    A.scene = d.scene;
    { const char* v = getenv("TRT_TUNE_BOXTMIN"); A.scene.box_tmin = (v && *v && atoi(v) == 0) ? 0.0f : d.scene.tri_tmin; }
    A.W = p->largeur_image; A.H = p->hauteur_image;
    A.bounces = p->nbRebondMax;
    A.use_ao = p->useAO ? 1 : 0;
    A.inv_ao = (float)(1.0 / p->AO_intensity);
    {
        const long long total = p->sample_total > 0 ? p->sample_total : p->nbRayonParPixel;      // the same on every GPU / rank of a job
        A.sample_clamp = total <= (1LL << 14) ? TRT_SAMPLE_CLAMP : (float)((double)(1LL << 34) / (double)total);
    }
    A.emit_scale_ao = (float)(1.5 * p->AO_intensity);
    A.focus = (float)p->focus_distance;
    A.aperture_x = (float)p->ouverture_x; A.aperture_y = (float)p->ouverture_y;
    A.inv_wm1 = (float)(1.0 / (double)(A.W - 1)); A.inv_hm1 = (float)(1.0 / (double)(A.H - 1));
    const trt_camera& c = p->cam;
    A.cam_origin = make_float3((float)c.origin.e[0], (float)c.origin.e[1], (float)c.origin.e[2]);
    A.cam_hor = make_float3((float)c.horizontal.e[0], (float)c.horizontal.e[1], (float)c.horizontal.e[2]);
    A.cam_ver = make_float3((float)c.vertical.e[0], (float)c.vertical.e[1], (float)c.vertical.e[2]);
    A.cam_corner = make_float3((float)(c.coin_bas_gauche.e[0] - c.origin.e[0]), (float)(c.coin_bas_gauche.e[1] - c.origin.e[1]),
                               (float)(c.coin_bas_gauche.e[2] - c.origin.e[2]));
    A.key0 = (uint32_t)(p->seed & 0xffffffffu); A.key1 = (uint32_t)(p->seed >> 32);
}

int launch_trace(trt_ctx* ctx, DeviceCtx& d, const trt_params* p, int spp, int offset, long long* accum, int tile_offset = 0, int tile_stride = 1)
{
    RenderArgs A{};
    fill_camera_args(A, d, p);
    A.spp = spp; A.sample_offset = offset;
    // tuning knobs (defaults chosen from measurements on B200); env overrides are for experiments
    auto knob = [](const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; };
    const bool brute = (p->flags & TRT_FLAG_BRUTE_FORCE) != 0, count_tests = (p->flags & TRT_FLAG_COUNT_TESTS) != 0;
    // engine 1 = register-resident megakernel (default), 3 = slot-pool kernel for meshes (opt-in, TRT_ENGINE=3)
    const int chunk_max = knob("TRT_TUNE_CHUNK", 16);
    A.min_steps = knob("TRT_TUNE_MIN_STEPS", 2);
    A.exit_num = knob("TRT_TUNE_EXIT", 10);
    // sphere scenes: the lanes of a warp start their next paths TOGETHER, once none of them has a path left (32 = all): closed
    // scenes give paths of similar length, and lanes that stay in step share the bounce index, the AO / bounce phase and most
    // branches (C1 16.4 -> 12.5 ms, C5 -24 %, C3 -18 %); meshes: every lane re-arms on its own (paths of 1-4 rays)
    A.start_min = knob("TRT_TUNE_START_MIN", d.scene.n_tris > 0 ? 1 : 32);
    A.max_hk = knob("TRT_TUNE_MAX_HK", 3);
    A.hk_min = knob("TRT_TUNE_HK_MIN", 28);          // another housekeeping pass while fewer than 28/32 of the live lanes have BVH work (and some lane has a hit to consume)
    A.defer = knob("TRT_TUNE_DEFER", 3);
    A.exp_flags = knob("TRT_TUNE_EXP", 0);
    A.top_nodes = knob("TRT_TUNE_TOPK", 0);
    if (A.top_nodes > ctx->info.n_nodes) A.top_nodes = ctx->info.n_nodes;
    if (A.top_nodes > 256) A.top_nodes = 256;
    A.item_order = knob("TRT_TUNE_ORDER", d.scene.n_tris > 0 ? 1 : 0);   // meshes: a warp starts on the samples of 2-3 neighbouring pixels (coherent primary rays, +2-3 %); spheres: one tile (+5 %)
    A.count_tests = count_tests ? 1 : 0;
    const int blocks_per_sm = knob("TRT_TUNE_BLOCKS", 0);     // 0: kernel default (4 CTAs of 256 threads per SM, 64 registers)
    A.chunk = spp < chunk_max ? spp : chunk_max;
    if (d.scene.n_tris > 0 && !getenv("TRT_TUNE_CHUNK")) {
        // meshes (chunk-fastest item order): about 20 chunks per pixel, so that the 32 items a warp reserves are the chunks of
        // one or two pixels — near-identical primary rays — while an item still amortises its flush (C2, 200 spp: chunk 10
        // 33.5 ms vs chunk 16 35.3 ms; 20 spp: chunk 4 7.3 ms vs 10.7 ms)
        int c = (spp + 10) / 20;
        A.chunk = c < 4 ? (spp < 4 ? spp : 4) : (c > chunk_max ? chunk_max : c);
    }
    A.n_chunks = (spp + A.chunk - 1) / A.chunk;
    A.tiles_x = (A.W + 7) / 8; A.tiles_y = (A.H + 3) / 4;
    // items a warp reserves per atomic: one per lane on sphere scenes; on meshes (chunk-fastest order) the chunks of 3-4
    // neighbouring pixels, so that the warp stays on them for a while (C2, 20 chunks: 64 items +1.8 % over 32; C4, 32 chunks:
    // 128 items +7.5 %)
    A.pool = 32;
    if (d.scene.n_tris > 0) { int p = ((7 * A.n_chunks / 2 + 16) / 32) * 32; A.pool = p < 32 ? 32 : (p > 256 ? 256 : p); }
    A.pool = knob("TRT_TUNE_POOL", A.pool);
    A.tile_offset = tile_offset; A.tile_stride = tile_stride < 1 ? 1 : tile_stride;
    const long long n_tiles = (long long)A.tiles_x * A.tiles_y;
    const long long my_tiles = n_tiles > A.tile_offset ? (n_tiles - A.tile_offset + A.tile_stride - 1) / A.tile_stride : 0;
    A.n_items = my_tiles * A.n_chunks * 32;
    {
        auto magic = [](unsigned d) { unsigned long long m = (1ULL << 32) / d; return (unsigned)(m > 0xffffffffULL ? 0xffffffffULL : m); };
        A.per_tile = 32u * (unsigned)A.n_chunks;
        A.m_per_tile = magic(A.per_tile); A.m_chunks = magic((unsigned)A.n_chunks); A.m_tiles_x = magic((unsigned)A.tiles_x);
        A.items_fit32 = (A.n_items + 64 < (1LL << 32)) ? 1 : 0;        // + the pool overshoot of the last reservations
    }
    A.item_counter = d.ctrs;
    A.counters = d.ctrs + 1;
    A.accum = accum;
    CU(ctx, cudaMemsetAsync(d.ctrs, 0, sizeof(unsigned long long) * 16, d.stream));
    // engine 3 = slot-pool kernel (pathtrace_pool.cu), opt-in (TRT_ENGINE=3): same image; measured slower than engine 1 so far
    // (C2 38.3 vs 31.2 ms: 30 of 32 lanes inside the BVH loop instead of 23, but more L1 wavefronts and 20 warps per SM)
    bool pool_engine = false;
    if (knob("TRT_ENGINE", 0) == 3 && d.scene.n_tris > 0 && !brute && p->nbRebondMax >= 1 && p->nbRebondMax <= 8191 && spp < (1 << 26)) {
        A.pp_slots = knob("TRT_TUNE_SLOTS_PER_WARP", 64);
        if (A.pp_slots < 32) A.pp_slots = 32;
        if (A.pp_slots > 64) A.pp_slots = 64;
        A.pp_shade_at = knob("TRT_TUNE_SHADE_AT", A.pp_slots >= 64 ? 32 : A.pp_slots - 24);
        if (A.pp_shade_at < 1) A.pp_shade_at = 1;
        if (A.pp_shade_at > 32) A.pp_shade_at = 32;
        A.pp_refill_min = knob("TRT_TUNE_REFILL_MIN", 8);
        A.pp_exit_below = knob("TRT_TUNE_EXIT", 10);
        if (A.pp_exit_below < 1) A.pp_exit_below = 1;
        auto magic = [](unsigned d) { unsigned long long m = (1ULL << 32) / d; return (unsigned)(m > 0xffffffffULL ? 0xffffffffULL : m); };
        A.pp_per_tile = 32u * (unsigned)spp;
        A.pp_m_per_tile = magic(A.pp_per_tile); A.pp_m_spp = magic((unsigned)spp);
        A.n_items = my_tiles * 32 * (long long)spp;                       // paths, edge-tile padding included
        A.pool = knob("TRT_TUNE_POOL", spp < 256 ? 256 : (spp > 2048 ? 2048 : spp));
        A.pp_fit32 = (A.n_items + 2 * (long long)A.pool < (1LL << 32)) ? 1 : 0;
        pool_engine = trt_pool_smem_bytes(A) <= (size_t)200 * 1024 / 4;   // 4+ CTAs of 128 threads per SM must fit
    }
    // engine 4 (opt-in, TRT_ENGINE=4): the primary rays of the frame as packets first (primary_packet.cu), then K2 starts
    // every path from its stored hit
    bool packet_engine = false;
    if (knob("TRT_ENGINE", 0) == 4 && d.scene.n_tris > 0 && !brute && A.defer > 0 && A.top_nodes == 0 && p->nbRebondMax >= 1 && spp < (1 << 26)) {
        const size_t npix = (size_t)A.W * (size_t)A.H, bytes = npix * (size_t)spp * sizeof(float4);
        if (bytes <= ((size_t)knob("TRT_TUNE_PREHIT_GB", 8) << 30) && d.prehit.ensure(bytes) == cudaSuccess) {
            auto magic = [](unsigned d) { unsigned long long m = (1ULL << 32) / d; return (unsigned)(m > 0xffffffffULL ? 0xffffffffULL : m); };
            RenderArgs K = A;
            K.pp_per_tile = 32u * (unsigned)spp;
            K.pp_m_per_tile = magic(K.pp_per_tile); K.pp_m_spp = magic((unsigned)spp);
            K.n_items = my_tiles * 32 * (long long)spp;
            K.pp_fit32 = (K.n_items + 64 < (1LL << 32)) ? 1 : 0;
            K.pool = knob("TRT_TUNE_PK_POOL", 8);                       // packets per reservation
            K.item_counter = d.ctrs + 15;
            CU(ctx, cudaEventRecord(d.ev_a, d.stream));
            CU(ctx, trt_launch_primary_packets(K, d.n_sms, (float4*)d.prehit.p, spp, count_tests, d.stream));
            A.prehit = (const float4*)d.prehit.p; A.prehit_slab = spp;
            packet_engine = true;
        } else cudaGetLastError();
    }
    if (packet_engine) {
        CU(ctx, trt_launch_pathtrace(A, d.n_sms, blocks_per_sm, brute, count_tests, d.stream));
        CU(ctx, cudaEventRecord(d.ev_b, d.stream));
        d.last_launches = 2;
    } else if (pool_engine) {
        CU(ctx, cudaEventRecord(d.ev_a, d.stream));
        CU(ctx, trt_launch_pathtrace_pool(A, d.n_sms, blocks_per_sm, count_tests, d.stream));
        CU(ctx, cudaEventRecord(d.ev_b, d.stream));
        d.last_launches = 1;
    } else {
        CU(ctx, cudaEventRecord(d.ev_a, d.stream));
        CU(ctx, trt_launch_pathtrace(A, d.n_sms, blocks_per_sm, brute, count_tests, d.stream));
        CU(ctx, cudaEventRecord(d.ev_b, d.stream));
        d.last_launches = 1;
    }
    CU(ctx, cudaMemcpyAsync(d.h_ctrs, d.ctrs, sizeof(unsigned long long) * 15, cudaMemcpyDeviceToHost, d.stream));
    return TRT_OK;
}

void collect_stats(trt_ctx* ctx, const trt_params* p, trt_stats* st, bool traced)
{
    if (!st) return;
    if (traced) {
        st->rays = st->box_tests = st->tri_tests = st->sphere_tests = st->rays_skipped = 0;
        st->ms_trace = 0.0;
        for (auto& d : ctx->devs) {
            float ms = 0.0f;
            cudaSetDevice(d.dev);
            if (cudaEventElapsedTime(&ms, d.ev_a, d.ev_b) == cudaSuccess && ms > st->ms_trace) st->ms_trace = ms;
            st->rays += d.h_ctrs[1]; st->box_tests += d.h_ctrs[2]; st->tri_tests += d.h_ctrs[3]; st->sphere_tests += d.h_ctrs[4];
            st->rays_skipped += d.h_ctrs[12];
        }
        st->paths = (uint64_t)p->largeur_image * (uint64_t)p->hauteur_image * (uint64_t)p->nbRayonParPixel;
    }
}

// pathtrace_pool.cu raises counters[13] when a warp has been scheduling for 20 s without finishing (a bug, not a long frame)
int check_watchdog(trt_ctx* ctx)
{
    for (auto& d : ctx->devs)
        if (d.h_ctrs[14] != 0) return fail(ctx, TRT_ERR_CUDA, "path-tracing kernel gave up after 20 s on device %d (scheduler watchdog)", d.dev);
#ifdef TRT_DEBUG_BOUNDS
    for (auto& d : ctx->devs)
        if (d.h_ctrs[13] != 0) return fail(ctx, TRT_ERR_CUDA, "TRT_DEBUG_BOUNDS: index out of range on device %d, check %llu (trt_isect.cuh / pathtrace*.cu)", d.dev, (unsigned long long)d.h_ctrs[13]);
#endif
    return TRT_OK;
}

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// No C++ exception may cross the C ABI (std::bad_alloc from a host vector, std::system_error from a thread): the entry points
// that allocate on the host run their body through this guard.
#define TRT_GUARDED(ctx, expr)                                                                       \
    try { return (expr); }                                                                           \
    catch (const std::bad_alloc&) { return fail(ctx, TRT_ERR_OOM, "out of host memory"); }           \
    catch (const std::exception& e) { return fail(ctx, TRT_ERR_INVALID, "internal error: %s", e.what()); } \
    catch (...) { return fail(ctx, TRT_ERR_INVALID, "internal error"); }

// error plumbing for the other translation units of the library (nccl_glue.cu)
int trt_set_error(trt_ctx* ctx, int code, const char* text) { return fail(ctx, code, "%s", text ? text : ""); }

// ================================================================================================
extern "C" {

TRT_API int trt_host_alloc(size_t bytes, void** out)
{
    if (!out) return TRT_ERR_INVALID;
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { *out = nullptr; return fail(nullptr, e == cudaErrorMemoryAllocation ? TRT_ERR_OOM : TRT_ERR_CUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e)); }
    return TRT_OK;
}

TRT_API void trt_host_free(void* p) { if (p) cudaFreeHost(p); }

TRT_API int trt_device_of(const trt_ctx* ctx, int k) { return (ctx && k >= 0 && k < (int)ctx->devs.size()) ? ctx->devs[k].dev : -1; }

TRT_API int trt_version(void) { return 100; }

TRT_API const char* trt_last_error(const trt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

TRT_API int trt_default_overrides(trt_mat_override* out, int capacity)
{
    if (!out || capacity < 3) return 3;
    memset(out, 0, sizeof(trt_mat_override) * 3);
    out[0].material_ordinal = 1; out[0].mask = TRT_OV_EMISSION_COLOR | TRT_OV_EMISSION_STRENGTH | TRT_OV_ALPHA;
    out[0].emissionColor[0] = out[0].emissionColor[1] = out[0].emissionColor[2] = 1.0;
    out[0].emissionStrength = 1.85; out[0].alpha = 1.0;
    out[1].material_ordinal = 4; out[1].mask = TRT_OV_ALPHA | TRT_OV_MATERIAL_INDEX | TRT_OV_REFLECTION;
    out[1].alpha = 0.6; out[1].materialIndex = 1.33; out[1].reflectionStrength = 0.93;
    out[2].material_ordinal = 3; out[2].mask = TRT_OV_ALPHA | TRT_OV_MATERIAL_INDEX | TRT_OV_REFLECTION;
    out[2].alpha = 0.1; out[2].materialIndex = 1.50; out[2].reflectionStrength = 0.3;
    return 3;
}

TRT_API int trt_init(const int* device_ids, int n_devices, trt_ctx** out)
{
    if (!out) return fail(nullptr, TRT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, TRT_ERR_NO_DEVICE, "no CUDA device (%s); libtipe_rt has no CPU fallback", cudaGetErrorString(e));
    if (n_devices < 1) n_devices = 1;
    if (n_devices > TRT_MAX_RESOLVE_SOURCES) return fail(nullptr, TRT_ERR_INVALID, "at most %d devices", TRT_MAX_RESOLVE_SOURCES);
    trt_ctx* ctx = new (std::nothrow) trt_ctx();
    if (!ctx) return fail(nullptr, TRT_ERR_OOM, "out of host memory");
    for (int k = 0; k < n_devices; k++) {
        DeviceCtx d;
        d.dev = device_ids ? device_ids[k] : k;
        if (d.dev < 0 || d.dev >= count) { delete ctx; return fail(nullptr, TRT_ERR_NO_DEVICE, "device %d not present (%d visible)", d.dev, count); }
        ctx->devs.push_back(d);
    }
    for (auto& d : ctx->devs) {
        cudaDeviceProp prop;
        if ((e = cudaSetDevice(d.dev)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, d.dev)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreate(&d.ev_a)) != cudaSuccess || (e = cudaEventCreate(&d.ev_b)) != cudaSuccess ||
            (e = cudaEventCreate(&d.ev_c)) != cudaSuccess || (e = cudaEventCreate(&d.ev_d)) != cudaSuccess ||
            (e = cudaMalloc(&d.ctrs, sizeof(unsigned long long) * 16)) != cudaSuccess ||
            (e = cudaMallocHost(&d.h_ctrs, sizeof(unsigned long long) * 16)) != cudaSuccess) {
            fail(nullptr, TRT_ERR_CUDA, "device %d set-up failed: %s", d.dev, cudaGetErrorString(e));
            trt_destroy(ctx);
            return TRT_ERR_CUDA;
        }
        if (prop.major < 10) {
            fail(nullptr, TRT_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", d.dev, prop.major, prop.minor);
            trt_destroy(ctx);
            return TRT_ERR_NO_DEVICE;
        }
        d.n_sms = prop.multiProcessorCount;
        memset(d.h_ctrs, 0, sizeof(unsigned long long) * 16);
    }
    // peer access between every pair of the ctx's devices: every device resolves its own band of the image from the sums of all of
    // them (trt_render), and the built scene goes from the first device to the others over NVLink (trt_scene_set)
    for (size_t i = 0; i < ctx->devs.size(); i++)
        for (size_t j = 0; j < ctx->devs.size(); j++) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, ctx->devs[i].dev, ctx->devs[j].dev);
            if (!can) { fail(nullptr, TRT_ERR_CUDA, "device %d cannot map device %d's memory (no NVLink/P2P)", ctx->devs[i].dev, ctx->devs[j].dev); trt_destroy(ctx); return TRT_ERR_CUDA; }
            cudaSetDevice(ctx->devs[i].dev);
            e = cudaDeviceEnablePeerAccess(ctx->devs[j].dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { fail(nullptr, TRT_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); trt_destroy(ctx); return TRT_ERR_CUDA; }
            cudaGetLastError();
        }
    *out = ctx;
    return TRT_OK;
}

TRT_API void trt_destroy(trt_ctx* ctx)
{
    if (!ctx) return;
    if (!ctx->devs.empty()) {
        cudaSetDevice(ctx->devs[0].dev);
        UploadRing& r = ctx->ring;
        if (r.stream) { cudaStreamSynchronize(r.stream); cudaStreamDestroy(r.stream); }
        for (cudaEvent_t ev : r.ev) if (ev) cudaEventDestroy(ev);
        if (r.done) cudaEventDestroy(r.done);
        if (r.h) cudaFreeHost(r.h);
    }
    for (auto& d : ctx->devs) {
        cudaSetDevice(d.dev);
        if (d.stream) cudaStreamSynchronize(d.stream);
        free_scene(d);
        d.sky.release();
        d.prehit.release();
        d.nodes64.release();
        if (d.accum) cudaFree(d.accum);
        if (d.out) cudaFree(d.out);
        if (d.h_stage) cudaFreeHost(d.h_stage);
        for (cudaEvent_t ev : d.chunk_ev) cudaEventDestroy(ev);
        if (d.ctrs) cudaFree(d.ctrs);
        if (d.h_ctrs) cudaFreeHost(d.h_ctrs);
        cudaEvent_t evs[] = {d.ev_a, d.ev_b, d.ev_c, d.ev_d};
        for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete ctx;
}

TRT_API void* trt_stream(trt_ctx* ctx) { return ctx && !ctx->devs.empty() ? (void*)ctx->devs[0].stream : nullptr; }

TRT_API int trt_synchronize(trt_ctx* ctx)
{
    if (!ctx) return TRT_ERR_INVALID;
    for (auto& d : ctx->devs) { CU(ctx, cudaSetDevice(d.dev)); CU(ctx, cudaStreamSynchronize(d.stream)); }
    return TRT_OK;
}

static int trt_scene_set_impl(trt_ctx* ctx,
                          const trt_sphere* sphere_list, int nbSpheres,
                          const trt_triangle* triangle_list, int nbTriangles,
                          const trt_material* mat_list, int tex_width, int tex_height, int nbMaterials,
                          const int* quelMatPourTri,
                          const trt_mat_override* overrides, int n_overrides)
{
    if (!ctx) return TRT_ERR_INVALID;
    if (nbSpheres < 0 || nbTriangles < 0) return fail(ctx, TRT_ERR_INVALID, "negative primitive count");
    if (nbSpheres > 0 && !sphere_list) return fail(ctx, TRT_ERR_INVALID, "sphere_list is NULL");
    if (nbSpheres > TRT_MAX_SMEM_SPHERES) return fail(ctx, TRT_ERR_INVALID, "at most %d spheres (they live in shared memory)", TRT_MAX_SMEM_SPHERES);
    if (nbTriangles > 0) {
        if (!triangle_list || !quelMatPourTri) return fail(ctx, TRT_ERR_INVALID, "triangle_list / quelMatPourTri is NULL");
        if (!mat_list || tex_width < 1 || tex_height < 1 || nbMaterials < 1)
            return fail(ctx, TRT_ERR_INVALID, "triangles need mat_list (the reference textures every triangle, main.c:86)");
        for (int k = 0; k < nbTriangles; k++)
            if (quelMatPourTri[k] < 0 || quelMatPourTri[k] >= nbMaterials)
                return fail(ctx, TRT_ERR_INVALID, "quelMatPourTri[%d] = %d outside 0..%d", k, quelMatPourTri[k], nbMaterials - 1);
    } else { tex_width = tex_height = nbMaterials = 0; }

    trt_mat_override builtin[3];
    if (!overrides && n_overrides < 0) { n_overrides = trt_default_overrides(builtin, 3); overrides = builtin; }
    if (n_overrides < 0 || (n_overrides > 0 && !overrides)) return fail(ctx, TRT_ERR_INVALID, "bad overrides");

    const size_t n_texels = (size_t)nbMaterials * (size_t)tex_width * (size_t)tex_height;
    ctx->has_scene = false;
    const double t0 = now_ms();
    const size_t ns = (size_t)(nbSpheres > 0 ? nbSpheres : 1), nt = (size_t)(nbTriangles > 0 ? nbTriangles : 1);
    const size_t tex_bytes = sizeof(float4) * 3 * (n_texels > 0 ? n_texels : 1);
    for (auto& d : ctx->devs) {            // the previous frame is over before its scene is overwritten
        CU(ctx, cudaSetDevice(d.dev));
        CU(ctx, cudaStreamSynchronize(d.stream));
    }

    // ---- root device: upload the caller's arrays, convert, build the BVH ----------------------------------
    DeviceCtx& d = ctx->devs[0];
    CU(ctx, cudaSetDevice(d.dev));
    CU(ctx, d.sph_raw.ensure(sizeof(trt_sphere) * ns));
    CU(ctx, d.sph.ensure(sizeof(float4) * ns));
    CU(ctx, d.sph_mat.ensure(sizeof(float4) * 3 * ns));
    CU(ctx, d.tri_raw.ensure(sizeof(trt_triangle) * nt));
    CU(ctx, d.quel_mat.ensure(sizeof(int) * nt));
    CU(ctx, d.tri.ensure(sizeof(float4) * 4 * nt + 256));
    CU(ctx, d.tri_uv.ensure(sizeof(float4) * 2 * nt));
    CU(ctx, d.tri_orig.ensure(sizeof(int) * nt));
    CU(ctx, d.nodes.ensure(sizeof(float4) * 8 * nt));
    CU(ctx, d.texels.ensure(tex_bytes));
    CU(ctx, d.scratch.ensure(trt_build_scratch_bytes(nbTriangles)));
    if (nbSpheres) CU(ctx, cudaMemcpyAsync(d.sph_raw.p, sphere_list, sizeof(trt_sphere) * nbSpheres, cudaMemcpyHostToDevice, d.stream));
    if (nbTriangles) {
        // texel slab: 80 B of FP64 per texel in the caller's array (hitinfo.h:6-13), packed to 10 floats on the way: half the
        // bytes over PCIe, and out of pinned memory (C4, 1024^2 texels: 84 MB pageable took 8 ms)
        CU(ctx, d.mats_raw.ensure(sizeof(float) * 10 * n_texels));
        CU(ctx, d.ov_raw.ensure(sizeof(trt_mat_override) * (size_t)(n_overrides > 0 ? n_overrides : 1)));
        auto pack_texels = [](char* dst, const char* src, size_t count) {
            const double* in = (const double*)src; float* out = (float*)dst;
            for (size_t k = 0; k < 10 * count; k++) out[k] = (float)in[k];
        };
        auto pack_copy_tri = [](char* dst, const char* src, size_t count) { memcpy(dst, src, count * sizeof(trt_triangle)); };
        const bool staged = getenv("TRT_TUNE_STAGED_UP") == nullptr || atoi(getenv("TRT_TUNE_STAGED_UP")) != 0;
        int rc;
        if (staged && sizeof(trt_triangle) * (size_t)nbTriangles >= STAGED_MIN_BYTES) {
            if ((rc = staged_upload(ctx, (char*)d.tri_raw.p, (const char*)triangle_list, (size_t)nbTriangles, sizeof(trt_triangle), sizeof(trt_triangle), pack_copy_tri)) != TRT_OK) return rc;
        } else {
            CU(ctx, cudaMemcpyAsync(d.tri_raw.p, triangle_list, sizeof(trt_triangle) * nbTriangles, cudaMemcpyHostToDevice, d.stream));
        }
        CU(ctx, cudaMemcpyAsync(d.quel_mat.p, quelMatPourTri, sizeof(int) * nbTriangles, cudaMemcpyHostToDevice, d.stream));
        if (staged && sizeof(trt_material) * n_texels >= STAGED_MIN_BYTES) {
            if ((rc = staged_upload(ctx, (char*)d.mats_raw.p, (const char*)mat_list, n_texels, sizeof(trt_material), sizeof(float) * 10, pack_texels)) != TRT_OK) return rc;
        } else {
            std::vector<float> packed(10 * n_texels);
            pack_texels((char*)packed.data(), (const char*)mat_list, n_texels);
            CU(ctx, cudaMemcpyAsync(d.mats_raw.p, packed.data(), sizeof(float) * 10 * n_texels, cudaMemcpyHostToDevice, d.stream));
            CU(ctx, cudaStreamSynchronize(d.stream));            // `packed` goes out of scope
        }
        if (n_overrides > 0) CU(ctx, cudaMemcpyAsync(d.ov_raw.p, overrides, sizeof(trt_mat_override) * n_overrides, cudaMemcpyHostToDevice, d.stream));
    }
    BuildInputs in{};
    in.spheres_raw = (const double*)d.sph_raw.p; in.n_spheres = nbSpheres;
    in.tris_raw = (const double*)d.tri_raw.p; in.n_tris = nbTriangles;
    in.mats_packed = (const float*)d.mats_raw.p; in.tex_w = tex_width; in.tex_h = tex_height; in.n_materials = nbMaterials;
    in.quel_mat = (const int*)d.quel_mat.p; in.overrides_raw = (const double*)d.ov_raw.p; in.n_overrides = nbTriangles ? n_overrides : 0;
    in.scratch = d.scratch.p; in.scratch_bytes = d.scratch.cap;
    BuildOutputs bo{};
    bo.sph = (float4*)d.sph.p; bo.sph_mat = (float4*)d.sph_mat.p; bo.tri = (float4*)d.tri.p; bo.tri_uv = (float4*)d.tri_uv.p;
    bo.tri_orig = (int*)d.tri_orig.p; bo.nodes = (float4*)d.nodes.p; bo.texels = (float4*)d.texels.p;
    CU(ctx, trt_build_scene(in, bo, d.stream));
    const bool node64 = getenv("TRT_TUNE_NODE64") && atoi(getenv("TRT_TUNE_NODE64")) != 0 && bo.n_wide_nodes > 0;
    if (node64) {
        CU(ctx, d.nodes64.ensure((size_t)bo.n_wide_nodes * 64));
        CU(ctx, trt_compress_nodes(bo.n_wide_nodes, bo.nodes, d.nodes64.p, d.stream));
    }
    CU(ctx, cudaEventRecord(d.ev_c, d.stream));
    if (3 * bo.max_depth + 2 > TRT_STACK_DEPTH)
        return fail(ctx, TRT_ERR_INVALID, "BVH depth %d needs a traversal stack of %d entries (limit %d)", bo.max_depth, 3 * bo.max_depth + 2, TRT_STACK_DEPTH);
    if (d.mats_raw.cap > ((size_t)256 << 20)) { CU(ctx, cudaStreamSynchronize(d.stream)); d.mats_raw.release(); }      // big texture slabs are not kept around
    {
        DevScene& s = d.scene;
        const float4* sky = s.sky; const int sky_w = s.sky_w, sky_h = s.sky_h;             // trt_sky_set is independent of the scene
        s = DevScene{};
        s.sky = sky; s.sky_w = sky_w; s.sky_h = sky_h;
        s.n_spheres = nbSpheres; s.sph = bo.sph; s.sph_mat = bo.sph_mat; s.sph64 = (const double*)d.sph_raw.p;
        s.n_tris = nbTriangles; s.tri = bo.tri; s.tri_uv = bo.tri_uv; s.tri_orig = bo.tri_orig; s.tri64 = (const double*)d.tri_raw.p;
        s.quel_mat = (const int*)d.quel_mat.p; s.nodes = bo.nodes; s.root = bo.root; s.n_nodes = bo.n_wide_nodes;
        s.nodes64 = node64 ? (const float4*)d.nodes64.p : nullptr;
        s.texels = bo.texels; s.tex_w = tex_width; s.tex_h = tex_height; s.n_materials = nbMaterials;
        s.tri_tmin = bo.tri_tmin;
        s.root_lo = bo.root_lo; s.root_hi = bo.root_hi;
    }
    ctx->info.n_nodes = bo.n_wide_nodes;
    ctx->info.n_leaves = nbTriangles;
    ctx->info.max_depth = bo.max_depth; ctx->info.n_twins = bo.n_twins; ctx->info.sah_cost = bo.sah_cost;

    // ---- the other devices of the ctx get the BUILT scene over NVLink: one build, N - 1 peer copies in parallel ----------
    const size_t n_nodes = (size_t)(bo.n_wide_nodes > 0 ? bo.n_wide_nodes : 1);
    for (size_t k = 1; k < ctx->devs.size(); k++) {
        DeviceCtx& q = ctx->devs[k];
        CU(ctx, cudaSetDevice(q.dev));
        CU(ctx, q.sph.ensure(sizeof(float4) * ns));
        CU(ctx, q.sph_mat.ensure(sizeof(float4) * 3 * ns));
        CU(ctx, q.tri.ensure(sizeof(float4) * 4 * nt + 256));
        CU(ctx, q.tri_uv.ensure(sizeof(float4) * 2 * nt));
        CU(ctx, q.tri_orig.ensure(sizeof(int) * nt));
        CU(ctx, q.nodes.ensure(sizeof(float4) * 8 * n_nodes));
        CU(ctx, q.texels.ensure(tex_bytes));
        CU(ctx, cudaStreamWaitEvent(q.stream, d.ev_c, 0));
        struct { Buf* dst; Buf* src; size_t bytes; } copies[] = {
            {&q.sph, &d.sph, sizeof(float4) * (size_t)nbSpheres}, {&q.sph_mat, &d.sph_mat, sizeof(float4) * 3 * (size_t)nbSpheres},
            {&q.tri, &d.tri, sizeof(float4) * 4 * (size_t)nbTriangles}, {&q.tri_uv, &d.tri_uv, sizeof(float4) * 2 * (size_t)nbTriangles},
            {&q.tri_orig, &d.tri_orig, sizeof(int) * (size_t)nbTriangles}, {&q.nodes, &d.nodes, nbTriangles ? sizeof(float4) * 8 * n_nodes : 0},
            {&q.texels, &d.texels, sizeof(float4) * 3 * n_texels},
        };
        for (auto& c : copies)
            if (c.bytes) CU(ctx, cudaMemcpyPeerAsync(c.dst->p, q.dev, c.src->p, d.dev, c.bytes, q.stream));
        DevScene& s = q.scene;
        const float4* sky = s.sky; const int sky_w = s.sky_w, sky_h = s.sky_h;
        s = d.scene;
        s.sky = sky; s.sky_w = sky_w; s.sky_h = sky_h;
        s.sph = (const float4*)q.sph.p; s.sph_mat = (const float4*)q.sph_mat.p; s.tri = (const float4*)q.tri.p; s.tri_uv = (const float4*)q.tri_uv.p;
        s.tri_orig = (const int*)q.tri_orig.p; s.nodes = (const float4*)q.nodes.p; s.texels = (const float4*)q.texels.p;
        s.sph64 = nullptr; s.tri64 = nullptr; s.quel_mat = nullptr;       // the FP64 primary pass (trt_primary) runs on the root only
        s.nodes64 = nullptr;
    }
    for (auto& q : ctx->devs) { CU(ctx, cudaSetDevice(q.dev)); CU(ctx, cudaStreamSynchronize(q.stream)); }
    ctx->info.ms_build = now_ms() - t0;
    ctx->has_scene = true;
    return TRT_OK;
}

static int trt_sky_set_impl(trt_ctx* ctx, const trt_material* sky_mat_list, int sky_width, int sky_height)
{
    if (!ctx) return TRT_ERR_INVALID;
    if (sky_mat_list && (sky_width < 1 || sky_height < 1 || (long long)sky_width * sky_height > (1LL << 30)))
        return fail(ctx, TRT_ERR_INVALID, "bad sky size %d x %d", sky_width, sky_height);
    const size_t n = sky_mat_list ? (size_t)sky_width * (size_t)sky_height : 0;
    std::vector<float4> rgb(n);
    for (size_t k = 0; k < n; k++)
        rgb[k] = make_float4((float)sky_mat_list[k].diffuseColor.e[0], (float)sky_mat_list[k].diffuseColor.e[1],
                             (float)sky_mat_list[k].diffuseColor.e[2], 0.0f);
    for (auto& d : ctx->devs) {
        CU(ctx, cudaSetDevice(d.dev));
        CU(ctx, cudaStreamSynchronize(d.stream));
        if (n == 0) { d.scene.sky = nullptr; d.scene.sky_w = d.scene.sky_h = 0; continue; }
        CU(ctx, d.sky.ensure(sizeof(float4) * n));
        CU(ctx, cudaMemcpyAsync(d.sky.p, rgb.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, d.stream));
        CU(ctx, cudaStreamSynchronize(d.stream));
        d.scene.sky = (const float4*)d.sky.p; d.scene.sky_w = sky_width; d.scene.sky_h = sky_height;
    }
    return TRT_OK;
}

TRT_API int trt_bvh_info_get(trt_ctx* ctx, trt_bvh_info* out)
{
    if (!ctx || !out) return TRT_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, TRT_ERR_NO_SCENE, "trt_scene_set has not been called");
    *out = ctx->info;
    return TRT_OK;
}

// Schedule counters of the last TRT_FLAG_COUNT_TESTS launch on device 0 (tools/ only; see pathtrace.cu):
// [0] BVH-loop trips (per warp) [1] lanes inside, summed over trips [2] lanes at a node [3] lanes at a leaf
// [4] loop sessions [5] housekeeping passes [6] lanes consuming a hit, summed over passes
TRT_API int trt_debug_schedule(trt_ctx* ctx, uint64_t* out7)
{
    if (!ctx || !out7) return TRT_ERR_INVALID;
    for (int k = 0; k < 7; k++) out7[k] = ctx->devs[0].h_ctrs[5 + k];
    return TRT_OK;
}

TRT_API int trt_accumulate(trt_ctx* ctx, const trt_params* params, int64_t* d_accum, trt_stats* stats)
{
    int rc = check_params(ctx, params);
    if (rc != TRT_OK) return rc;
    if (!d_accum) return fail(ctx, TRT_ERR_INVALID, "d_accum is NULL");
    if (ctx->devs.size() != 1) return fail(ctx, TRT_ERR_INVALID, "trt_accumulate works on a single-device ctx (one process per GPU); use trt_render for in-process multi-GPU");
    const double t0 = now_ms();
    DeviceCtx& d = ctx->devs[0];
    CU(ctx, cudaSetDevice(d.dev));
    rc = launch_trace(ctx, d, params, params->nbRayonParPixel, params->sample_offset, (long long*)d_accum,
                      params->tile_part ? (params->tile_part & 0xffff) : 0, params->tile_part ? ((params->tile_part >> 16) & 0xffff) : 1);
    if (rc != TRT_OK) return rc;
    CU(ctx, cudaStreamSynchronize(d.stream));
    if ((rc = check_watchdog(ctx)) != TRT_OK) return rc;
    if (stats) {
        memset(stats, 0, sizeof *stats);
        collect_stats(ctx, params, stats, true);
        stats->kernel_launches = d.last_launches;
        stats->ms_total = now_ms() - t0;
    }
    return TRT_OK;
}

static int resolve_common(trt_ctx* ctx, const trt_params* p, const long long* const* srcs, int n_src,
                          double* d_canva, double* d_albedo, double* d_normal, unsigned char* d_canva8 = nullptr,
                          DeviceCtx* dev = nullptr, size_t p0 = 0, size_t p1 = 0)
{
    DeviceCtx& d = dev ? *dev : ctx->devs[0];
    ResolveArgs R{};
    R.W = p->largeur_image; R.H = p->hauteur_image;
    R.divisor = p->sample_total > 0 ? p->sample_total : p->nbRayonParPixel;
    R.n_src = n_src;
    for (int k = 0; k < n_src; k++) R.src[k] = srcs[k];
    R.p0 = p0; R.p1 = p1 > p0 ? p1 : (size_t)R.W * (size_t)R.H;
    R.canva = d_canva; R.albedo = d_albedo; R.normal = d_normal; R.canva8 = d_canva8;
    CU(ctx, trt_launch_resolve(R, d.n_sms, d.stream));
    return TRT_OK;
}

// Is this host pointer page-locked (trt_host_alloc, cudaHostAlloc, cudaHostRegister)?  Then the DMA engine can write it directly.
static bool host_is_pinned(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// Tonemap pixels [p0, p1) on device d and start sending them to the caller's arrays (pointers already offset to p0).
//   mode 2: all three arrays are page-locked: canva is written as doubles and the three bands are copied straight into the caller's
//           memory by the DMA engine (no staging, no host threads: with N devices the N bands cross N PCIe links at once);
//   mode 1: pageable arrays, large: canva as bytes + AOV tables through the pinned staging ring (copy_out_issue / copy_out_drain);
//   mode 0: pageable arrays, small: plain copies in out_finish.
static int resolve_and_send(trt_ctx* ctx, const trt_params* rp, const long long* const* srcs, int n_src, DeviceCtx& d, size_t p0, size_t p1,
                            trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab, int* mode)
{
    const size_t nb = p1 - p0;
    int rc;
    *mode = 0;
    if (nb == 0) return TRT_OK;
    const char* v = getenv("TRT_TUNE_DIRECT_D2H");
    const bool direct = !(v && *v && atoi(v) == 0) && host_is_pinned(canva) && host_is_pinned(albedo_tab) && host_is_pinned(normal_tab);
    if (direct) {
        double* ob = d.out;
        if ((rc = resolve_common(ctx, rp, srcs, n_src, ob, ob + 3 * nb, ob + 6 * nb, nullptr, &d, p0, p1)) != TRT_OK) return rc;
        CU(ctx, cudaEventRecord(d.ev_d, d.stream));
        const size_t part = sizeof(double) * 3 * nb;
        CU(ctx, cudaMemcpyAsync(canva, ob, part, cudaMemcpyDeviceToHost, d.stream));
        CU(ctx, cudaMemcpyAsync(albedo_tab, ob + 3 * nb, part, cudaMemcpyDeviceToHost, d.stream));
        CU(ctx, cudaMemcpyAsync(normal_tab, ob + 6 * nb, part, cudaMemcpyDeviceToHost, d.stream));
        *mode = 2;
        return TRT_OK;
    }
    char* ob = (char*)d.out;
    if ((rc = resolve_common(ctx, rp, srcs, n_src, nullptr, (double*)(ob + canva8_bytes(nb)), (double*)(ob + canva8_bytes(nb)) + 3 * nb,
                             (unsigned char*)ob, &d, p0, p1)) != TRT_OK) return rc;
    CU(ctx, cudaEventRecord(d.ev_d, d.stream));
    rc = TRT_OK;
    *mode = copy_out_issue(ctx, d, nb, &rc) ? 1 : 0;
    return rc;
}

static int out_finish(trt_ctx* ctx, DeviceCtx& d, size_t nb, int mode, trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab)
{
    const OutJob job = {&d, nb, mode, canva, albedo_tab, normal_tab};
    return drain_jobs(ctx, &job, 1);
}

TRT_API int trt_resolve_device(trt_ctx* ctx, const trt_params* params, const int64_t* d_accum,
                               double* d_canva, double* d_albedo, double* d_normal)
{
    int rc = check_params(ctx, params);
    if (rc != TRT_OK) return rc;
    if (!d_accum || !d_canva || !d_albedo || !d_normal) return fail(ctx, TRT_ERR_INVALID, "NULL device buffer");
    CU(ctx, cudaSetDevice(ctx->devs[0].dev));
    const long long* src = (const long long*)d_accum;
    rc = resolve_common(ctx, params, &src, 1, d_canva, d_albedo, d_normal);
    if (rc != TRT_OK) return rc;
    CU(ctx, cudaStreamSynchronize(ctx->devs[0].stream));
    return TRT_OK;
}

static int trt_resolve_impl(trt_ctx* ctx, const trt_params* params, const int64_t* d_accum,
                        trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab)
{
    int rc = check_params(ctx, params);
    if (rc != TRT_OK) return rc;
    if (!d_accum) return fail(ctx, TRT_ERR_INVALID, "d_accum is NULL");
    DeviceCtx& d = ctx->devs[0];
    const size_t npix = (size_t)params->largeur_image * params->hauteur_image;
    CU(ctx, cudaSetDevice(d.dev));
    if ((rc = ensure_out(ctx, d, npix)) != TRT_OK) return rc;
    const long long* src = (const long long*)d_accum;
    if (!(host_is_pinned(canva) && host_is_pinned(albedo_tab) && host_is_pinned(normal_tab))) ensure_stage(d, sizeof(double) * 9 * npix);
    int mode = 0;
    if ((rc = resolve_and_send(ctx, params, &src, 1, d, 0, npix, canva, albedo_tab, normal_tab, &mode)) != TRT_OK) return rc;
    return out_finish(ctx, d, npix, mode, canva, albedo_tab, normal_tab);
}

TRT_API int trt_resolve_hdr_device(trt_ctx* ctx, const trt_params* params, const int64_t* d_accum,
                                   float* d_radiance, float* d_albedo, float* d_normal)
{
    int rc = check_params(ctx, params);
    if (rc != TRT_OK) return rc;
    if (!d_accum || !d_radiance || !d_albedo || !d_normal) return fail(ctx, TRT_ERR_INVALID, "NULL device buffer");
    DeviceCtx& d = ctx->devs[0];
    const size_t npix = (size_t)params->largeur_image * params->hauteur_image;
    CU(ctx, cudaSetDevice(d.dev));
    const int divisor = params->sample_total > 0 ? params->sample_total : params->nbRayonParPixel;
    CU(ctx, trt_launch_resolve_hdr((int)npix, divisor, (const long long*)d_accum, d_radiance, d_albedo, d_normal, d.n_sms, d.stream));
    CU(ctx, cudaStreamSynchronize(d.stream));
    return TRT_OK;
}

TRT_API int trt_resolve_hdr(trt_ctx* ctx, const trt_params* params, const int64_t* d_accum,
                            float* radiance, float* albedo, float* normal)
{
    int rc = check_params(ctx, params);
    if (rc != TRT_OK) return rc;
    if (!d_accum || !radiance || !albedo || !normal) return fail(ctx, TRT_ERR_INVALID, "NULL buffer");
    DeviceCtx& d = ctx->devs[0];
    const size_t npix = (size_t)params->largeur_image * params->hauteur_image;
    CU(ctx, cudaSetDevice(d.dev));
    if ((rc = ensure_out(ctx, d, npix)) != TRT_OK) return rc;
    float* buf = (float*)d.out;                    // 9*npix doubles are room for 9*npix floats
    if ((rc = trt_resolve_hdr_device(ctx, params, d_accum, buf, buf + 3 * npix, buf + 6 * npix)) != TRT_OK) return rc;
    CU(ctx, cudaMemcpyAsync(radiance, buf, sizeof(float) * 3 * npix, cudaMemcpyDeviceToHost, d.stream));
    CU(ctx, cudaMemcpyAsync(albedo, buf + 3 * npix, sizeof(float) * 3 * npix, cudaMemcpyDeviceToHost, d.stream));
    CU(ctx, cudaMemcpyAsync(normal, buf + 6 * npix, sizeof(float) * 3 * npix, cudaMemcpyDeviceToHost, d.stream));
    CU(ctx, cudaStreamSynchronize(d.stream));
    return TRT_OK;
}

static int trt_render_impl(trt_ctx* ctx, const trt_params* params,
                       trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab, trt_stats* stats)
{
    int rc = check_params(ctx, params);
    if (rc != TRT_OK) return rc;
    const double t0 = now_ms();
    const size_t npix = (size_t)params->largeur_image * params->hauteur_image;
    const int n_dev = (int)ctx->devs.size();
    const int spp = params->nbRayonParPixel;
    int launches = 0;

    // spp split: device k renders samples [k*spp/N, (k+1)*spp/N) of every pixel
    std::vector<const long long*> srcs;
    for (int k = 0; k < n_dev; k++) {
        DeviceCtx& d = ctx->devs[k];
        // spp split (default): device k renders samples [k*spp/N, (k+1)*spp/N) of every pixel; tile split
        // (TRT_FLAG_TILE_SPLIT, or a tile_part of the caller's own): all samples of the tiles k mod N
        const bool tile_split = n_dev > 1 && (params->flags & TRT_FLAG_TILE_SPLIT) != 0 && params->tile_part == 0;
        int s0 = (int)((long long)spp * k / n_dev), s1 = (int)((long long)spp * (k + 1) / n_dev);
        int t_off = params->tile_part ? (params->tile_part & 0xffff) : 0, t_stride = params->tile_part ? ((params->tile_part >> 16) & 0xffff) : 1;
        if (tile_split) { s0 = 0; s1 = spp; t_off = k; t_stride = n_dev; }
        CU(ctx, cudaSetDevice(d.dev));
        if ((rc = ensure_accum(ctx, d, npix)) != TRT_OK) return rc;
        CU(ctx, cudaMemsetAsync(d.accum, 0, sizeof(long long) * 9 * npix, d.stream));
        if (s1 > s0) {
            if ((rc = launch_trace(ctx, d, params, s1 - s0, params->sample_offset + s0, d.accum, t_off, t_stride)) != TRT_OK) return rc;
            launches += d.last_launches;
        } else {
            CU(ctx, cudaEventRecord(d.ev_a, d.stream));
            CU(ctx, cudaEventRecord(d.ev_b, d.stream));
            memset(d.h_ctrs, 0, sizeof(unsigned long long) * 16);
        }
        CU(ctx, cudaEventRecord(d.ev_c, d.stream));
        srcs.push_back(d.accum);
    }
    // Every device resolves its own BAND of the image (pixels [npix*k/N, npix*(k+1)/N)): it reads that band of every device's int64
    // sums over NVLink peer mappings while it tonemaps (the reduction is fused into the resolve and spread over all devices and
    // all their NVLink ports), then sends its band to the host over its own PCIe link; the host drains the bands one after the other
    // while the later ones are still arriving.
    DeviceCtx& root = ctx->devs[0];
    trt_params rp = *params;
    if (rp.sample_total <= 0) rp.sample_total = spp;
    const bool all_pinned = host_is_pinned(canva) && host_is_pinned(albedo_tab) && host_is_pinned(normal_tab);      // then no staging buffer is needed
    std::vector<size_t> band(n_dev + 1);
    for (int k = 0; k <= n_dev; k++) band[k] = npix * (size_t)k / (size_t)n_dev;
    for (int k = 0; k < n_dev; k++) {                      // while the GPUs trace: output and staging buffers for the bands
        DeviceCtx& d = ctx->devs[k];
        const size_t nb = band[k + 1] - band[k];
        CU(ctx, cudaSetDevice(d.dev));
        if ((rc = ensure_out(ctx, d, nb)) != TRT_OK) return rc;
        if (!all_pinned) ensure_stage(d, sizeof(double) * 9 * nb);
    }
    std::vector<int> mode(n_dev, 0);
    for (int k = 0; k < n_dev; k++) {
        DeviceCtx& d = ctx->devs[k];
        CU(ctx, cudaSetDevice(d.dev));
        for (int j = 0; j < n_dev; j++) if (j != k) CU(ctx, cudaStreamWaitEvent(d.stream, ctx->devs[j].ev_c, 0));
        if ((rc = resolve_and_send(ctx, &rp, srcs.data(), n_dev, d, band[k], band[k + 1], canva ? canva + band[k] : nullptr,
                                   albedo_tab ? albedo_tab + band[k] : nullptr, normal_tab ? normal_tab + band[k] : nullptr, &mode[k])) != TRT_OK) return rc;
        launches++;
    }
    std::vector<OutJob> jobs(n_dev);
    for (int k = 0; k < n_dev; k++)
        jobs[k] = OutJob{&ctx->devs[k], band[k + 1] - band[k], mode[k], canva ? canva + band[k] : nullptr,
                         albedo_tab ? albedo_tab + band[k] : nullptr, normal_tab ? normal_tab + band[k] : nullptr};
    if ((rc = drain_jobs(ctx, jobs.data(), n_dev)) != TRT_OK) return rc;
    for (auto& d : ctx->devs) { CU(ctx, cudaSetDevice(d.dev)); CU(ctx, cudaStreamSynchronize(d.stream)); }
    if ((rc = check_watchdog(ctx)) != TRT_OK) return rc;
    if (stats) {
        memset(stats, 0, sizeof *stats);
        collect_stats(ctx, params, stats, true);
        float ms = 0.0f;
        cudaSetDevice(root.dev);
        if (cudaEventElapsedTime(&ms, root.ev_c, root.ev_d) == cudaSuccess) stats->ms_reduce_resolve = ms;
        stats->kernel_launches = launches;
        stats->ms_total = now_ms() - t0;
        stats->ms_d2h = stats->ms_total - stats->ms_trace - stats->ms_reduce_resolve;
        if (stats->ms_d2h < 0) stats->ms_d2h = 0;
    }
    return TRT_OK;
}

// One process per GPU (nccl_glue.cu): this rank's share of the samples (or tiles), NCCL reduce onto rank 0, rank 0 resolves.
static int trt_render_distributed_impl(trt_ctx* ctx, const trt_params* params, void* nccl_comm, int rank, int n_ranks,
                                   trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab, trt_stats* stats)
{
    int rc = check_params(ctx, params);
    if (rc != TRT_OK) return rc;
    if (ctx->devs.size() != 1) return fail(ctx, TRT_ERR_INVALID, "trt_render_distributed works on a single-device ctx (one process per GPU)");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ctx, TRT_ERR_INVALID, "rank %d of %d", rank, n_ranks);
    if (n_ranks > 1 && !nccl_comm) return fail(ctx, TRT_ERR_INVALID, "nccl_comm is NULL");
    if (params->tile_part != 0) return fail(ctx, TRT_ERR_INVALID, "tile_part is chosen by trt_render_distributed (TRT_FLAG_TILE_SPLIT)");
    const double t0 = now_ms();
    DeviceCtx& d = ctx->devs[0];
    const size_t npix = (size_t)params->largeur_image * params->hauteur_image;
    const int spp = params->nbRayonParPixel;
    CU(ctx, cudaSetDevice(d.dev));
    if ((rc = ensure_accum(ctx, d, npix)) != TRT_OK) return rc;
    CU(ctx, cudaMemsetAsync(d.accum, 0, sizeof(long long) * 9 * npix, d.stream));
    const bool tile_split = n_ranks > 1 && (params->flags & TRT_FLAG_TILE_SPLIT) != 0;
    int s0 = (int)((long long)spp * rank / n_ranks), s1 = (int)((long long)spp * (rank + 1) / n_ranks), t_off = 0, t_stride = 1;
    if (tile_split) { s0 = 0; s1 = spp; t_off = rank; t_stride = n_ranks; }
    if (s1 > s0) {
        if ((rc = launch_trace(ctx, d, params, s1 - s0, params->sample_offset + s0, d.accum, t_off, t_stride)) != TRT_OK) return rc;
    } else {
        CU(ctx, cudaEventRecord(d.ev_a, d.stream));
        CU(ctx, cudaEventRecord(d.ev_b, d.stream));
        memset(d.h_ctrs, 0, sizeof(unsigned long long) * 16);
        d.last_launches = 0;
    }
    CU(ctx, cudaEventRecord(d.ev_c, d.stream));
    if (rank == 0 && !(host_is_pinned(canva) && host_is_pinned(albedo_tab) && host_is_pinned(normal_tab))) ensure_stage(d, sizeof(double) * 9 * npix);      // while the GPU traces
    if (n_ranks > 1 && (rc = trt_reduce_nccl(ctx, nccl_comm, (int64_t*)d.accum, 9 * npix, 0)) != TRT_OK) return rc;
    int launches = d.last_launches;
    if (rank == 0) {
        if ((rc = ensure_out(ctx, d, npix)) != TRT_OK) return rc;
        trt_params rp = *params;
        if (rp.sample_total <= 0) rp.sample_total = spp;
        const long long* src = d.accum;
        int mode = 0;
        if ((rc = resolve_and_send(ctx, &rp, &src, 1, d, 0, npix, canva, albedo_tab, normal_tab, &mode)) != TRT_OK) return rc;
        launches++;
        if ((rc = out_finish(ctx, d, npix, mode, canva, albedo_tab, normal_tab)) != TRT_OK) return rc;
    } else {
        CU(ctx, cudaEventRecord(d.ev_d, d.stream));
    }
    CU(ctx, cudaStreamSynchronize(d.stream));
    if ((rc = check_watchdog(ctx)) != TRT_OK) return rc;
    if (stats) {
        memset(stats, 0, sizeof *stats);
        trt_params mine = *params;
        mine.nbRayonParPixel = s1 - s0;
        collect_stats(ctx, &mine, stats, true);                          // this rank's paths and rays
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, d.ev_c, d.ev_d) == cudaSuccess) stats->ms_reduce_resolve = ms;
        stats->kernel_launches = launches;
        stats->ms_total = now_ms() - t0;
        stats->ms_d2h = stats->ms_total - stats->ms_trace - stats->ms_reduce_resolve;
        if (stats->ms_d2h < 0) stats->ms_d2h = 0;
    }
    return TRT_OK;
}

TRT_API int trt_primary(trt_ctx* ctx, const trt_params* params, double* t, int* id, trt_vec3* normal, trt_vec3* albedo)
{
    int rc = check_params(ctx, params);
    if (rc != TRT_OK) return rc;
    if (params->hauteur_image > 65535) return fail(ctx, TRT_ERR_INVALID, "trt_primary: at most 65535 rows (one grid row per image row)");
    DeviceCtx& d = ctx->devs[0];
    const size_t npix = (size_t)params->largeur_image * params->hauteur_image;
    CU(ctx, cudaSetDevice(d.dev));
    double* buf = nullptr;                    // t | normal | albedo | id
    CU(ctx, cudaMalloc(&buf, sizeof(double) * 8 * npix));
    PrimaryArgs P{};
    P.scene = d.scene; P.W = params->largeur_image; P.H = params->hauteur_image;
    P.focus = params->focus_distance;
    memcpy(P.cam, &params->cam, sizeof(double) * 12);
    P.t = buf; P.normal = buf + npix; P.albedo = buf + 4 * npix; P.id = (int*)(buf + 7 * npix);
    cudaError_t e = trt_launch_primary(P, d.stream);
    if (e == cudaSuccess && t) e = cudaMemcpyAsync(t, P.t, sizeof(double) * npix, cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess && normal) e = cudaMemcpyAsync(normal, P.normal, sizeof(double) * 3 * npix, cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess && albedo) e = cudaMemcpyAsync(albedo, P.albedo, sizeof(double) * 3 * npix, cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess && id) e = cudaMemcpyAsync(id, P.id, sizeof(int) * npix, cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    cudaFree(buf);
    CU(ctx, e);
    return TRT_OK;
}

TRT_API int trt_primary_fp32(trt_ctx* ctx, const trt_params* params, float* t, int* id)
{
    int rc = check_params(ctx, params);
    if (rc != TRT_OK) return rc;
    DeviceCtx& d = ctx->devs[0];
    const size_t npix = (size_t)params->largeur_image * params->hauteur_image;
    CU(ctx, cudaSetDevice(d.dev));
    RenderArgs A{};
    fill_camera_args(A, d, params);
    int* d_id = nullptr;
    CU(ctx, cudaMalloc(&d_id, (sizeof(int) + sizeof(float)) * npix));
    float* d_t = (float*)(d_id + npix);
    cudaError_t e = trt_launch_primary_fp32(A, d_id, d_t, d.stream);
    if (e == cudaSuccess && id) e = cudaMemcpyAsync(id, d_id, sizeof(int) * npix, cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess && t) e = cudaMemcpyAsync(t, d_t, sizeof(float) * npix, cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    cudaFree(d_id);
    CU(ctx, e);
    return TRT_OK;
}

// ---- guarded entry points (bodies above, *_impl) ----------------------------------------------------------
TRT_API int trt_scene_set(trt_ctx* ctx,
                          const trt_sphere* sphere_list, int nbSpheres,
                          const trt_triangle* triangle_list, int nbTriangles,
                          const trt_material* mat_list, int tex_width, int tex_height, int nbMaterials,
                          const int* quelMatPourTri,
                          const trt_mat_override* overrides, int n_overrides)
{
    TRT_GUARDED(ctx, trt_scene_set_impl(ctx, sphere_list, nbSpheres, triangle_list, nbTriangles, mat_list, tex_width, tex_height, nbMaterials, quelMatPourTri, overrides, n_overrides))
}

TRT_API int trt_sky_set(trt_ctx* ctx, const trt_material* sky_mat_list, int sky_width, int sky_height)
{
    TRT_GUARDED(ctx, trt_sky_set_impl(ctx, sky_mat_list, sky_width, sky_height))
}

TRT_API int trt_render(trt_ctx* ctx, const trt_params* params,
                       trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab, trt_stats* stats)
{
    TRT_GUARDED(ctx, trt_render_impl(ctx, params, canva, albedo_tab, normal_tab, stats))
}

TRT_API int trt_resolve(trt_ctx* ctx, const trt_params* params, const int64_t* d_accum,
                        trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab)
{
    TRT_GUARDED(ctx, trt_resolve_impl(ctx, params, d_accum, canva, albedo_tab, normal_tab))
}

TRT_API int trt_render_distributed(trt_ctx* ctx, const trt_params* params, void* nccl_comm, int rank, int n_ranks,
                                   trt_vec3* canva, trt_vec3* albedo_tab, trt_vec3* normal_tab, trt_stats* stats)
{
    TRT_GUARDED(ctx, trt_render_distributed_impl(ctx, params, nccl_comm, rank, n_ranks, canva, albedo_tab, normal_tab, stats))
}

// ---- known-answer entry points ---------------------------------------------------------------------
namespace {
struct DevBuf {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    ~DevBuf() { if (p) cudaFree(p); }
};
}  // namespace

#define KAT_BEGIN                                                   \
    if (!ctx) return TRT_ERR_INVALID;                               \
    if (n < 0) return fail(ctx, TRT_ERR_INVALID, "n < 0");          \
    if (n == 0) return TRT_OK;                                      \
    DeviceCtx& d = ctx->devs[0];                                    \
    CU(ctx, cudaSetDevice(d.dev));
#define UP(buf, src, bytes) CU(ctx, buf.alloc(bytes)); CU(ctx, cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, d.stream))
#define OUTBUF(buf, bytes) CU(ctx, buf.alloc(bytes))
#define DOWN(dst, buf, bytes) CU(ctx, cudaMemcpyAsync(dst, buf.p, bytes, cudaMemcpyDeviceToHost, d.stream))

TRT_API int trt_kat_hit_sphere(trt_ctx* ctx, int n, const float* cr, const float* ray, int* hit, float* dst, float* point, float* normal)
{
    KAT_BEGIN
    DevBuf a, b, h, t, p, q;
    UP(a, cr, sizeof(float) * 4 * n); UP(b, ray, sizeof(float) * 6 * n);
    OUTBUF(h, sizeof(int) * n); OUTBUF(t, sizeof(float) * n); OUTBUF(p, sizeof(float) * 3 * n); OUTBUF(q, sizeof(float) * 3 * n);
    CU(ctx, trt_kat_launch_sphere(n, (const float4*)a.p, (const float*)b.p, (int*)h.p, (float*)t.p, (float*)p.p, (float*)q.p, d.stream));
    DOWN(hit, h, sizeof(int) * n); DOWN(dst, t, sizeof(float) * n); DOWN(point, p, sizeof(float) * 3 * n); DOWN(normal, q, sizeof(float) * 3 * n);
    CU(ctx, cudaStreamSynchronize(d.stream));
    return TRT_OK;
}

TRT_API int trt_kat_hit_triangle(trt_ctx* ctx, int n, const float* abc, const float* ray, int* hit, float* dst_u_v)
{
    KAT_BEGIN
    DevBuf a, b, h, t;
    UP(a, abc, sizeof(float) * 9 * n); UP(b, ray, sizeof(float) * 6 * n);
    OUTBUF(h, sizeof(int) * n); OUTBUF(t, sizeof(float) * 3 * n);
    CU(ctx, trt_kat_launch_triangle(n, (const float*)a.p, (const float*)b.p, (int*)h.p, (float*)t.p, d.stream));
    DOWN(hit, h, sizeof(int) * n); DOWN(dst_u_v, t, sizeof(float) * 3 * n);
    CU(ctx, cudaStreamSynchronize(d.stream));
    return TRT_OK;
}

TRT_API int trt_kat_refract(trt_ctx* ctx, int n, const float* in8, float* out)
{
    KAT_BEGIN
    DevBuf a, o;
    UP(a, in8, sizeof(float) * 8 * n); OUTBUF(o, sizeof(float) * 3 * n);
    CU(ctx, trt_kat_launch_refract(n, (const float*)a.p, (float*)o.p, d.stream));
    DOWN(out, o, sizeof(float) * 3 * n);
    CU(ctx, cudaStreamSynchronize(d.stream));
    return TRT_OK;
}

TRT_API int trt_kat_philox(trt_ctx* ctx, int n, const uint32_t* ctr_key, uint32_t* out)
{
    KAT_BEGIN
    DevBuf a, o;
    UP(a, ctr_key, sizeof(uint32_t) * 6 * n); OUTBUF(o, sizeof(uint32_t) * 4 * n);
    CU(ctx, trt_kat_launch_philox(n, (const uint32_t*)a.p, (uint32_t*)o.p, d.stream));
    DOWN(out, o, sizeof(uint32_t) * 4 * n);
    CU(ctx, cudaStreamSynchronize(d.stream));
    return TRT_OK;
}

TRT_API int trt_kat_tonemap(trt_ctx* ctx, int n, const int64_t* sums, int spp, double* out)
{
    KAT_BEGIN
    if (spp < 1) return fail(ctx, TRT_ERR_INVALID, "spp < 1");
    DevBuf a, o;
    UP(a, sums, sizeof(int64_t) * n); OUTBUF(o, sizeof(double) * n);
    CU(ctx, trt_kat_launch_tonemap(n, (const long long*)a.p, spp, (double*)o.p, d.stream));
    DOWN(out, o, sizeof(double) * n);
    CU(ctx, cudaStreamSynchronize(d.stream));
    return TRT_OK;
}

}  // extern "C"
