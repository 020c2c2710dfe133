// nccl_glue.cu — the one collective of the path, inside the C library: the int64 sample sums of a render split over
// several PROCESSES (one per GPU) are added onto one rank with ncclReduce over NVLink / NVSwitch (BASELINE
// config 5).  The reference has no multi-process path (pthreads in one address space, main.c:402-453); this is what lets
// the drop-in host program run one process per GPU without Python or torch.distributed.
//
// NCCL is resolved at run time: dlopen("libnccl.so.2", RTLD_NOLOAD) first, so that a host process that already carries
// an NCCL (e.g. PyTorch's bundled one) keeps using that instance, then a plain dlopen.  libtipe_rt.so has no link-time
// dependency on NCCL; without the library these entry points fail with a message, everything else works.
#include <dlfcn.h>
#include <nccl.h>       // types and enums only

#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/tipe_rt.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    char error[256] = "";
};

NcclApi g_nccl;
std::once_flag g_once;

void load_nccl()
{
    NcclApi& n = g_nccl;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { n.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD); if (n.handle) break; }
    if (!n.handle) for (const char* nm : names) { n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (n.handle) break; }
    if (!n.handle) { snprintf(n.error, sizeof n.error, "NCCL not found (dlopen libnccl.so.2: %s)", dlerror()); return; }
#define SYM(field, name)                                                                          \
    *(void**)(&n.field) = dlsym(n.handle, name);                                                  \
    if (!n.field) { snprintf(n.error, sizeof n.error, "NCCL symbol %s missing", name); n.handle = nullptr; return; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Reduce, "ncclReduce") SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString") SYM(GetVersion, "ncclGetVersion")
#undef SYM
}

NcclApi* nccl()
{
    std::call_once(g_once, load_nccl);
    return g_nccl.handle ? &g_nccl : nullptr;
}

}  // namespace

int trt_set_error(trt_ctx* ctx, int code, const char* text);      // api.cu

extern "C" {

TRT_API int trt_nccl_version(void)
{
    NcclApi* n = nccl();
    int v = 0;
    if (!n || n->GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

TRT_API int trt_nccl_unique_id(void* out128)
{
    NcclApi* n = nccl();
    if (!n) return trt_set_error(nullptr, TRT_ERR_NCCL, g_nccl.error);
    if (!out128) return trt_set_error(nullptr, TRT_ERR_INVALID, "out128 is NULL");
    ncclUniqueId id;
    const ncclResult_t r = n->GetUniqueId(&id);
    if (r != ncclSuccess) return trt_set_error(nullptr, TRT_ERR_NCCL, n->GetErrorString(r));
    memcpy(out128, &id, TRT_NCCL_UNIQUE_ID_BYTES);
    return TRT_OK;
}

TRT_API int trt_nccl_comm_init(trt_ctx* ctx, int n_ranks, int rank, const void* unique_id128, void** comm_out)
{
    NcclApi* n = nccl();
    if (!ctx) return TRT_ERR_INVALID;
    if (!n) return trt_set_error(ctx, TRT_ERR_NCCL, g_nccl.error);
    if (!unique_id128 || !comm_out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return trt_set_error(ctx, TRT_ERR_INVALID, "bad rank / communicator arguments");
    const int dev = trt_device_of(ctx, 0);
    if (dev < 0 || cudaSetDevice(dev) != cudaSuccess) return trt_set_error(ctx, TRT_ERR_CUDA, "cudaSetDevice failed");
    ncclUniqueId id;
    memcpy(&id, unique_id128, TRT_NCCL_UNIQUE_ID_BYTES);
    ncclComm_t comm = nullptr;
    const ncclResult_t r = n->CommInitRank(&comm, n_ranks, id, rank);
    if (r != ncclSuccess) return trt_set_error(ctx, TRT_ERR_NCCL, n->GetErrorString(r));
    *comm_out = (void*)comm;
    return TRT_OK;
}

TRT_API int trt_nccl_comm_destroy(void* comm)
{
    NcclApi* n = nccl();
    if (!n || !comm) return TRT_OK;
    return n->CommDestroy((ncclComm_t)comm) == ncclSuccess ? TRT_OK : TRT_ERR_NCCL;
}

TRT_API int trt_reduce_nccl(trt_ctx* ctx, void* comm, int64_t* d_accum, size_t count, int root)
{
    NcclApi* n = nccl();
    if (!ctx) return TRT_ERR_INVALID;
    if (!n) return trt_set_error(ctx, TRT_ERR_NCCL, g_nccl.error);
    if (!comm || !d_accum) return trt_set_error(ctx, TRT_ERR_INVALID, "comm / d_accum is NULL");
    const int dev = trt_device_of(ctx, 0);
    if (dev < 0 || cudaSetDevice(dev) != cudaSuccess) return trt_set_error(ctx, TRT_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t stream = (cudaStream_t)trt_stream(ctx);
    // integer sum: the result does not depend on the order NCCL adds in, so the image is the same for any number of ranks
    const ncclResult_t r = root >= 0 ? n->Reduce(d_accum, d_accum, count, ncclInt64, ncclSum, root, (ncclComm_t)comm, stream)
                                     : n->AllReduce(d_accum, d_accum, count, ncclInt64, ncclSum, (ncclComm_t)comm, stream);
    if (r != ncclSuccess) return trt_set_error(ctx, TRT_ERR_NCCL, n->GetErrorString(r));
    return TRT_OK;
}

}  // extern "C"
