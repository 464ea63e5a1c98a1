// NCCL plumbing for the sharded 1-D quantile path.  NCCL is resolved lazily with dlopen so that
// libndstats_b200.so loads (and single-GPU calls work) where no NCCL is installed, and so that inside a
// torch process the already-loaded bundled libnccl.so.2 is reused instead of a second copy.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "nds_internal.h"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi &api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) {
            a.error = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?");
            return;
        }
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
        a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
        a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
        if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.AllGather || !a.GetErrorString)
            a.error = "libnccl is missing a required symbol";
    });
    return a;
}

nds_status nccl_fail(nds_ctx *ctx, const char *where, ncclResult_t r) {
    char buf[256];
    std::snprintf(buf, sizeof buf, "%s: %s", where, api().GetErrorString ? api().GetErrorString(r) : "?");
    if (ctx) ctx->last_error = buf;
    return NDS_NCCL_ERROR;
}

}  // namespace

namespace nds {
int nccl_allreduce_sum_u64(nds_ctx *ctx, void *buf, size_t count, cudaStream_t stream, std::string *err) {
    NcclApi &a = api();
    if (!a.error.empty() || !ctx->nccl_comm) {
        if (err) *err = a.error.empty() ? "no communicator" : a.error;
        return -1;
    }
    ncclResult_t r = a.AllReduce(buf, buf, count, ncclUint64, ncclSum, (ncclComm_t)ctx->nccl_comm, stream);
    if (r != ncclSuccess) {
        if (err) *err = std::string("ncclAllReduce: ") + a.GetErrorString(r);
        return -1;
    }
    return 0;
}
int nccl_allgather_bytes(nds_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t stream,
                         std::string *err) {
    NcclApi &a = api();
    if (!a.error.empty() || !ctx->nccl_comm) {
        if (err) *err = a.error.empty() ? "no communicator" : a.error;
        return -1;
    }
    ncclResult_t r = a.AllGather(send, recv, bytes_per_rank, ncclUint8, (ncclComm_t)ctx->nccl_comm, stream);
    if (r != ncclSuccess) {
        if (err) *err = std::string("ncclAllGather: ") + a.GetErrorString(r);
        return -1;
    }
    return 0;
}
}  // namespace nds

extern "C" {

nds_status nds_comm_get_unique_id(void *id_out) {
    static_assert(sizeof(ncclUniqueId) == NDS_UNIQUE_ID_BYTES, "ncclUniqueId size");
    if (!id_out) return NDS_BAD_ARGUMENT;
    NcclApi &a = api();
    if (!a.error.empty()) return NDS_NCCL_ERROR;
    ncclUniqueId id;
    if (a.GetUniqueId(&id) != ncclSuccess) return NDS_NCCL_ERROR;
    std::memcpy(id_out, &id, sizeof id);
    return NDS_OK;
}

nds_status nds_comm_init_rank(nds_ctx *ctx, int nranks, int rank, const void *id) {
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return NDS_BAD_ARGUMENT;
    NcclApi &a = api();
    if (!a.error.empty()) {
        ctx->last_error = a.error;
        return NDS_NCCL_ERROR;
    }
    if (ctx->nccl_comm) nds_comm_destroy(ctx);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return NDS_CUDA_ERROR;
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof uid);
    ncclComm_t comm;
    ncclResult_t r = a.CommInitRank(&comm, nranks, uid, rank);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclCommInitRank", r);
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    nds::xchg_setup(ctx);  // collective: every rank maps every peer's inbox, or all of them stay on the NCCL collectives
    return NDS_OK;
}

nds_status nds_comm_stats(nds_ctx *ctx, uint64_t out[4]) {
    if (!ctx || !out) return NDS_BAD_ARGUMENT;
    out[0] = nds::xchg_ready(ctx) ? 1 : 0;
    out[1] = ctx->xchg_steps;
    out[2] = ctx->host_syncs;
    out[3] = (uint64_t)ctx->nranks;
    return NDS_OK;
}

nds_status nds_comm_destroy(nds_ctx *ctx) {
    if (!ctx) return NDS_BAD_ARGUMENT;
    if (ctx->xchg) nds::xchg_teardown(ctx, true);
    if (ctx->nccl_comm) {
        cudaStreamSynchronize(ctx->stream);
        api().CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->nranks = 1;
    ctx->rank = 0;
    return NDS_OK;
}

}  // extern "C"
