// comm.cu -- multi-GPU plumbing for the query-sharded mode.
//
// The reference is single-process; this is the one real exchange step of the B200 design: the
// 37-value normal-equation vector of every rank is summed before the solve, so that every rank
// applies a bit-identical transform.  NCCL is loaded lazily with dlopen so that the library has no
// link-time NCCL dependency (single-GPU users and the CPU-only symbol test never touch it).
#include <dlfcn.h>

#include <cstdio>
#include <cstring>

#include "internal.h"

namespace {

struct NcclUniqueId { char internal[128]; };
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclUniqueId, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_comm_destroy)(void*);
typedef const char* (*fn_error_string)(int);

struct NcclApi {
    void* lib = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_error_string error_string = nullptr;
};

NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return false;
    g_nccl.get_unique_id = (fn_get_unique_id)dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.all_reduce = (fn_all_reduce)dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.comm_destroy = (fn_comm_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.error_string = (fn_error_string)dlsym(g_nccl.lib, "ncclGetErrorString");
    return g_nccl.get_unique_id && g_nccl.comm_init_rank && g_nccl.all_reduce && g_nccl.comm_destroy;
}

constexpr int kNcclFloat64 = 8, kNcclSum = 0;

}  // namespace

int icp_comm_allreduce(icpsymm_ctx* ctx, double* buf, size_t count) {
    if (!ctx->nccl_comm) return ICPSYMM_ERR_COMM;
    int r = g_nccl.all_reduce(buf, buf, count, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
    return r == 0 ? ICPSYMM_OK : ICPSYMM_ERR_COMM;
}

extern "C" int icpsymm_comm_unique_id(void* id_out128) {
    if (!id_out128) return ICPSYMM_ERR_INVALID;
    if (!load_nccl()) return ICPSYMM_ERR_COMM;
    NcclUniqueId id;
    if (g_nccl.get_unique_id(&id) != 0) return ICPSYMM_ERR_COMM;
    memcpy(id_out128, &id, sizeof id);
    return ICPSYMM_OK;
}

extern "C" int icpsymm_comm_init(icpsymm_ctx* ctx, const void* id128, int n_ranks, int rank) {
    if (!ctx || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return ICPSYMM_ERR_INVALID;
    if (n_ranks > ICP_NPARTS) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "query sharding supports at most 8 ranks");
    if (!load_nccl()) return icp_fail_msg(ctx, ICPSYMM_ERR_COMM, "libnccl.so.2 not found");
    if (ctx->nccl_comm) { g_nccl.comm_destroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return icp_fail_msg(ctx, ICPSYMM_ERR_CUDA, "cudaSetDevice");
    NcclUniqueId id;
    memcpy(&id, id128, sizeof id);
    void* comm = nullptr;
    int r = g_nccl.comm_init_rank(&comm, n_ranks, id, rank);
    if (r != 0) {
        char msg[256];
        snprintf(msg, sizeof msg, "ncclCommInitRank: %s", g_nccl.error_string ? g_nccl.error_string(r) : "error");
        return icp_fail_msg(ctx, ICPSYMM_ERR_COMM, msg);
    }
    ctx->nccl_comm = comm;
    ctx->shard_rank = rank;
    ctx->shard_world = n_ranks;
    return ICPSYMM_OK;
}

extern "C" int icpsymm_comm_destroy(icpsymm_ctx* ctx) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    if (ctx->nccl_comm && g_nccl.comm_destroy) g_nccl.comm_destroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->shard_rank = 0;
    ctx->shard_world = 1;
    return ICPSYMM_OK;
}

extern "C" int icpsymm_set_query_shard(icpsymm_ctx* ctx, int rank, int n_ranks) {
    if (!ctx || n_ranks < 1 || n_ranks > ICP_NPARTS || rank < 0 || rank >= n_ranks) return ICPSYMM_ERR_INVALID;
    ctx->shard_rank = rank;
    ctx->shard_world = n_ranks;
    return ICPSYMM_OK;
}

// ---- NVLink peer exchange (used by solve_kernel, icp_loop.cu) ------------------------------------
static size_t peer_buffer_bytes() { return sizeof(PeerBuf); }

extern "C" int icpsymm_peer_export(icpsymm_ctx* ctx, void* handle64_out) {
    if (!ctx || !handle64_out) return ICPSYMM_ERR_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return icp_fail_msg(ctx, ICPSYMM_ERR_CUDA, "cudaSetDevice");
    if (!ctx->peer_local) {
        ICP_CUDA_CHECK(cudaMalloc(&ctx->peer_local, peer_buffer_bytes()));
        ICP_CUDA_CHECK(cudaMemset(ctx->peer_local, 0, peer_buffer_bytes()));
    }
    cudaIpcMemHandle_t h;
    ICP_CUDA_CHECK(cudaIpcGetMemHandle(&h, ctx->peer_local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64_out, &h, 64);
    return ICPSYMM_OK;
}

extern "C" int icpsymm_peer_attach(icpsymm_ctx* ctx, const void* handles, int n_ranks, int rank) {
    if (!ctx || !handles || n_ranks < 1 || n_ranks > ICP_MAX_PEERS || rank < 0 || rank >= n_ranks) return ICPSYMM_ERR_INVALID;
    if (!ctx->peer_local) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "call icpsymm_peer_export first");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return icp_fail_msg(ctx, ICPSYMM_ERR_CUDA, "cudaSetDevice");
    for (int r = 0; r < n_ranks; r++) {
        if (r == rank) { ctx->peer_ptrs[r] = ctx->peer_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        ICP_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_ptrs[r] = p;
    }
    ctx->peer_world = n_ranks;
    ctx->shard_rank = rank;
    ctx->shard_world = n_ranks;
    // a fresh exchange: every rank starts at sequence 0 with zeroed slots (attach is collective: all ranks
    // export, gather the handles and attach; re-attaching needs a detach on every rank first)
    if (ctx->d_ctrl) ICP_CUDA_CHECK(cudaMemsetAsync(ctx->d_ctrl, 0, sizeof(LoopCtrl), ctx->stream));
    ICP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    // the captured loop graph holds the old peer pointers
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    return ICPSYMM_OK;
}

extern "C" int icpsymm_peer_detach(icpsymm_ctx* ctx) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < ICP_MAX_PEERS; r++) {
        if (ctx->peer_ptrs[r] && ctx->peer_ptrs[r] != ctx->peer_local) cudaIpcCloseMemHandle(ctx->peer_ptrs[r]);
        ctx->peer_ptrs[r] = nullptr;
    }
    if (ctx->peer_local) { cudaFree(ctx->peer_local); ctx->peer_local = nullptr; }
    if (ctx->peer_world > 1) { ctx->shard_rank = 0; ctx->shard_world = 1; }
    ctx->peer_world = 0;
    if (ctx->d_ctrl) cudaMemsetAsync(ctx->d_ctrl, 0, sizeof(LoopCtrl), ctx->stream);   // sequence, error flag, tickets
    cudaStreamSynchronize(ctx->stream);
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    cudaGetLastError();
    return ICPSYMM_OK;
}
