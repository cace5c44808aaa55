// internal.h -- host-side structures shared by the translation units of libicpsymm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/icpsymm.h"
#include "common.cuh"

// device storage of one cloud set (nseg independent clouds, fixed pitch)
struct CloudSet {
    uint32_t nseg = 0;
    uint32_t max_count = 0;       // largest segment
    size_t total = 0;             // sum of counts
    TreeShape shape{};
    float4* pts = nullptr;
    float4* nrm = nullptr;
    WideNode* nodes = nullptr;
    uint32_t* pos_of_orig = nullptr;
    uint32_t* unit_split = nullptr;   // per 32-point unit: lane mask of its second cluster (0 = compact unit)
    size_t cap_units = 0;
    uint32_t* counts = nullptr;   // device, nseg
    uint32_t* bbox = nullptr;     // device, nseg*8 ordered-uint (lo xyz, pad, hi xyz, pad) during build
    float4* bb_lo = nullptr;      // device, nseg
    float4* bb_hi = nullptr;
    uint32_t* h_counts = nullptr; // host copy
    size_t cap_pts = 0, cap_nodes = 0, cap_seg = 0;
    bool valid = false, has_normals = false;
    int slot = 0;                 // 0 = source, 1 = target: selects the build scratch and the aux stream
    bool on_aux = false;          // work of this cloud is outstanding on its aux stream (async setup)
    bool pending_check = false;   // the non-finite flag of the last build has not been read yet

    CloudView view() const {
        CloudView v;
        v.pts = pts; v.nrm = nrm; v.nodes = nodes; v.pos_of_orig = pos_of_orig; v.unit_split = unit_split; v.counts = counts;
        v.shape = shape; v.nseg = nseg;
        return v;
    }
};

struct GraphKey {
    int max_iters = -1;
    uint32_t nseg = 0, tiles_max = 0, src_pitch = 0, tgt_pitch = 0;
    // every device pointer the captured kernels hold belongs to this context; `generation` counts the
    // (re)allocations of all of them, so an unchanged generation means unchanged pointers
    unsigned long long generation = ~0ull;
    int mode = 0, shard_rank = 0, shard_world = 1, trace = 0, trace_stride = 0;   // mode: graph flavour + peer world
    bool operator==(const GraphKey& o) const {
        return max_iters == o.max_iters && nseg == o.nseg && tiles_max == o.tiles_max &&
               src_pitch == o.src_pitch && tgt_pitch == o.tgt_pitch && generation == o.generation && mode == o.mode &&
               shard_rank == o.shard_rank && shard_world == o.shard_world && trace == o.trace &&
               trace_stride == o.trace_stride;
    }
};

// scratch of one index build; one per cloud slot so that the two builds may overlap
struct BuildScratch {
    void* staging = nullptr;      size_t staging_bytes = 0;
    unsigned long long *keys_a = nullptr, *keys_b = nullptr;
    uint32_t *vals_a = nullptr, *vals_b = nullptr;
    size_t sort_cap = 0;
    void* cub_temp = nullptr;     size_t cub_temp_bytes = 0;
    int* d_flags = nullptr;       // [0] non-finite count
    int* h_flag = nullptr;        // pinned mirror
};

struct icpsymm_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    char err[512] = {0};

    CloudSet src, tgt;

    // build scratch (grow-only) and the aux streams of the asynchronous setup
    BuildScratch scratch[2];
    cudaStream_t aux[2] = {nullptr, nullptr};
    cudaEvent_t ready[2] = {nullptr, nullptr};   // end of the cloud's outstanding work on its aux stream
    cudaEvent_t built[2] = {nullptr, nullptr};   // end of its index build (the non-finite flag has been copied back)
    cudaEvent_t fork_ev = nullptr;

    // loop state
    SegState* d_state = nullptr;  size_t state_cap = 0;
    SegState* h_state = nullptr;  size_t h_state_cap = 0;   // pinned mirror
    LoopParams* d_lp = nullptr;
    void* d_ctrl = nullptr;       // LoopCtrl: ticket + done counter of the device-side loop
    uint32_t* corr_pos = nullptr; size_t corr_cap = 0;
    uint32_t* tile_seed = nullptr; size_t tile_seed_cap = 0;   // per tile: neighbour of its middle query (probe kernel)
    float* margin2 = nullptr;     size_t margin_cap = 0;   // per query: margin of its last walk (K1 skip test)
    uint8_t* ref_search = nullptr; size_t ref_cap = 0;     // per query: search of its last tracked walk (0xff: none)
    float* hist = nullptr;        size_t hist_cap = 0;     // per pair: fp32 transform of every search
    double* partial = nullptr;    size_t partial_cap = 0;
    double* gsum = nullptr;       size_t gsum_cap = 0;
    double* gparts = nullptr;     size_t gparts_cap = 0;   // part sums: exchange buffer of the NCCL mode
    float* trace_T = nullptr;     size_t trace_T_cap = 0;      // floats
    double* trace_diff = nullptr; size_t trace_diff_cap = 0;   // doubles
    int trace_stride = 0;
    void* io_buf = nullptr;       size_t io_bytes = 0;   // device scratch for exports/imports
    uint32_t* order_buf = nullptr; size_t order_cap = 0;  // tile order of the first search: keys/values, double-buffered
    void* order_temp = nullptr;   size_t order_temp_bytes = 0;
    bool cost_attr_set = false;
    bool have_corr = false;

    // options
    double opt_rel_diff_tol = 0.0;
    int opt_identity = 0;
    float opt_max_pair_dist = 0.f;
    float opt_min_normal_dot = -2.f;
    int opt_use_graph = 1;
    int opt_kernel_timing = 0;
    int opt_warm_start = 1;
    int opt_trace = 0;
    int opt_grid_blocks = 0;
    int opt_async_setup = 0;
    int opt_tile_order = 1;
    int opt_skip_walk = 1;
    int opt_probe = 1;
    double opt_peer_timeout_ms = 10000.0;

    // graph cache
    unsigned long long generation = 0;   // bumped by every device (re)allocation of this context (icp_reserve)
    cudaGraphExec_t graph_exec = nullptr;
    GraphKey graph_key;

    // a registration whose loop has been enqueued but whose result has not been handed out yet
    struct Pending {
        bool active = false, timing = false, use_comm = false, single_search = false;
        uint32_t nseg = 0, part0 = 0, nparts = ICP_NPARTS;
        int n_search = 0;
    } pending;
    cudaEvent_t done_ev = nullptr;

    // timing
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    cudaEvent_t* ev_pool = nullptr; int ev_pool_n = 0; int timed_searches = 0;

    // multi-GPU
    int shard_rank = 0, shard_world = 1;
    void* nccl_comm = nullptr;
    // NVLink peer exchange (comm.cu): my buffer + the mapped buffers of all ranks
    void* peer_local = nullptr;
    void* peer_ptrs[ICP_MAX_PEERS] = {nullptr};
    int peer_world = 0;
};

int icp_fail(icpsymm_ctx* ctx, cudaError_t e, const char* expr, const char* file, int line);
int icp_fail_msg(icpsymm_ctx* ctx, int status, const char* msg);

template <typename T>
int icp_reserve(icpsymm_ctx* ctx, T** p, size_t* cap, size_t need);   // grow-only cudaMalloc, element counts

// index_build.cu
// stream the work of cloud `cs` goes to: the context stream, or (ICPSYMM_OPT_ASYNC_SETUP) the
// cloud's aux stream, forked from the context stream; icp_cloud_done marks the end of that work
cudaStream_t icp_cloud_stream(icpsymm_ctx* ctx, CloudSet& cs);
void icp_cloud_done(icpsymm_ctx* ctx, CloudSet& cs);
// make the context stream wait for outstanding aux work; with `check` also resolve deferred
// non-finite checks (blocks the host on the build)
int icp_join_clouds(icpsymm_ctx* ctx, bool check);
TreeShape icp_make_shape(uint32_t max_count);
int icp_build_cloudset(icpsymm_ctx* ctx, CloudSet& cs, const float* xyz, bool on_device, const uint32_t* counts,
                       size_t nseg, size_t max_count, size_t stride_bytes);
void icp_free_cloudset(CloudSet& cs);

// normals.cu
int icp_estimate_normals(icpsymm_ctx* ctx, CloudSet& cs, int k, const float vp[3]);
int icp_export_normals(icpsymm_ctx* ctx, const CloudSet& cs, float* out_host);
int icp_import_normals(icpsymm_ctx* ctx, CloudSet& cs, const float* in_host);

// icp_loop.cu
int icp_run_loop(icpsymm_ctx* ctx, int max_iters, float thr, const float* T_init /*12 floats (3x4) or null*/,
                 const double* origin /*3 or null*/, bool single_search, float* T_out, icpsymm_stats* stats);
int icp_loop_enqueue(icpsymm_ctx* ctx, int max_iters, float thr, const float* T_init, const double* origin, bool single_search);
int icp_loop_finish(icpsymm_ctx* ctx, float* T_out, icpsymm_stats* stats);
int icp_run_single(icpsymm_ctx* ctx, const float* T16, const double* origin, icpsymm_stats* stats);
int icp_export_correspondences(icpsymm_ctx* ctx, int32_t* idx_out, float* d2_out);
int icp_export_transformed(icpsymm_ctx* ctx, const float* T, float* xyz_out, size_t stride_bytes);
int icp_solve_compose_device(icpsymm_ctx* ctx, const double* sums, const double* origin, double* Tinc, double* x,
                             double* ratio);
int icp_read_sums(icpsymm_ctx* ctx, double* sums, double* origin_out);
int icp_search_times(icpsymm_ctx* ctx, float* search_ms, float* solve_ms, int n);

// kabsch.cu
int icp_kabsch(icpsymm_ctx* ctx, const double* src, const double* dst, size_t n, double R[9], double T[3]);
int icp_register_p2p(icpsymm_ctx* ctx, int max_iters, float thr, float* T_out, icpsymm_stats* stats);

// comm.cu
int icp_comm_allreduce(icpsymm_ctx* ctx, double* buf, size_t count);
