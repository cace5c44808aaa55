// capi.cu -- the extern "C" boundary declared in include/icpsymm.h.
//
// Mirrors the roles of the reference's public interface: MyICP (ICP/myicp.h:7-36) and
// registrateNPoint (ICP/regist.h:8).  No CPU fallback: every compute entry needs a CUDA device.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "internal.h"

int icp_fail(icpsymm_ctx* ctx, cudaError_t e, const char* expr, const char* file, int line) {
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, expr);
    cudaGetLastError();
    return ICPSYMM_ERR_CUDA;
}

int icp_fail_msg(icpsymm_ctx* ctx, int status, const char* msg) {
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s", msg);
    return status;
}

extern "C" {

const char* icpsymm_status_string(int s) {
    switch (s) {
        case ICPSYMM_OK: return "ok";
        case ICPSYMM_WARN_DEGENERATE: return "degenerate geometry: a pivot of the 6x6 system was dropped";
        case ICPSYMM_ERR_INVALID: return "invalid argument";
        case ICPSYMM_ERR_CUDA: return "CUDA error (no CPU fallback exists)";
        case ICPSYMM_ERR_NOT_READY: return "clouds or normals missing";
        case ICPSYMM_ERR_SIZE: return "size mismatch";
        case ICPSYMM_ERR_NONFINITE: return "non-finite coordinate";
        case ICPSYMM_ERR_COMM: return "communication error";
        default: return "unknown status";
    }
}

const char* icpsymm_last_error(const icpsymm_ctx* ctx) { return ctx ? ctx->err : "null context"; }

int icpsymm_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) { cudaGetLastError(); return ICPSYMM_ERR_CUDA; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return ICPSYMM_OK;
}

int icpsymm_create_on_stream(icpsymm_ctx** out, int device, void* cuda_stream) {
    if (!out) return ICPSYMM_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return ICPSYMM_ERR_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ICPSYMM_ERR_CUDA; }
    icpsymm_ctx* ctx = new icpsymm_ctx();
    ctx->device = device;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) == cudaSuccess) ctx->sm_count = p.multiProcessorCount;
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    *out = ctx;
    return ICPSYMM_OK;
}

int icpsymm_create(icpsymm_ctx** out, int device) {
    int rc = icpsymm_create_on_stream(out, device, nullptr);
    if (rc != ICPSYMM_OK) return rc;
    icpsymm_ctx* ctx = *out;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        *out = nullptr;
        return ICPSYMM_ERR_CUDA;
    }
    ctx->own_stream = true;
    return ICPSYMM_OK;
}

void icpsymm_destroy(icpsymm_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    icp_join_clouds(ctx, false);
    cudaStreamSynchronize(ctx->stream);
    icpsymm_comm_destroy(ctx);
    icpsymm_peer_detach(ctx);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    icp_free_cloudset(ctx->src);
    icp_free_cloudset(ctx->tgt);
    for (int k = 0; k < 2; k++) {
        if (ctx->aux[k]) { cudaStreamSynchronize(ctx->aux[k]); cudaStreamDestroy(ctx->aux[k]); }
        if (ctx->ready[k]) cudaEventDestroy(ctx->ready[k]);
        if (ctx->built[k]) cudaEventDestroy(ctx->built[k]);
        BuildScratch& B = ctx->scratch[k];
        cudaFree(B.staging); cudaFree(B.keys_a); cudaFree(B.keys_b); cudaFree(B.vals_a); cudaFree(B.vals_b);
        cudaFree(B.cub_temp); cudaFree(B.d_flags);
        if (B.h_flag) cudaFreeHost(B.h_flag);
    }
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    cudaFree(ctx->d_state); cudaFree(ctx->d_lp); cudaFree(ctx->d_ctrl);
    cudaFree(ctx->corr_pos); cudaFree(ctx->partial); cudaFree(ctx->gsum); cudaFree(ctx->gparts); cudaFree(ctx->trace_T);
    cudaFree(ctx->trace_diff); cudaFree(ctx->io_buf); cudaFree(ctx->order_buf); cudaFree(ctx->order_temp);
    cudaFree(ctx->margin2); cudaFree(ctx->ref_search); cudaFree(ctx->hist); cudaFree(ctx->tile_seed);
    if (ctx->h_state) cudaFreeHost(ctx->h_state);
    if (ctx->done_ev) cudaEventDestroy(ctx->done_ev);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    for (int i = 0; i < ctx->ev_pool_n; i++) cudaEventDestroy(ctx->ev_pool[i]);
    free(ctx->ev_pool);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
}

void* icpsymm_get_stream(const icpsymm_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int icpsymm_synchronize(icpsymm_ctx* ctx) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    { int jr = icp_join_clouds(ctx, true); if (jr < 0) return jr; }
    ICP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return ICPSYMM_OK;
}

int icpsymm_set_option(icpsymm_ctx* ctx, int key, double value) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    switch (key) {
        case ICPSYMM_OPT_REL_DIFF_TOL: if (value < 0) return ICPSYMM_ERR_INVALID; ctx->opt_rel_diff_tol = value; break;
        case ICPSYMM_OPT_IDENTITY_PAIRS: ctx->opt_identity = value != 0; break;
        case ICPSYMM_OPT_MAX_PAIR_DIST: if (value < 0) return ICPSYMM_ERR_INVALID; ctx->opt_max_pair_dist = (float)value; break;
        case ICPSYMM_OPT_MIN_NORMAL_DOT: ctx->opt_min_normal_dot = (value <= -1.0) ? -2.f : (float)value; break;
        case ICPSYMM_OPT_USE_GRAPH: ctx->opt_use_graph = (value == 2) ? 2 : (value != 0); break;
        case ICPSYMM_OPT_KERNEL_TIMING: ctx->opt_kernel_timing = value != 0; break;
        case ICPSYMM_OPT_WARM_START: ctx->opt_warm_start = (value == 2) ? 2 : (value != 0); break;
        case ICPSYMM_OPT_TRACE: ctx->opt_trace = value != 0; break;
        case ICPSYMM_OPT_GRID_BLOCKS: if (value < 0) return ICPSYMM_ERR_INVALID; ctx->opt_grid_blocks = (int)value; break;
        case ICPSYMM_OPT_TILE_ORDER: ctx->opt_tile_order = value != 0; break;
        case ICPSYMM_OPT_SKIP_WALK: ctx->opt_skip_walk = value != 0; break;
        case ICPSYMM_OPT_PROBE: ctx->opt_probe = value != 0; break;
        case ICPSYMM_OPT_ASYNC_SETUP: {
            int jr = icp_join_clouds(ctx, true);   // never switch modes with work in flight
            ctx->opt_async_setup = value != 0;
            if (jr < 0) return jr;
            break;
        }
        case ICPSYMM_OPT_PEER_TIMEOUT_MS: if (!(value > 0)) return ICPSYMM_ERR_INVALID; ctx->opt_peer_timeout_ms = value; break;
        default: return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "unknown option key");
    }
    return ICPSYMM_OK;
}

static int set_cloud(icpsymm_ctx* ctx, CloudSet& cs, const float* xyz, bool dev, const uint32_t* counts, size_t nseg,
                     size_t max_count, size_t stride) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return icp_fail_msg(ctx, ICPSYMM_ERR_CUDA, "cudaSetDevice");
    ctx->have_corr = false;
    cs.slot = (&cs == &ctx->tgt) ? 1 : 0;
    return icp_build_cloudset(ctx, cs, xyz, dev, counts, nseg, max_count, stride);
}

int icpsymm_set_source(icpsymm_ctx* ctx, const float* xyz, size_t n, size_t stride) {
    return ctx ? set_cloud(ctx, ctx->src, xyz, false, nullptr, 1, n, stride) : ICPSYMM_ERR_INVALID;
}
int icpsymm_set_target(icpsymm_ctx* ctx, const float* xyz, size_t n, size_t stride) {
    return ctx ? set_cloud(ctx, ctx->tgt, xyz, false, nullptr, 1, n, stride) : ICPSYMM_ERR_INVALID;
}
int icpsymm_set_source_dev(icpsymm_ctx* ctx, const float* xyz, size_t n, size_t stride) {
    return ctx ? set_cloud(ctx, ctx->src, xyz, true, nullptr, 1, n, stride) : ICPSYMM_ERR_INVALID;
}
int icpsymm_set_target_dev(icpsymm_ctx* ctx, const float* xyz, size_t n, size_t stride) {
    return ctx ? set_cloud(ctx, ctx->tgt, xyz, true, nullptr, 1, n, stride) : ICPSYMM_ERR_INVALID;
}
int icpsymm_set_source_batch(icpsymm_ctx* ctx, const float* xyz, const uint32_t* counts, size_t n_pairs, size_t max_count,
                             size_t stride) {
    return ctx ? set_cloud(ctx, ctx->src, xyz, false, counts, n_pairs, max_count, stride) : ICPSYMM_ERR_INVALID;
}
int icpsymm_set_target_batch(icpsymm_ctx* ctx, const float* xyz, const uint32_t* counts, size_t n_pairs, size_t max_count,
                             size_t stride) {
    return ctx ? set_cloud(ctx, ctx->tgt, xyz, false, counts, n_pairs, max_count, stride) : ICPSYMM_ERR_INVALID;
}

int icpsymm_estimate_normals(icpsymm_ctx* ctx, int k, const float viewpoint[3]) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    if (!ctx->src.valid || !ctx->tgt.valid) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "set_source / set_target first");
    int rc = icp_estimate_normals(ctx, ctx->src, k, viewpoint);
    if (rc) return rc;
    rc = icp_estimate_normals(ctx, ctx->tgt, k, viewpoint);
    if (rc) return rc;
    if (!ctx->opt_async_setup) ICP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return ICPSYMM_OK;
}

int icpsymm_set_normals(icpsymm_ctx* ctx, int which, const float* normals4) {
    if (!ctx || (which != 0 && which != 1)) return ICPSYMM_ERR_INVALID;
    return icp_import_normals(ctx, which ? ctx->tgt : ctx->src, normals4);
}

int icpsymm_get_normals(icpsymm_ctx* ctx, int which, float* normals4) {
    if (!ctx || !normals4 || (which != 0 && which != 1)) return ICPSYMM_ERR_INVALID;
    return icp_export_normals(ctx, which ? ctx->tgt : ctx->src, normals4);
}

int icpsymm_register_symm(icpsymm_ctx* ctx, int max_iters, float diff_threshold, float* T_out, icpsymm_stats* stats) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    return icp_run_loop(ctx, max_iters, diff_threshold, nullptr, nullptr, false, T_out, stats);
}

int icpsymm_register_symm_async(icpsymm_ctx* ctx, int max_iters, float diff_threshold) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    return icp_loop_enqueue(ctx, max_iters, diff_threshold, nullptr, nullptr, false);
}

int icpsymm_wait(icpsymm_ctx* ctx, float* T_out, icpsymm_stats* stats) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    return icp_loop_finish(ctx, T_out, stats);
}

int icpsymm_find_correspondences(icpsymm_ctx* ctx, const float T[16], int32_t* idx_out, float* d2_out) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    int rc = icp_run_single(ctx, T, nullptr, nullptr);
    if (rc < 0) return rc;
    return icp_export_correspondences(ctx, idx_out, d2_out);
}

int icpsymm_get_correspondences(icpsymm_ctx* ctx, int32_t* idx_out, float* d2_out) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    return icp_export_correspondences(ctx, idx_out, d2_out);
}

int icpsymm_eval_diff(icpsymm_ctx* ctx, const float T[16], double* diff) {
    if (!ctx || !diff) return ICPSYMM_ERR_INVALID;
    icpsymm_stats st;
    int rc = icp_run_single(ctx, T, nullptr, &st);
    if (rc < 0) return rc;
    *diff = st.diff;
    return ICPSYMM_OK;
}

int icpsymm_normal_equations(icpsymm_ctx* ctx, const float T[16], const double origin[3], double sums[ICPSYMM_NSUMS]) {
    if (!ctx || !sums) return ICPSYMM_ERR_INVALID;
    if (!ctx->src.has_normals || !ctx->tgt.has_normals) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "normals missing");
    int rc = icp_run_single(ctx, T, origin, nullptr);
    if (rc < 0) return rc;
    return icp_read_sums(ctx, sums, nullptr);
}

int icpsymm_solve_compose(icpsymm_ctx* ctx, const double sums[ICPSYMM_NSUMS], const double origin[3], double Tinc[16],
                          double x[6], double* min_pivot_ratio) {
    if (!ctx || !sums) return ICPSYMM_ERR_INVALID;
    return icp_solve_compose_device(ctx, sums, origin, Tinc, x, min_pivot_ratio);
}

int icpsymm_get_transformed_source(icpsymm_ctx* ctx, const float T[16], float* xyz_out, size_t stride_bytes) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    return icp_export_transformed(ctx, T, xyz_out, stride_bytes);
}

int icpsymm_get_trace(icpsymm_ctx* ctx, size_t pair, int k, float T[16], double* diff) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    if (!ctx->opt_trace || !ctx->trace_T || !ctx->h_state) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "trace not enabled");
    if (pair >= ctx->src.nseg || k < 0 || k >= ctx->h_state[pair].searches || k >= ctx->trace_stride)
        return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "trace index out of range");
    float t12[12];
    double d;
    size_t o = pair * (size_t)ctx->trace_stride + (size_t)k;
    ICP_CUDA_CHECK(cudaMemcpyAsync(t12, ctx->trace_T + o * 12, sizeof t12, cudaMemcpyDeviceToHost, ctx->stream));
    ICP_CUDA_CHECK(cudaMemcpyAsync(&d, ctx->trace_diff + o, sizeof d, cudaMemcpyDeviceToHost, ctx->stream));
    ICP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (T) {
        for (int i = 0; i < 12; i++) T[i] = t12[i];
        T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
    }
    if (diff) *diff = d;
    return ICPSYMM_OK;
}

int icpsymm_register_p2p(icpsymm_ctx* ctx, int max_iters, float diff_threshold, float* T_out, icpsymm_stats* stats) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    return icp_register_p2p(ctx, max_iters, diff_threshold, T_out, stats);
}

int icpsymm_get_search_times(icpsymm_ctx* ctx, float* search_ms, float* solve_ms, int n) {
    if (!ctx || n < 0) return ICPSYMM_ERR_INVALID;
    return icp_search_times(ctx, search_ms, solve_ms, n);
}

int icpsymm_register_npoint(icpsymm_ctx* ctx, const double* src, const double* dst, size_t n, double R[9], double T[3]) {
    if (!ctx) return ICPSYMM_ERR_INVALID;
    return icp_kabsch(ctx, src, dst, n, R, T);
}

}  // extern "C"
