/*
 * icpsymm.h -- C ABI of the B200-native symmetric-ICP registration path.
 *
 * This is the drop-in boundary for the hot path of StephenNG59/ICP-symm.  The reference has
 * no FFI layer: its boundary is two C++ headers, ICP/myicp.h:7-36 (class MyICP) and
 * ICP/regist.h:8 (registrateNPoint), whose types are PCL / OpenCV objects.  Every entry point
 * below names the reference interface it replaces.  Plain pointers and sizes only: no CUDA,
 * torch, PCL or OpenCV types.  The C++ facade that keeps the reference's class shape lives in
 * icp-symm_b200/host/myicp.h + regist.h and is written purely on top of this ABI.
 *
 * Conventions
 *   - xyz inputs: float, `stride_bytes` between consecutive points (12 for packed xyz, 16 for
 *     pcl::PointXYZ-style x,y,z,pad -- myicp.h:5).  Host pointers unless the name ends in _dev.
 *   - 4x4 transforms: float[16], ROW-major, map SOURCE -> TARGET (myicp.cpp:138 left-multiplies
 *     increments: transform = incre * transform).
 *   - normals: float[4] per point (nx, ny, nz, curvature), i.e. pcl::Normal.
 *   - indices returned to the caller are ORIGINAL point indices (order of the arrays passed in).
 *   - all calls on one context are serialised by the caller (one context per host thread).
 *   - return value: ICPSYMM_OK (0), a positive warning, or a negative error.  No CPU fallback
 *     exists: without a CUDA device every compute call returns ICPSYMM_ERR_CUDA.
 */
#ifndef ICPSYMM_H
#define ICPSYMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden: only this header is exported */
#endif

#define ICPSYMM_OK 0
#define ICPSYMM_WARN_DEGENERATE 1    /* a 6x6 pivot was dropped (rank-deficient geometry) */
#define ICPSYMM_ERR_INVALID (-1)     /* bad argument */
#define ICPSYMM_ERR_CUDA (-2)        /* CUDA runtime error (see icpsymm_last_error) */
#define ICPSYMM_ERR_NOT_READY (-3)   /* clouds or normals missing */
#define ICPSYMM_ERR_SIZE (-4)        /* size mismatch (identity pairing needs equal sizes, func.cpp:21) */
#define ICPSYMM_ERR_NONFINITE (-5)   /* NaN/Inf coordinate in an input cloud */
#define ICPSYMM_ERR_COMM (-6)        /* NCCL / peer-exchange error */

#define ICPSYMM_NSUMS 37             /* normal-equation vector length, layout in DESIGN.md */

typedef struct icpsymm_ctx icpsymm_ctx;

typedef struct icpsymm_stats {
    int32_t iters;            /* solves applied (trips of the while loop, myicp.cpp:123) */
    int32_t converged;        /* 1: diff <= threshold, 2: relative-diff tolerance, 0: hit max_iters */
    int32_t degenerate;       /* a pivot was dropped in some iteration */
    int32_t searches;         /* correspondence searches run (= iters + 1 unless stopped early) */
    double diff;              /* sum_i ||T p_i - q_nn(i)|| of the final search (evalDiff, func.cpp:19-40) */
    double rmse;              /* sqrt(mean d2) of the final search */
    double min_pivot_ratio;   /* min over iterations of min_j d_j/A_jj (weak-constraint indicator) */
    uint64_t pairs;           /* pairs used by the last solve */
    float loop_ms;            /* device time of the whole loop, CUDA events on the context stream */
    float search_ms;          /* summed duration of the NN+accumulate kernel (ICPSYMM_OPT_KERNEL_TIMING) */
    float solve_ms;           /* summed duration of the reduce+solve kernel  (ICPSYMM_OPT_KERNEL_TIMING) */
    int32_t search_launches;
    int32_t solve_launches;
    int32_t reserved;
} icpsymm_stats;

/* option keys for icpsymm_set_option */
enum {
    ICPSYMM_OPT_REL_DIFF_TOL = 1,   /* 0 = off (reference behaviour); stop when |d(k-1)-d(k)| <= tol*d(k-1) */
    ICPSYMM_OPT_IDENTITY_PAIRS = 2, /* 1 = pair row i with row i (what myicp.cpp:128-131 actually does) */
    ICPSYMM_OPT_MAX_PAIR_DIST = 3,  /* 0 = off; reject pairs farther than this */
    ICPSYMM_OPT_USE_GRAPH = 4,      /* 1 (default) = CUDA graph with a device-side WHILE node; 2 = unrolled graph
                                       ((max_iters+1) x [search, solve] with early exits); 0 = plain stream launches */
    ICPSYMM_OPT_KERNEL_TIMING = 5,  /* 1 = per-kernel CUDA-event timing (forces stream mode) */
    ICPSYMM_OPT_WARM_START = 6,     /* 1 (default) = seed each search with the previous correspondence; 2 = also seed the
                                       first search with the correspondences the previous call left for the SAME clouds
                                       (re-registration from another start; a seed is only a candidate: results stay exact) */
    ICPSYMM_OPT_TRACE = 7,          /* 1 = keep per-search transform + diff (icpsymm_get_trace) */
    ICPSYMM_OPT_GRID_BLOCKS = 8,    /* accepted and ignored: the tile -> block map of the search kernel is canonical (a function
                                       of the pair's point count only), which is what makes results independent of the GPU count */
    ICPSYMM_OPT_MIN_NORMAL_DOT = 10, /* <= -1 = off; reject pairs whose normals disagree: n_p . n_q < value */
    ICPSYMM_OPT_TILE_ORDER = 12,    /* 1 (default) = hand the 256-query tiles of the first search out most expensive first (a cost
                                       estimate per tile); 0 = curve order.  Scheduling only: results are identical */
    ICPSYMM_OPT_SKIP_WALK = 13,     /* 1 (default) = a 32-query unit whose nearest neighbours provably cannot have changed since its
                                       last walk (each query moved by less than the gap between its neighbour and everything
                                       else) keeps them without walking the index; 0 = always walk.
                                       Results are identical either way */
    ICPSYMM_OPT_PROBE = 14,         /* 1 (default) = a search that follows a long step (and the first one) begins with one exact
                                       single-query search per 256-query tile whose result seeds the whole tile; 0 = off.
                                       Seeds are candidates only: results are identical */
    ICPSYMM_OPT_PEER_TIMEOUT_MS = 11, /* NVLink peer exchange: how long a rank waits for its peers' sums (default 10000 ms)
                                        before the call fails with ICPSYMM_ERR_COMM */
    ICPSYMM_OPT_ASYNC_SETUP = 9     /* 1 = set_source / set_target / estimate_normals return without waiting: each cloud is
                                       uploaded, indexed and given normals on its own stream, so the upload of one cloud
                                       overlaps the build of the other; a NaN/Inf input is then reported by the next call
                                       that returns data (register, find_correspondences, get_*, synchronize) */
};

/* ---- lifetime (MyICP::MyICP / ~MyICP, myicp.cpp:6-18) ---- */
int icpsymm_create(icpsymm_ctx** ctx, int device_ordinal);
/* same, but all work is enqueued on the caller's cudaStream_t (passed as void*) */
int icpsymm_create_on_stream(icpsymm_ctx** ctx, int device_ordinal, void* cuda_stream);
void icpsymm_destroy(icpsymm_ctx* ctx);
const char* icpsymm_status_string(int status);
const char* icpsymm_last_error(const icpsymm_ctx* ctx);
void* icpsymm_get_stream(const icpsymm_ctx* ctx); /* the cudaStream_t all work of this context runs on */
int icpsymm_synchronize(icpsymm_ctx* ctx);
int icpsymm_set_option(icpsymm_ctx* ctx, int key, double value);
int icpsymm_device_info(int device_ordinal, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ---- clouds (MyICP::LoadCloud, myicp.cpp:20-31; pasteInMatrix, func.cpp:5-15) ----
 * Copies the points to the device, sorts them along a Hilbert curve and builds the search index.  Any normals
 * of that cloud are invalidated. */
int icpsymm_set_source(icpsymm_ctx* ctx, const float* xyz, size_t n, size_t stride_bytes);
int icpsymm_set_target(icpsymm_ctx* ctx, const float* xyz, size_t n, size_t stride_bytes);
int icpsymm_set_source_dev(icpsymm_ctx* ctx, const float* xyz_dev, size_t n, size_t stride_bytes);
int icpsymm_set_target_dev(icpsymm_ctx* ctx, const float* xyz_dev, size_t n, size_t stride_bytes);
/* batches of independent pairs: pair b occupies rows [b*max_count, b*max_count+counts[b]) of xyz;
 * counts == NULL means every cloud holds max_count points. */
int icpsymm_set_source_batch(icpsymm_ctx* ctx, const float* xyz, const uint32_t* counts, size_t n_pairs,
                             size_t max_count, size_t stride_bytes);
int icpsymm_set_target_batch(icpsymm_ctx* ctx, const float* xyz, const uint32_t* counts, size_t n_pairs,
                             size_t max_count, size_t stride_bytes);

/* ---- normals (MyICP::estimateNormals, myicp.cpp:152-172: NormalEstimation, setKSearch(10)) ----
 * exact kNN-k (self included), covariance, eigenvector of the smallest eigenvalue, flipped towards
 * `viewpoint` (NULL = (0,0,0), PCL's default), curvature in .w.  3 <= k <= 32. */
int icpsymm_estimate_normals(icpsymm_ctx* ctx, int k, const float viewpoint[3]);
/* which: 0 = source, 1 = target.  normals4: n x 4 floats in original point order. */
int icpsymm_set_normals(icpsymm_ctx* ctx, int which, const float* normals4);
int icpsymm_get_normals(icpsymm_ctx* ctx, int which, float* normals4);

/* ---- registration (MyICP::RegisterSymm, myicp.cpp:100-150, with the parameters its
 * "todo add params to specify iters & diff" asks for, myicp.h:19).  Reference defaults: 10, 1.0f. ----
 * while (diff > diff_threshold && iters++ < max_iters) { correspondences; estimateTransformSymm;
 * apply; transform = incre * transform; diff = evalDiff }.  T_out: 16 floats (n_pairs*16 for a batch),
 * stats: one record (n_pairs records for a batch); both may be NULL. */
int icpsymm_register_symm(icpsymm_ctx* ctx, int max_iters, float diff_threshold, float* T_out,
                          icpsymm_stats* stats);

/* The same registration in two halves, for callers that pipeline pairs: _async ENQUEUES the whole loop on the context's
 * stream and returns without waiting for the device; icpsymm_wait blocks until it is done and hands out what
 * icpsymm_register_symm would have returned.  With ICPSYMM_OPT_ASYNC_SETUP the set-up calls do not block either, so one
 * host thread can upload, index and start pair k+1 on a second context while pair k is still iterating on the first.
 * Exactly one registration may be in flight per context; any other call on that context must come after icpsymm_wait. */
int icpsymm_register_symm_async(icpsymm_ctx* ctx, int max_iters, float diff_threshold);
int icpsymm_wait(icpsymm_ctx* ctx, float* T_out, icpsymm_stats* stats);

/* ---- building blocks of the loop, exposed because the reference exposes them (func.h:5-27) ---- */
/* the missing search of myicp.cpp:128-131: exact 1-NN of T*source in target, lowest index on ties.
 * idx_out: n_source int32 (original target indices), d2_out (optional): fp32 squared distances. */
int icpsymm_find_correspondences(icpsymm_ctx* ctx, const float T[16], int32_t* idx_out, float* d2_out);
/* correspondences / distances of the last search run by icpsymm_register_symm */
int icpsymm_get_correspondences(icpsymm_ctx* ctx, int32_t* idx_out, float* d2_out);
/* evalDiff(src, tgt, indices) (func.cpp:36-40, a stub in the reference): sum ||T p - q_nn|| */
int icpsymm_eval_diff(icpsymm_ctx* ctx, const float T[16], double* diff);
/* calculateMatrixNotation + the normal equations (func.cpp:43-60): the 37 raw sums about `origin`
 * for the pairs found under T (pair 0 of a batch). */
int icpsymm_normal_equations(icpsymm_ctx* ctx, const float T[16], const double origin[3],
                             double sums[ICPSYMM_NSUMS]);
/* estimateTransformSymm's solve + composition (func.cpp:76-102) for a given sums vector, on device */
int icpsymm_solve_compose(icpsymm_ctx* ctx, const double sums[ICPSYMM_NSUMS], const double origin[3],
                          double Tinc[16], double x[6], double* min_pivot_ratio);
/* applyTransform (func.cpp:104-121) of the source by T, original order, `stride_bytes` between points */
int icpsymm_get_transformed_source(icpsymm_ctx* ctx, const float T[16], float* xyz_out, size_t stride_bytes);
/* per-search trace (ICPSYMM_OPT_TRACE): k = 0..searches-1 of pair `pair` */
int icpsymm_get_trace(icpsymm_ctx* ctx, size_t pair, int k, float T[16], double* diff);

/* per-launch device times of the last registration run with ICPSYMM_OPT_KERNEL_TIMING: search k (the
 * NN+accumulate kernel) and the reduce+solve kernel that followed it, in ms; either array may be NULL.
 * Returns the number of entries written (<= n), or a negative error. */
int icpsymm_get_search_times(icpsymm_ctx* ctx, float* search_ms, float* solve_ms, int n);

/* ---- closed-form point-to-point fit (registrateNPoint, regist.h:8-72): dst = R src + T, fp64 ---- */
int icpsymm_register_npoint(icpsymm_ctx* ctx, const double* src_xyz, const double* dst_xyz, size_t n,
                            double R[9], double T[3]);

/* ---- point-to-point ICP (MyICP::RegisterP2P, myicp.h:18; a stub in the reference, myicp.cpp:43-59): the loop of
 * myicp.cpp:121-142 with the closed-form fit of the matched pairs (registrateNPoint) as its step; single pairs, no normals
 * needed; correspondences, pair gathering, fit and composition all run on the device ---- */
int icpsymm_register_p2p(icpsymm_ctx* ctx, int max_iters, float diff_threshold, float T_out[16], icpsymm_stats* stats);

/* ---- multi-GPU.  Pair batches need nothing: each rank registers its own pairs.
 * Query sharding: every rank holds the full clouds; rank r of n_ranks <= 8 searches its slice of the sorted
 * source and the 37-value normal equations are summed across ranks before the solve.  The sums have a
 * canonical order (8 fixed parts of the source, a shard is a run of whole parts), so the
 * transform, the iteration count and the correspondences are BIT-IDENTICAL for 1, 2, 4 and 8 ranks and for
 * both exchange mechanisms below. ---- */
int icpsymm_comm_unique_id(void* id_out128);                       /* ncclGetUniqueId on the root */
int icpsymm_comm_init(icpsymm_ctx* ctx, const void* id128, int n_ranks, int rank);
int icpsymm_comm_destroy(icpsymm_ctx* ctx);
int icpsymm_set_query_shard(icpsymm_ctx* ctx, int rank, int n_ranks); /* without comm: shard only */
/* NCCL-free exchange for the query-sharded mode: the reduce+solve kernel itself stores its 37 sums
 * into every peer's exchange buffer over NVLink (CUDA IPC peer mapping), waits for the peers' sums
 * of the same iteration and adds them in rank order -- one kernel, usable inside the device-side
 * loop graph.  Each rank exports a 64-byte handle, the host application gathers the handles of all
 * ranks (any transport) and every rank attaches with the full table.  Single pairs only.
 * Attach and detach are collective (all ranks, each followed by a host-side barrier of the application before
 * the first sharded registration); they reset the exchange sequence.  Only icpsymm_register_symm is collective:
 * the single-search calls (find_correspondences, eval_diff, normal_equations) work on the rank's own shard.
 * A peer that does not publish within ICPSYMM_OPT_PEER_TIMEOUT_MS fails the call with ICPSYMM_ERR_COMM. */
int icpsymm_peer_export(icpsymm_ctx* ctx, void* handle64_out);
int icpsymm_peer_attach(icpsymm_ctx* ctx, const void* handles /* n_ranks x 64 bytes */, int n_ranks, int rank);
int icpsymm_peer_detach(icpsymm_ctx* ctx);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ICPSYMM_H */
