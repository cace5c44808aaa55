// icp_loop.cu -- the registration loop of MyICP::RegisterSymm (ICP/myicp.cpp:100-150) on device.
//
// One iteration = up to three kernels on the context stream:
//   probe_kernel           (after a long step only) one budgeted single-query search per 256-query tile; its
//                          neighbour seeds the whole tile in K1.
//   nn_accum_kernel  (K1)  for every source point: apply the accumulated transform on load
//                          (applyTransform, func.cpp:104-121, never materialised), exact nearest
//                          neighbour in the target hierarchy (the search myicp.cpp:128-131 leaves
//                          as a todo) -- or the proof that the previous neighbour still stands (walk
//                          skipping) --, evalDiff term (func.cpp:19-40), and the per-pair rows
//                          m = (p+q) x n, n = n_p+n_q, c = (p-q).n of calculateMatrixNotation
//                          (func.cpp:43-60) folded straight into the 27 unique normal-equation
//                          sums (+ centroid, diff and count terms: 37 values), reduced by warp
//                          shuffles to one fp64 row per 256-query tile.
//   solve_kernel     (K2)  canonical fp64 reduction of the tile rows (8 parts, fixed order: the result
//                          does not depend on the grid, the batch or the number of GPUs), the loop
//                          test of myicp.cpp:123, re-centring to the pair centroids, the 6x6 solve that
//                          replaces the two JacobiSVD least-squares solves (func.cpp:64-73,87-88),
//                          theta = atan|a| and the half-rotation composition (func.cpp:91-99, in
//                          the order of the comment at :94), transform = incre * transform
//                          (myicp.cpp:138); decides probe / margin recording for the next search.
// The loop is one CUDA graph whose WHILE node repeats [probe,K1,K2] until K2 clears the condition on
// the device (default), or an unrolled graph of (max_iters+1) x [probe,K1,K2] in which a device-side
// `done` flag in SegState turns the remaining launches into early exits, or plain stream launches.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>

#include "internal.h"
#ifndef ICP_WIDE_LEAF
#define ICP_WIDE_LEAF 1   // 256-bit loads in the walks of this translation unit (search.cuh)
#endif
#ifndef ICP_WIDE_NODE
#define ICP_WIDE_NODE 1
#endif
#include "search.cuh"

#ifdef ICP_BLOCK_TIMES   // developer instrumentation: per-block start/end/SM of the last K1 launch
__device__ unsigned long long icp_k1_times[3 * 8192];
struct K1Timer {
    unsigned long long t0;
    __device__ K1Timer() { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); }
    __device__ void stop() {
        if (threadIdx.x == 0 && blockIdx.x < 8192) {
            unsigned long long t1; unsigned int sm;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            icp_k1_times[3 * blockIdx.x] = t0; icp_k1_times[3 * blockIdx.x + 1] = t1; icp_k1_times[3 * blockIdx.x + 2] = sm;
        }
    }
};
__device__ unsigned long long icp_k2_stamps[8];
#define ICP_K2_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); icp_k2_stamps[i] = t_; } } while (0)
extern "C" __attribute__((visibility("default"))) int icpsymm_debug_k2_stamps(unsigned long long* out) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, icp_k2_stamps, sizeof(unsigned long long) * 8);
}
extern "C" __attribute__((visibility("default"))) int icpsymm_debug_k1_times(unsigned long long* out, int n) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, icp_k1_times, sizeof(unsigned long long) * 3 * n);
}
#else
#define ICP_K2_STAMP(i) do { } while (0)
#endif

#ifdef ICP_TILE_TIMES   // developer instrumentation: start/end of every tile of the last K1 launch
__device__ unsigned long long icp_tile_times[2 * 65536];
extern "C" __attribute__((visibility("default"))) int icpsymm_debug_tile_times(unsigned long long* out, int n) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, icp_tile_times, sizeof(unsigned long long) * 2 * n);
}
#endif

namespace {

constexpr int kNNThreads = ICP_TILE;   // one block = one 256-query tile at a time
#ifndef ICP_K1_BLOCKS
#define ICP_K1_BLOCKS (1024 / ICP_TILE)   // 64 registers per thread: 1024 threads per SM (5 blocks = 51 registers: measured)
#endif
constexpr int kNNBlocksPerSM = ICP_K1_BLOCKS;
#ifndef ICP_TILES_PER_TICKET
#define ICP_TILES_PER_TICKET 1
#endif
constexpr int kTilesPerTicket = ICP_TILES_PER_TICKET;
#ifndef ICP_TILES_PER_TICKET_SKIP
#define ICP_TILES_PER_TICKET_SKIP 2
#endif
constexpr int kTilesPerTicketSkip = ICP_TILES_PER_TICKET_SKIP > ICP_TILES_PER_TICKET ? ICP_TILES_PER_TICKET_SKIP : ICP_TILES_PER_TICKET;
#ifndef ICP_PROBE_BUDGET
#define ICP_PROBE_BUDGET 24   // leaf scans a probe may spend (an ordinary far query needs about 40; 24, 48, 96 measured alike)
#endif
#ifndef ICP_PROBE_STEP_SPACINGS
#define ICP_PROBE_STEP_SPACINGS 64.0   // probe the next search when a step moves some point by more than this many point spacings
#endif
#ifndef ICP_TRACK_STEP_SPACINGS
#define ICP_TRACK_STEP_SPACINGS 4.0   // record margins once a step moves no point by more than this many point spacings
#endif
constexpr int kNNWarps = kNNThreads / 32;
#ifndef ICP_K2_THREADS
#define ICP_K2_THREADS 1024
#endif
constexpr int kSolveThreads = ICP_K2_THREADS;
constexpr int kSolveParts = kSolveThreads / ICP_NSUMS;  // 27
#ifndef ICP_K2_BATCH
#define ICP_K2_BATCH 16  // loads of the K2 row reduction issued back to back
#endif

struct NNParams {
    CloudView src, tgt;
    SegState* state;
    const LoopParams* lp;
    uint32_t* corr_pos;
    double* partial;          // [pair][tiles_max] rows of ICP_PART_STRIDE doubles: one row per 256-query tile
    LoopCtrl* ctrl;           // ticket counter of the dynamic tile hand-out
    const uint32_t* order;    // ticket -> work item (null: identity)
    // walk skipping (see the skip test in K1): per query the margin of its last tracked walk and the search in which
    // that walk happened, per pair the fp32 transform of every search so far
    uint32_t* tile_seed;      // [pair][tiles_max]: sorted target position of the neighbour of the tile's middle query (probe)
    float* margin2;
    uint8_t* ref_search;      // 0xff = the query has no margin (never walked with tracking in this registration)
    const float* hist;
    uint32_t hist_stride;     // searches kept per pair
    uint32_t tiles_max;       // rows reserved per pair = icp_tiles(largest pair)
    uint32_t part0, nparts;   // parts of every pair this launch covers: 0, 8 = whole pairs; a query shard = a run of parts
    uint32_t work_per_seg;    // tile slots per pair handed out by this launch (the parts' tile range of the largest pair)
    uint32_t total_work;      // pairs x work_per_seg
    uint32_t grid;            // persistent blocks of the launch
};

struct SolveParams {
    cudaGraphConditionalHandle cond;   // while-node condition (use_cond != 0)
    LoopCtrl* ctrl;
    int use_cond;
    SegState* state;
    const LoopParams* lp;
    const double* partial;
    const uint32_t* src_counts;
    double* gsum;             // [pair][ICP_PART_STRIDE]: totals of the local parts (single searches)
    double* gparts;           // [pair][ICP_NPARTS][ICP_PART_STRIDE]: part sums, exchange buffer of the NCCL mode
    float* hist;              // [pair][hist_stride][12]: the fp32 transform of every search (K1's skip test)
    uint32_t hist_stride;
    float* trace_T;
    double* trace_diff;
    uint32_t tiles_max, part0, nparts, k1_grid;
    int trace_stride;
    int mode;  // 0 reduce+solve, 1 reduce only (-> gparts), 2 solve from gparts, 3 reduce + record (-> gsum), no solve
    PeerBuf* peers[ICP_MAX_PEERS];   // NVLink peer mode: exchange buffers of all ranks (mine included)
    int peer_world, peer_rank;
    unsigned long long peer_timeout_ns;
};

struct InitArgs {
    float T[12];
    double origin[3];
    int has_T, has_origin;
    int probe;   // the first search is probed
};

__global__ void set_lp_kernel(LoopParams* dst, LoopParams v) { *dst = v; }

__global__ void init_state_kernel(SegState* state, uint32_t nseg, InitArgs a, const float4* __restrict__ bb_lo,
                                  const float4* __restrict__ bb_hi, const uint32_t* __restrict__ tgt_counts, LoopCtrl* ctrl,
                                  uint32_t k1_grid, float* hist, uint32_t hist_stride) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) ctrl->work = k1_grid;   // tickets below the grid size are implicit
    if (s >= nseg) return;
    SegState st;
    memset(&st, 0, sizeof st);
    for (int i = 0; i < 12; i++) {
        float v = a.has_T ? a.T[i] : ((i % 5 == 0) ? 1.f : 0.f);
        st.T[i] = (double)v;
        st.Tf[i] = v;
        if (hist) hist[(size_t)s * hist_stride * 12 + i] = v;   // the transform of search 0
    }
    float4 lo = bb_lo[s], hi = bb_hi[s];
    // the origin of the raw sums must be exactly representable in fp32 (K1 subtracts it in fp32)
    st.origin[0] = a.has_origin ? (double)(float)a.origin[0] : (double)(0.5f * (lo.x + hi.x));
    st.origin[1] = a.has_origin ? (double)(float)a.origin[1] : (double)(0.5f * (lo.y + hi.y));
    st.origin[2] = a.has_origin ? (double)(float)a.origin[2] : (double)(0.5f * (lo.z + hi.z));
    st.min_ratio = 1.0;
    st.probe = a.probe;
    const double ex = (double)hi.x - lo.x, ey = (double)hi.y - lo.y, ez = (double)hi.z - lo.z;
    st.radius = 0.5 * sqrt(ex * ex + ey * ey + ez * ez);
    st.spacing = 2.0 * st.radius / sqrt((double)(tgt_counts[s] ? tgt_counts[s] : 1u));
    state[s] = st;
}

// ------------------------------------------------------------------------------------------------
// Tile order of the first search.  The cost of an exact search is set by how much of the target lies
// near the query's ball: a patch close to the medial axis of the target (inside a bulge, 30 degrees off)
// sees hundreds of candidate boxes at almost the same distance and takes 10-20x longer than a patch just
// off the surface (measured: median tile 126 us, worst 930 us in a 1.5 ms launch).  With tiles handed out
// in curve order such tiles start whenever the curve reaches them and the launch ends with a long tail.
// One warp per tile estimates the cost up front -- r = an upper bound of the NN distance of the tile's
// middle query (min over the boxes of one coarse level of the distance to the farthest corner), cost =
// the number of boxes of that level within r -- and the tiles are handed out most expensive first
// (stable radix sort: equal costs keep curve order, so neighbouring blocks still share lines).
// ------------------------------------------------------------------------------------------------
constexpr int kCostBoxesMax = 4096;   // boxes of the coarse level staged in shared memory (96 KB)
__global__ void __launch_bounds__(256) tile_cost_kernel(CloudView src, CloudView tgt, const SegState* __restrict__ state,
                                                        uint32_t part0, uint32_t nparts, uint32_t work, int rec_level,
                                                        uint32_t tiles_per_warp, uint32_t* __restrict__ key_out,
                                                        uint32_t* __restrict__ val_out) {
    // the boxes of the coarse level, once per block: [6][nbox] floats (lo xyz, hi xyz); padding boxes have lo > hi
    extern __shared__ float sbox[];
    const uint32_t nbox = min(tgt.shape.level_cnt[rec_level] * ICP_FAN, (uint32_t)kCostBoxesMax);
    {
        const float* recs = reinterpret_cast<const float*>(tgt.nodes + tgt.shape.level_off[rec_level]);
        for (uint32_t t = threadIdx.x; t < nbox * 6u; t += 256u) {
            // record r holds 48 floats: lox[8] loy[8] loz[8] hix[8] hiy[8] hiz[8]
            const uint32_t r = t / 48u, k = (t % 48u) / 8u, c = t % 8u;
            sbox[k * nbox + r * 8u + c] = __ldg(recs + t);
        }
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp_global = (blockIdx.x * 256u + threadIdx.x) >> 5;
    const uint32_t ns = src.counts[0], tiles = icp_tiles(ns);
    const float *blx = sbox, *bly = sbox + nbox, *blz = sbox + 2 * nbox, *bhx = sbox + 3 * nbox, *bhy = sbox + 4 * nbox,
                *bhz = sbox + 5 * nbox;
    for (uint32_t k = 0; k < tiles_per_warp; k++) {
        const uint32_t wid = warp_global * tiles_per_warp + k;
        if (wid >= work) return;
        const uint32_t tile = icp_shard_tile(tiles, part0, nparts, wid);
        uint32_t cost = 0u;
        if (tile < tiles && tgt.counts[0] > 0) {
            uint32_t i = tile * ICP_TILE + ICP_TILE / 2;
            if (i >= ns) i = ns - 1;
            const float4 sp = ldg4(src.pts + i);
            const float3 q = apply_rn(state[0].Tf, sp.x, sp.y, sp.z);
            float r2 = __int_as_float(0x7f800000);
            for (uint32_t b = lane; b < nbox; b += 32u) {
                const float lx = blx[b], ly = bly[b], lz = blz[b], hx = bhx[b], hy = bhy[b], hz = bhz[b];
                if (lx <= hx) {   // not a padding box
                    const float fx = fmaxf(fabsf(q.x - lx), fabsf(q.x - hx)), fy = fmaxf(fabsf(q.y - ly), fabsf(q.y - hy));
                    const float fz = fmaxf(fabsf(q.z - lz), fabsf(q.z - hz));
                    r2 = fminf(r2, fx * fx + fy * fy + fz * fz);
                }
            }
            r2 = warp_min_f(r2);
            for (uint32_t b = lane; b < nbox; b += 32u) {
                const float lx = blx[b], ly = bly[b], lz = blz[b], hx = bhx[b], hy = bhy[b], hz = bhz[b];
                const float ex = fmaxf(fmaxf(lx - q.x, q.x - hx), 0.f), ey = fmaxf(fmaxf(ly - q.y, q.y - hy), 0.f);
                const float ez = fmaxf(fmaxf(lz - q.z, q.z - hz), 0.f);
                cost += (lx <= hx && ex * ex + ey * ey + ez * ez <= r2) ? 1u : 0u;
            }
            cost = __reduce_add_sync(0xffffffffu, cost);
        }
        if (lane == 0) {
            key_out[wid] = 0xffffu - (cost > 0xffffu ? 0xffffu : cost);   // ascending key = descending cost
            val_out[wid] = wid;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1: search + accumulate
// ------------------------------------------------------------------------------------------------
// PROBE.  While the clouds are far apart a search starts blind: the previous correspondences are useless as bounds
// (the last step moved every point a long way) and a packet pays for that until its first good candidates turn up
// (measured: a far search seeded with a neighbouring query's true neighbour is 20-25 % cheaper).  So before such a
// search one warp per 256-query tile finds the exact neighbour of the tile's MIDDLE query -- the ordinary walk with a
// single live lane, about 1/3 of a packet walk for 1/8 of the units -- and K1 offers that point to all 256 queries of
// the tile first.  A seed is only ever a candidate: the result stays exact.  K2 turns the probe off once a step is small.
__global__ void __launch_bounds__(kNNThreads, kNNBlocksPerSM) probe_kernel(const __grid_constant__ NNParams P) {
    const uint32_t w = (blockIdx.x * (uint32_t)kNNThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (w >= P.total_work) return;
    const uint32_t seg = w / P.work_per_seg;
    const SegState* st = P.state + seg;
    if (st->done || !st->probe) return;
    const uint32_t ns = P.src.counts[seg], nt = P.tgt.counts[seg];
    const uint32_t tiles = icp_tiles(ns);
    const uint32_t tile = icp_shard_tile(tiles, P.part0, P.nparts, w - seg * P.work_per_seg);
    if (tile >= tiles || nt == 0) return;
    ICP_ASSERT(tile < P.tiles_max);
    uint32_t i = tile * ICP_TILE + ICP_TILE / 2;
    if (i >= ns) i = ns - 1;
    const float4 s = ldg4(P.src.pts + (size_t)seg * P.src.shape.pitch + i);
    const float3 p = apply_rn(st->Tf, s.x, s.y, s.z);
    const float4* tpts = P.tgt.pts + (size_t)seg * P.tgt.shape.pitch;
    Best b;
    best_init(b);
    const uint32_t prev = P.corr_pos[(size_t)seg * P.src.shape.pitch + i];
    if (prev < nt) best_offer(b, ldg4(tpts + prev), p.x, p.y, p.z);
    const unsigned long long key = nn_walk_budget(&P.tgt.shape, P.tgt.nodes + (size_t)seg * P.tgt.shape.nodes_per_seg, tpts, p.x,
                                                  p.y, p.z, lane == 0, b.key, ICP_PROBE_BUDGET);
    if (lane == 0) {
        const uint32_t id = (uint32_t)key;
        P.tile_seed[(size_t)seg * P.tiles_max + tile] =
            id != (uint32_t)ICP_INVALID_ID ? __ldg(P.tgt.pos_of_orig + (size_t)seg * P.tgt.shape.pitch + id) : 0xffffffffu;
    }
}

// Persistent blocks (4 per SM) draw 256-query tiles from a global ticket counter; ticket k is tile
// order[k] (most expensive first, see tile_cost_kernel) or simply tile k.  The 8 warps of a block always
// work on the SAME tile: they share its node records and leaf lines in L1, which is worth more than
// letting them run apart (measured: warps decoupled over the block's tiles +20 % time).  An SM that drew
// cheap tiles draws more of them -- the cost of a far search varies 3-5x over a cloud.  Which block
// handles a tile does not touch the result: every tile leaves its own fp64 row (the 8 warps' fp32 shuffle
// totals added in warp order), and K2 adds the rows in a fixed order (common.cuh).
__global__ void __launch_bounds__(kNNThreads, kNNBlocksPerSM) nn_accum_kernel(const __grid_constant__ NNParams P) {
#ifdef ICP_BLOCK_TIMES
    K1Timer timer_;
#endif
    __shared__ float sT[12];
    __shared__ float sO[3];
    // the warps' fp32 totals; double-buffered by ticket parity: one barrier per ticket.  Kept small on purpose: shared
    // memory comes out of the SM's L1 (4 blocks x this), and the walk lives on its L1 hit rate
    __shared__ float wacc[2][kTilesPerTicketSkip][kNNWarps][ICP_PART_STRIDE];
    __shared__ int sDone;
    __shared__ uint32_t sSearch;      // index of the running search of the block's current pair
    __shared__ int sTrack;            // walks of this search record margins (K2 expects the next step to be small)
    __shared__ int sProbe;            // the probe kernel has left one seed per tile
    __shared__ uint32_t sWork[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const LoopParams lp = *P.lp;
    uint32_t cur_seg = 0xffffffffu;
    // tickets 0 .. gridDim-1 are implicit (the counter starts at gridDim); the ticket of the NEXT tile is drawn
    // at the start of the current one, so the latency of the atomic hides behind the tile's work
    uint32_t w = blockIdx.x;
    const bool use_order = P.order != nullptr && P.state[0].searches == 0;   // the order is made for the first search
    // a ticket is one tile while the searches walk (fine grain: the tiles of a far search differ 20x in cost; two
    // tiles per ticket measured 4-5 % slower even near convergence), and two once margins exist (K1 skip test): most warps then only re-check their winners (a few dependent
    // loads) while one or two of the eight walk, and with one tile per barrier the block spends its time waiting --
    // ncu: barrier stall 6-7 cycles per instruction, issue slots 40 % busy.  (launch-uniform: every block reads pair 0)
    const uint32_t tpt = use_order ? 1u : (P.state[0].have_margins ? (uint32_t)kTilesPerTicketSkip : (uint32_t)kTilesPerTicket);
    const uint32_t tickets_per_seg = (P.work_per_seg + tpt - 1u) / tpt;
    const uint32_t total_tickets = (P.total_work / P.work_per_seg) * tickets_per_seg;
    for (uint32_t it = 0; w < total_tickets; it++) {
        uint32_t wnext = 0u;
        if (tid == 0) wnext = atomicAdd(&P.ctrl->work, 1u);
        const uint32_t we = use_order ? __ldg(P.order + w) : w;
        const uint32_t seg = we / tickets_per_seg;
        if (seg != cur_seg) {   // block-uniform: (re)load the pair's transform and origin
            const SegState* st = P.state + seg;
            if (tid == 0) { sDone = st->done; sSearch = (uint32_t)st->searches; sTrack = st->track; sProbe = st->probe; }
            if (tid < 12) sT[tid] = st->Tf[tid];
            if (tid < 3) sO[tid] = (float)st->origin[tid];
            __syncthreads();
            cur_seg = seg;
        }
        const uint32_t ns = P.src.counts[seg], nt = P.tgt.counts[seg];
        const uint32_t tiles = icp_tiles(ns);
        const uint32_t item0 = (we - seg * tickets_per_seg) * tpt;   // work items of this ticket: item0 .. item0 + tpt - 1
#ifdef ICP_TILE_TIMES
        unsigned long long tt0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt0));
#endif
#pragma unroll 1
        for (uint32_t sub = 0; sub < tpt; sub++) {
        const uint32_t tile = icp_shard_tile(tiles, P.part0, P.nparts, item0 + sub);
        // block-uniform; a finished pair has nothing to do and a smaller pair of a batch has fewer tiles
        const bool valid = !sDone && item0 + sub < P.work_per_seg && tile < tiles;
        float (*acc)[ICP_PART_STRIDE] = wacc[it & 1u][sub];
        ICP_ASSERT(seg < P.src.nseg && (!valid || (size_t)tile * ICP_TILE < P.src.shape.pitch + ICP_TILE));
        if (valid) {
        ICP_ASSERT(tile < P.tiles_max);
        const float4* __restrict__ spts = P.src.pts + (size_t)seg * P.src.shape.pitch;
        const float4* __restrict__ snrm = P.src.nrm + (size_t)seg * P.src.shape.pitch;
        const float4* __restrict__ tpts = P.tgt.pts + (size_t)seg * P.tgt.shape.pitch;
        const float4* __restrict__ tnrm = P.tgt.nrm + (size_t)seg * P.tgt.shape.pitch;
        const WideNode* __restrict__ tbox = P.tgt.nodes + (size_t)seg * P.tgt.shape.nodes_per_seg;
        uint32_t* corr = P.corr_pos + (size_t)seg * P.src.shape.pitch;

        const uint32_t i = tile * kNNThreads + tid;
        const bool active = i < ns && nt > 0;
        float3 p = make_float3(0.f, 0.f, 0.f);
        Best b;
        best_init(b);
        // SKIP TEST.  The last tracked walk of a query (search `ref`, query position x_ref = T_ref p) left its winner and
        // its margin m: every other target point is at least m away from x_ref.  The query now sits at x = T p,
        // delta = |x - x_ref| away, so every other point is at least m - delta from x (triangle inequality), while
        // the old winner is d1 = |x - q| away.  If d1 < m - delta the winner cannot have changed: the lane keeps it
        // and sits out the walk; a unit whose 32 lanes all keep theirs does not walk at all.  The three lengths are
        // evaluated in fp32 with the rounding directed against keeping and a relative slack of 1e-5, three orders
        // above the rounding of the pinned distance formula, so the statement also holds for the fp32 values the
        // oracle compares (a tie has zero margin and never keeps).
        const bool may_skip = lp.skip_walk && lp.warm_start && !lp.identity_pairs;
        bool keep = false;   // this lane's winner provably stands
        // the previous correspondent and its normal: a lane that keeps its winner accumulates with them as they are,
        // without the two dependent gathers (inverse permutation, then point and normal) -- the searches in which most
        // lanes keep are bound by exactly that chain of loads
        uint32_t prev = 0xffffffffu;
        float4 sq = make_float4(0.f, 0.f, 0.f, 0.f), snq = sq, sn0 = sq;
        if (active) {
            const float4 s = ldg4(spts + i);
            p = apply_rn(sT, s.x, s.y, s.z);
            if (lp.identity_pairs) {
                uint32_t pos = P.tgt.pos_of_orig[(size_t)seg * P.tgt.shape.pitch + __float_as_uint(s.w)];
                best_offer(b, ldg4(tpts + pos), p.x, p.y, p.z);
            } else {
                prev = lp.warm_start ? corr[i] : 0xffffffffu;
                if (prev < nt) { sq = ldg4(tpts + prev); best_offer(b, sq, p.x, p.y, p.z); }
                if (sProbe) {   // the neighbour of the tile's middle query (probe_kernel): a good first bound for all 256
                    const uint32_t sp = P.tile_seed[(size_t)seg * P.tiles_max + tile];
                    if (sp < nt) best_offer(b, ldg4(tpts + sp), p.x, p.y, p.z);
                }
                const uint32_t ref = may_skip ? (uint32_t)P.ref_search[(size_t)seg * P.src.shape.pitch + i] : 0xffu;
                ICP_ASSERT(ref == 0xffu || ref < sSearch);
                if (ref < P.hist_stride && ref != 0xffu && prev < nt) {
                    snq = ldg4(tnrm + prev);   // issued with the loads of the test, used only if the lane keeps
                    sn0 = ldg4(snrm + i);
                    const float3 xr = apply_rn(P.hist + ((size_t)seg * P.hist_stride + ref) * 12, s.x, s.y, s.z);
                    const float delta = __fsqrt_ru(dist2_rn(p.x, p.y, p.z, xr.x, xr.y, xr.z));
                    const float d1 = __fsqrt_ru(b.d2());
                    const float m = __fsqrt_rd(P.margin2[(size_t)seg * P.src.shape.pitch + i]);
                    keep = __fmul_ru(__fadd_ru(d1, delta), 1.00001f) < __fmul_rd(m, 0.99999f);
                }
            }
        }
        // warp-synchronous exact search: the 32 lanes are 32 neighbouring source points; lanes that keep their winner
        // are bystanders (split units, see finish_units_kernel in index_build.cu, are not walked per cluster here:
        // wrapping this walk in a pass loop costs the common case 15 %, far more than the rare split unit costs)
        const bool walk = active && !keep;
        // the winner's sorted position, point and normal.  A unit that walks looks them up afterwards (the key carries the
        // original index; the sorted position comes from the inverse permutation); in a unit that does not walk -- every
        // lane keeps, or identity pairing -- a keeping lane already holds all three (nothing of it is live across the walk)
        uint32_t bpos = 0xffffffffu;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f), nq = q, n0 = q;
        auto look_up = [&]() {
            bpos = (b.id() != (uint32_t)ICP_INVALID_ID) ? __ldg(P.tgt.pos_of_orig + (size_t)seg * P.tgt.shape.pitch + b.id())
                                                        : 0xffffffffu;
            ICP_ASSERT(bpos == 0xffffffffu || bpos < nt);
            corr[i] = bpos;
            if (bpos < nt) { q = ldg4(tpts + bpos); nq = ldg4(tnrm + bpos); n0 = ldg4(snrm + i); }
        };
        if (!lp.identity_pairs && nt > 0 && __any_sync(0xffffffffu, walk)) {
            uint8_t* ref_slot = P.ref_search + (size_t)seg * P.src.shape.pitch + i;
            if (sTrack && may_skip && sSearch < 0xffu) {
                const WalkResult r = nn_walk<true>(&P.tgt.shape, tbox, tpts, p.x, p.y, p.z, walk, b.key);
                b.key = r.key;
                if (walk) { P.margin2[(size_t)seg * P.src.shape.pitch + i] = __uint_as_float(r.m2); *ref_slot = (uint8_t)sSearch; }
            } else {
                const WalkResult r = nn_walk<false>(&P.tgt.shape, tbox, tpts, p.x, p.y, p.z, walk, b.key);
                b.key = r.key;
                if (walk && may_skip) *ref_slot = 0xffu;   // walked without tracking: the old margin belongs to another winner
            }
            if (active) look_up();
        } else if (active) {
            if (keep && b.id() == __float_as_uint(sq.w)) { bpos = prev; q = sq; nq = snq; n0 = sn0; }   // corr[i] already holds it
            else look_up();
        }
        float v[40];   // declared after the walk so that none of it is live across it; ICP_NSUMS = 37 used
#pragma unroll
        for (int k = 0; k < 40; k++) v[k] = 0.f;
        if (active) {
            const float bd2 = b.d2();
            if (bpos < nt) {
                const float3 np = rotate_rn(sT, n0.x, n0.y, n0.z);
                v[33] = __fsqrt_rn(bd2);
                v[34] = bd2;
                if (bd2 <= lp.max_pair_d2 && (np.x * nq.x + np.y * nq.y + np.z * nq.z) >= lp.min_normal_dot) {
                    const float nx = np.x + nq.x, ny = np.y + nq.y, nz = np.z + nq.z;
                    const float ppx = p.x - sO[0], ppy = p.y - sO[1], ppz = p.z - sO[2];
                    const float qqx = q.x - sO[0], qqy = q.y - sO[1], qqz = q.z - sO[2];
                    const float sx = ppx + qqx, sy = ppy + qqy, sz = ppz + qqz;
                    const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
                    const float mx = sy * nz - sz * ny, my = sz * nx - sx * nz, mz = sx * ny - sy * nx;
                    const float r = dx * nx + dy * ny + dz * nz;
                    v[0] = mx * mx; v[1] = mx * my; v[2] = mx * mz; v[3] = my * my; v[4] = my * mz; v[5] = mz * mz;
                    v[6] = mx * nx; v[7] = mx * ny; v[8] = mx * nz;
                    v[9] = my * nx; v[10] = my * ny; v[11] = my * nz;
                    v[12] = mz * nx; v[13] = mz * ny; v[14] = mz * nz;
                    v[15] = nx * nx; v[16] = nx * ny; v[17] = nx * nz; v[18] = ny * ny; v[19] = ny * nz; v[20] = nz * nz;
                    v[21] = mx * r; v[22] = my * r; v[23] = mz * r;
                    v[24] = nx * r; v[25] = ny * r; v[26] = nz * r;
                    v[27] = ppx; v[28] = ppy; v[29] = ppz;
                    v[30] = qqx; v[31] = qqy; v[32] = qqz;
                    v[35] = r * r;
                    v[36] = 1.f;
                }
            }
        }
        // transposing warp reduction: after the steps 16, 8, 4, 2, 1 lane L holds the warp total of
        // v[L] (31 shuffles for 32 values instead of 160): a lane keeps the half of the values whose
        // index bit matches its own lane bit and hands the other half to its partner
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const bool hi = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < off; j++) {
                const float keep = hi ? v[j + off] : v[j];
                const float send = hi ? v[j] : v[j + off];
                v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        // slots 32..39 (five used): transpose over lane bits 4, 3, 2, then a butterfly over bits 1, 0
#pragma unroll
        for (int off = 16, n = 4; n >= 1; off >>= 1, n >>= 1) {
            const bool hi = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < n; j++) {
                const float keep = hi ? v[32 + j + n] : v[32 + j];
                const float send = hi ? v[32 + j] : v[32 + j + n];
                v[32 + j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        v[32] += __shfl_xor_sync(0xffffffffu, v[32], 2);
        v[32] += __shfl_xor_sync(0xffffffffu, v[32], 1);
        // every lane leaves its own total; the tile's row = the 8 warp totals added in warp order (fp64)
        acc[warp][lane] = v[0];
        if ((lane & 3) == 0) acc[warp][32 + (lane >> 2)] = v[32];
        }   // valid
        }   // sub
        if (tid == 0) sWork[it & 1u] = wnext;
        __syncthreads();
        if (!sDone && tid < (int)(tpt * ICP_PART_STRIDE)) {
            const uint32_t sub = tid / ICP_PART_STRIDE, vi = tid - sub * ICP_PART_STRIDE;
            const uint32_t tile = icp_shard_tile(tiles, P.part0, P.nparts, item0 + sub);
            if (vi < ICP_NSUMS && item0 + sub < P.work_per_seg && tile < tiles) {
                double s = 0.0;
#pragma unroll
                for (int ww = 0; ww < kNNWarps; ww++) s += (double)wacc[it & 1u][sub][ww][vi];
                P.partial[((size_t)seg * P.tiles_max + tile) * ICP_PART_STRIDE + vi] = s;
            }
        }
#ifdef ICP_TILE_TIMES
        if (tid == 0 && item0 < 65536u) {
            unsigned long long tt1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt1));
            icp_tile_times[2 * item0] = tt0; icp_tile_times[2 * item0 + 1] = tt1;
        }
#endif
        w = sWork[it & 1u];
    }
#ifdef ICP_BLOCK_TIMES
    timer_.stop();
#endif
}

// ------------------------------------------------------------------------------------------------
// K2: reduce + decide + solve + compose
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat34_mul(const double (&A)[12], const double (&B)[12], double (&C)[12]) {
    // C = A*B as 4x4 affine matrices in 3x4 storage
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            double s = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
            if (c == 3) s += A[4 * r + 3];
            C[4 * r + c] = s;
        }
    }
}

// Re-centre the raw sums (taken about `origin`) to the pair centroids, solve, compose.
// Tinc: 3x4 row-major.  Returns 1 when a pivot was dropped.  Every loop has compile-time bounds
// and is fully unrolled so that the 6x6 system lives in registers (one thread runs this; with
// run-time indexed arrays it spent ~25 us in local memory).
__device__ __forceinline__ int solve_compose_dev(const double* __restrict__ S, const double (&origin)[3],
                                                 double (&Tinc)[12], double (&x)[6], double& ratio_out) {
#pragma unroll
    for (int i = 0; i < 12; i++) Tinc[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = 0.0;
    ratio_out = 0.0;
    const double cnt = S[36];
    if (cnt < 1.0) return 1;
    const double A[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};
    double B[9];
#pragma unroll
    for (int i = 0; i < 9; i++) B[i] = S[6 + i];
    const double C[9] = {S[15], S[16], S[17], S[16], S[18], S[19], S[17], S[19], S[20]};
    const double icnt = 1.0 / cnt;
    double dp[3], dq[3], de[3], ep[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        dp[k] = S[27 + k] * icnt;
        dq[k] = S[30 + k] * icnt;
        de[k] = dp[k] + dq[k];
        ep[k] = dp[k] - dq[k];
    }
    // m = m' - K n with K = skew(de);  r = r' - ep.n
    const double K[9] = {0, -de[2], de[1], de[2], 0, -de[0], -de[1], de[0], 0};
    double BKt[9], KC[9], KCKt[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            BKt[3 * r + c] = B[3 * r] * K[3 * c] + B[3 * r + 1] * K[3 * c + 1] + B[3 * r + 2] * K[3 * c + 2];
            KC[3 * r + c] = K[3 * r] * C[c] + K[3 * r + 1] * C[3 + c] + K[3 * r + 2] * C[6 + c];
        }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
            KCKt[3 * r + c] = KC[3 * r] * K[3 * c] + KC[3 * r + 1] * K[3 * c + 1] + KC[3 * r + 2] * K[3 * c + 2];
    double H[36];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double at = A[3 * r + c] - BKt[3 * r + c] - BKt[3 * c + r] + KCKt[3 * r + c];
            const double bt = B[3 * r + c] - KC[3 * r + c];
            H[6 * r + c] = at;
            H[6 * r + 3 + c] = bt;
            H[6 * (3 + c) + r] = bt;
            H[6 * (3 + r) + 3 + c] = C[3 * r + c];
        }
    double Ce[3], Be[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        Ce[r] = C[3 * r] * ep[0] + C[3 * r + 1] * ep[1] + C[3 * r + 2] * ep[2];
        Be[r] = B[3 * r] * ep[0] + B[3 * r + 1] * ep[1] + B[3 * r + 2] * ep[2];
    }
    const double bn[3] = {S[24], S[25], S[26]};
    const double Kbn[3] = {de[1] * bn[2] - de[2] * bn[1], de[2] * bn[0] - de[0] * bn[2], de[0] * bn[1] - de[1] * bn[0]};
    const double KCe[3] = {de[1] * Ce[2] - de[2] * Ce[1], de[2] * Ce[0] - de[0] * Ce[2], de[0] * Ce[1] - de[1] * Ce[0]};
    double g[6];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        g[r] = S[21 + r] - Be[r] - Kbn[r] + KCe[r];
        g[3 + r] = bn[r] - Ce[r];
    }
    // LDL^T with relative pivot guard: d_j <= 1e-6 * H_jj drops component j.  A dropped pivot
    // keeps D = 0 and a zero column in L, so the unguarded sums below simply skip it.
    double L[36], D[6], iD[6];
    bool dead[6];
#pragma unroll
    for (int i = 0; i < 36; i++) L[i] = 0.0;
    int degenerate = 0;
    double minratio = 1.0;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = H[6 * j + j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[6 * j + k] * L[6 * j + k] * D[k];
        const double hjj = H[6 * j + j];
        const double ratio = hjj > 0 ? d / hjj : 0.0;
        minratio = ratio < minratio ? ratio : minratio;
        dead[j] = !(d > 1e-6 * hjj) || !(hjj > 0);
        degenerate |= dead[j] ? 1 : 0;
        D[j] = dead[j] ? 0.0 : d;
        iD[j] = dead[j] ? 0.0 : 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double vv = H[6 * i + j];
#pragma unroll
            for (int k = 0; k < j; k++) vv -= L[6 * i + k] * L[6 * j + k] * D[k];
            L[6 * i + j] = vv * iD[j];
        }
    }
    ratio_out = minratio;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double vv = -g[i];
#pragma unroll
        for (int k = 0; k < i; k++) vv -= L[6 * i + k] * y[k];
        y[i] = dead[i] ? 0.0 : vv;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] = y[i] * iD[i];
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double vv = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) vv -= L[6 * k + i] * x[k];
        x[i] = dead[i] ? 0.0 : vv;
    }
    // Tinc = Tr(qbar) R Tr(t cos th) R Tr(-pbar):  x -> R(R(x - pbar) + t cos th) + qbar
    const double an = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    const double th = atan(an);
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double sn, ct;
    sincos(th, &sn, &ct);
    if (an > 1e-300) {
        const double ia = 1.0 / an;
        const double u[3] = {x[0] * ia, x[1] * ia, x[2] * ia};
        const double Cc = 1.0 - ct;
        R[0] = ct + u[0] * u[0] * Cc;        R[1] = u[0] * u[1] * Cc - u[2] * sn; R[2] = u[0] * u[2] * Cc + u[1] * sn;
        R[3] = u[1] * u[0] * Cc + u[2] * sn; R[4] = ct + u[1] * u[1] * Cc;        R[5] = u[1] * u[2] * Cc - u[0] * sn;
        R[6] = u[2] * u[0] * Cc - u[1] * sn; R[7] = u[2] * u[1] * Cc + u[0] * sn; R[8] = ct + u[2] * u[2] * Cc;
    }
    double pbar[3], qbar[3], tc[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { pbar[k] = origin[k] + dp[k]; qbar[k] = origin[k] + dq[k]; tc[k] = x[3 + k] * ct; }
    double R2[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) R2[3 * r + c] = R[3 * r] * R[c] + R[3 * r + 1] * R[3 + c] + R[3 * r + 2] * R[6 + c];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        Tinc[4 * r] = R2[3 * r]; Tinc[4 * r + 1] = R2[3 * r + 1]; Tinc[4 * r + 2] = R2[3 * r + 2];
        Tinc[4 * r + 3] = qbar[r] + (R[3 * r] * tc[0] + R[3 * r + 1] * tc[1] + R[3 * r + 2] * tc[2]) -
                          (R2[3 * r] * pbar[0] + R2[3 * r + 1] * pbar[1] + R2[3 * r + 2] * pbar[2]);
    }
    return degenerate;
}

// Device-side loop control: the registration loop is a CUDA-graph WHILE node whose body is
// [K1, K2].  Every K2 block reports whether its pair is finished; the last block to arrive keeps
// the loop running while any pair is still iterating (myicp.cpp:123 evaluated on the device).
__device__ __forceinline__ void loop_ctrl_arrive(const SolveParams& P, bool pair_done) {
    if (!P.use_cond) return;
    if (pair_done) atomicAdd(&P.ctrl->done_count, 1u);
    __threadfence();
    const unsigned int t = atomicAdd(&P.ctrl->ticket, 1u);
    const unsigned int pairs = gridDim.x / cooperative_groups::this_cluster().num_blocks();   // one arrival per pair
    if (t == pairs - 1) {
        const unsigned int done = atomicAdd(&P.ctrl->done_count, 0u);
        P.ctrl->ticket = 0u;
        P.ctrl->done_count = 0u;
        cudaGraphSetConditional(P.cond, done < pairs ? 1u : 0u);
    }
}

// Canonical reduction (common.cuh): the rows of a part are added in a fixed order (27 interleaved
// chains, then the chains in order), then the 8 part sums in part order, starting from +0.0.  A single
// large pair is reduced by a thread-block CLUSTER of 8 blocks, one part each (one block cannot pull
// rows out of L2 faster than ~100 GB/s); block 0 collects the part sums through distributed shared
// memory.  Batches of small pairs launch plain blocks that take the parts in turn.  In the
// query-sharded modes a rank holds a run of parts and the part sums of the others arrive over NVLink
// (peer stores, fused here) or through NCCL (between a reduce-only and a solve-only launch); the final
// sum over the 8 parts is the same expression on every rank and on a single GPU.
__global__ void __launch_bounds__(kSolveThreads) solve_kernel(const __grid_constant__ SolveParams P) {
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const uint32_t cn = cluster.num_blocks(), cr = cluster.block_rank();
    const uint32_t seg = blockIdx.x / cn;
    SegState* st = P.state + seg;
    __shared__ double sPart[kSolveParts][ICP_NSUMS];
    __shared__ double sOct[ICP_NPARTS][ICP_NSUMS];   // part sums; indexed by LOCAL part until the exchange
    __shared__ double sS[ICP_NSUMS];
    const int tid = threadIdx.x;
    ICP_K2_STAMP(0);
    if (blockIdx.x == 0 && tid == 0) P.ctrl->work = P.k1_grid;   // K1 of this iteration is complete: rewind its ticket counter
    if (st->searches == 0) ICP_K2_STAMP(5);
    if (st->done) {   // the same for every block of the cluster: all leave, nobody waits
        if (tid == 0 && cr == 0) loop_ctrl_arrive(P, true);
        return;
    }
    if (P.mode == 2 && cr != 0) return;   // solve-only launch: nothing to add up
    if (P.mode != 2) {
        const uint32_t tiles = icp_tiles(P.src_counts[seg]);
        const int v = tid % ICP_NSUMS, chain = tid / ICP_NSUMS;
        for (uint32_t ol = cr; ol < P.nparts; ol += cn) {
            if (chain < kSolveParts) {
                // chain c adds the part's tile rows c, c + 27, c + 54, ... in that order; ICP_K2_BATCH loads are
                // issued back to back so the (L2-latency) loads overlap instead of serialising with the adds
                double s = 0.0;
                const uint32_t part = P.part0 + ol;
                const uint32_t own_blocks = icp_part_tiles(tiles, part);   // rows of this part: its tiles in curve order
                ICP_ASSERT(part < ICP_NPARTS && (own_blocks == 0 || icp_part_tile(tiles, part, own_blocks - 1) < P.tiles_max));
                const uint32_t C = icp_part_chunk(tiles), lc = 31u - (uint32_t)__clz((int)C);   // C is a power of two
                const double* row = P.partial + (size_t)seg * P.tiles_max * ICP_PART_STRIDE + v;
                for (uint32_t b = chain; b < own_blocks; b += ICP_K2_BATCH * kSolveParts) {
                    double x[ICP_K2_BATCH];   // rows past the end count as 0.0, which leaves the sum unchanged
#pragma unroll
                    for (int u = 0; u < ICP_K2_BATCH; u++)
                        x[u] = (b + u * kSolveParts < own_blocks)
                                   ? row[(size_t)(((((b + u * kSolveParts) >> lc) * ICP_NPARTS + part) << lc) +
                                                  ((b + u * kSolveParts) & (C - 1u))) * ICP_PART_STRIDE] : 0.0;   // = icp_part_tile
#pragma unroll
                    for (int u = 0; u < ICP_K2_BATCH; u++) s += x[u];
                }
                sPart[chain][v] = s;
            }
            __syncthreads();
            if (tid < ICP_NSUMS) {
                double s = 0.0;
                for (int p = 0; p < kSolveParts; p++) s += sPart[p][tid];
                sOct[ol][tid] = s;
            }
            __syncthreads();
        }
        ICP_K2_STAMP(1);
        if (cn > 1) {
            cluster.sync();                         // block ol holds the sum of local part ol
            if (cr == 0 && tid < ICP_NSUMS)
                for (uint32_t ol = 1; ol < P.nparts; ol++)
                    sOct[ol][tid] = cluster.map_shared_rank(&sOct[0][0], ol % cn)[ol * ICP_NSUMS + tid];
            cluster.sync();                         // block 0 has read: the others may leave
            if (cr != 0) return;
            __syncthreads();
        }
        if (P.mode == 1) {
            // NCCL mode: my parts' sums, zeros for the parts of the other ranks -- the all-reduce then adds
            // exact zeros to every value, so its result does not depend on NCCL's reduction order
            for (int t = tid; t < ICP_NPARTS * ICP_PART_STRIDE; t += kSolveThreads) {
                const uint32_t o = t / ICP_PART_STRIDE, j = t % ICP_PART_STRIDE;
                const bool mine = o >= P.part0 && o < P.part0 + P.nparts && j < ICP_NSUMS;
                P.gparts[(size_t)seg * ICP_NPARTS * ICP_PART_STRIDE + t] = mine ? sOct[o - P.part0][j] : 0.0;
            }
            return;
        }
        if (P.peer_world > 1) {
            // ---- fused exchange over NVLink: store my part sums into every rank's buffer, publish, wait
            // for the same exchange of every peer, then read all 8 parts from my own buffer ----
            const unsigned long long seq = P.ctrl->seq + 1ull;
            const int par = (int)(seq & 1ull);
            if (tid < (int)(P.nparts * ICP_PART_STRIDE)) {
                const uint32_t ol = tid / ICP_PART_STRIDE, j = tid % ICP_PART_STRIDE;
                const double val = j < ICP_NSUMS ? sOct[ol][j] : 0.0;
                for (int r = 0; r < P.peer_world; r++) P.peers[r]->sums[par][P.part0 + ol][j] = val;
                __threadfence_system();
            }
            __syncthreads();
            if (tid < P.peer_world) {
                volatile unsigned long long* flag = &P.peers[tid]->flag[par][P.peer_rank].seq;
                *flag = seq;
                __threadfence_system();
                // wait until rank `tid` has published this exchange into MY buffer
                volatile unsigned long long* mine = &P.peers[P.peer_rank]->flag[par][tid].seq;
                unsigned long long t0, t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                while (*mine != seq) {   // bounded: a rank that never publishes must not hang the GPU
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > P.peer_timeout_ns) { P.ctrl->comm_error = 1; break; }
                }
                __threadfence_system();
            }
            __syncthreads();
            if (tid < ICP_NSUMS) {
                const volatile double* all = &P.peers[P.peer_rank]->sums[par][0][0];
                double s = 0.0;
                for (int o = 0; o < ICP_NPARTS; o++) s += all[o * ICP_PART_STRIDE + tid];
                sS[tid] = s;
            }
            __syncthreads();
            if (tid == 0) P.ctrl->seq = seq;
        } else {
            if (tid < ICP_NSUMS) {
                double s = 0.0;
                for (uint32_t ol = 0; ol < P.nparts; ol++) s += sOct[ol][tid];
                sS[tid] = s;
            }
            __syncthreads();
        }
        if (tid < ICP_NSUMS && P.mode == 3) P.gsum[(size_t)seg * ICP_PART_STRIDE + tid] = sS[tid];
    } else {
        if (tid < ICP_NSUMS) {
            const double* all = P.gparts + (size_t)seg * ICP_NPARTS * ICP_PART_STRIDE;
            double s = 0.0;
            for (int o = 0; o < ICP_NPARTS; o++) s += all[o * ICP_PART_STRIDE + tid];
            sS[tid] = s;
        }
        __syncthreads();
    }
    if (tid != 0) return;
    ICP_K2_STAMP(2);

    const LoopParams lp = *P.lp;
    const double diff = sS[33];
    const int k = st->searches;
    st->searches = k + 1;
    st->diff = diff;
    st->sumd2 = sS[34];
    if (lp.trace && P.trace_T && k < P.trace_stride) {
        float* tt = P.trace_T + ((size_t)seg * P.trace_stride + k) * 12;
        for (int i = 0; i < 12; i++) tt[i] = st->Tf[i];
        P.trace_diff[(size_t)seg * P.trace_stride + k] = diff;
    }
    int stop = 0;
    if (P.peer_world > 1 && P.ctrl->comm_error) stop = 5;
    else if (P.mode == 3) stop = 4;
    else if (!(diff > (double)lp.diff_threshold)) stop = 1;                                    // myicp.cpp:123
    else if (st->iter >= lp.max_iters) stop = 3;                                               // iters++ < max_iters
    else if (lp.rel_diff_tol > 0.0 && k > 0 && fabs(st->prev_diff - diff) <= lp.rel_diff_tol * st->prev_diff) stop = 2;
    if (stop) {
        st->done = 1;
        st->converged = (stop == 1) ? 1 : (stop == 2 ? 2 : 0);
        loop_ctrl_arrive(P, true);
        return;
    }
    st->prev_diff = diff;
    double Tinc[12], x[6], ratio;
    const double origin[3] = {st->origin[0], st->origin[1], st->origin[2]};
    const int deg = solve_compose_dev(sS, origin, Tinc, x, ratio);
    if (deg) st->degenerate = 1;
    if (ratio < st->min_ratio) st->min_ratio = ratio;
    double Told[12], Tn[12];
#pragma unroll
    for (int i = 0; i < 12; i++) Told[i] = st->T[i];
    mat34_mul(Tinc, Told, Tn);                                                                 // myicp.cpp:138
#pragma unroll
    for (int i = 0; i < 12; i++) { st->T[i] = Tn[i]; st->Tf[i] = (float)Tn[i]; }
    if (P.hist && (uint32_t)(k + 1) < P.hist_stride) {   // the transform search k+1 will use
        float* h = P.hist + ((size_t)seg * P.hist_stride + (uint32_t)(k + 1)) * 12;
#pragma unroll
        for (int i = 0; i < 12; i++) h[i] = (float)Tn[i];
    }
    {
        // how far can this step move a source point?  |step at the box centre| + rotation angle x box radius.  When it
        // is within a few point spacings the NEXT step will be far smaller (the loop converges geometrically), so the
        // margins the next search records will let most of the searches after it skip their walks (K1 skip test);
        // before that, recording margins is wasted work (+15-20 % per walk), so the early searches do without.
        const double cx = origin[0], cy = origin[1], cz = origin[2];
        const double mx = Tinc[0] * cx + Tinc[1] * cy + Tinc[2] * cz + Tinc[3] - cx;
        const double my = Tinc[4] * cx + Tinc[5] * cy + Tinc[6] * cz + Tinc[7] - cy;
        const double mz = Tinc[8] * cx + Tinc[9] * cy + Tinc[10] * cz + Tinc[11] - cz;
        const double an = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        const double step = sqrt(mx * mx + my * my + mz * mz) + 2.0 * atan(an) * st->radius;
        st->have_margins = st->track;   // what the search that has just run recorded
        st->track = step < ICP_TRACK_STEP_SPACINGS * st->spacing ? 1 : 0;
        st->probe = (lp.probe && step >= ICP_PROBE_STEP_SPACINGS * st->spacing) ? 1 : 0;   // after a long step the old pairs are no bounds
    }
    st->iter = st->iter + 1;
    st->pairs = (unsigned long long)(sS[36] + 0.5);
    ICP_K2_STAMP(3);
    loop_ctrl_arrive(P, false);
    ICP_K2_STAMP(4);
}

__global__ void solve_compose_kernel(const double* in /*37 sums + 3 origin*/, double* out /*12 T + 6 x + ratio + deg*/) {
    double Tinc[12], x[6], ratio;
    const double origin[3] = {in[ICP_NSUMS], in[ICP_NSUMS + 1], in[ICP_NSUMS + 2]};
    int deg = solve_compose_dev(in, origin, Tinc, x, ratio);
    for (int i = 0; i < 12; i++) out[i] = Tinc[i];
    for (int i = 0; i < 6; i++) out[12 + i] = x[i];
    out[18] = ratio;
    out[19] = (double)deg;
}

// ------------------------------------------------------------------------------------------------
// exports
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) export_corr_kernel(CloudView src, CloudView tgt, const SegState* __restrict__ state,
                                                          const uint32_t* __restrict__ corr, uint32_t src_max_count,
                                                          int32_t* __restrict__ idx_out, float* __restrict__ d2_out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pitch = src.shape.pitch;
    if (j >= (size_t)src.nseg * pitch) return;
    uint32_t seg = (uint32_t)(j / pitch), i = (uint32_t)(j - (size_t)seg * pitch);
    if (i >= src.counts[seg]) return;
    float4 s = src.pts[j];
    uint32_t pos = corr[j];
    size_t o = (size_t)seg * src_max_count + __float_as_uint(s.w);
    if (pos < tgt.counts[seg]) {
        float4 t = tgt.pts[(size_t)seg * tgt.shape.pitch + pos];
        float3 p = apply_rn(state[seg].Tf, s.x, s.y, s.z);
        idx_out[o] = (int32_t)__float_as_uint(t.w);
        d2_out[o] = dist2_rn(p.x, p.y, p.z, t.x, t.y, t.z);
    } else {
        idx_out[o] = -1;
        d2_out[o] = __int_as_float(0x7f800000);
    }
}

struct T12 { float m[12]; };

__global__ void __launch_bounds__(256) export_transformed_kernel(CloudView src, T12 T, uint32_t src_max_count,
                                                                 float* __restrict__ out, uint32_t out_stride_floats) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pitch = src.shape.pitch;
    if (j >= (size_t)src.nseg * pitch) return;
    uint32_t seg = (uint32_t)(j / pitch), i = (uint32_t)(j - (size_t)seg * pitch);
    if (i >= src.counts[seg]) return;
    float4 s = src.pts[j];
    float3 p = apply_rn(T.m, s.x, s.y, s.z);
    float* o = out + ((size_t)seg * src_max_count + __float_as_uint(s.w)) * out_stride_floats;
    o[0] = p.x; o[1] = p.y; o[2] = p.z;
}

bool stream_capturable(cudaStream_t s) {
    return s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread;
}

}  // namespace

static int ensure_io(icpsymm_ctx* ctx, size_t bytes) {
    char* p = (char*)ctx->io_buf;
    int rc = icp_reserve(ctx, &p, &ctx->io_bytes, bytes);
    ctx->io_buf = p;
    return rc;
}

// one block per pair, or a cluster of 8 blocks per pair when there are many rows and few pairs
static void launch_solve(const SolveParams& sp, uint32_t nseg, cudaStream_t st) {
    const unsigned cn = (sp.tiles_max >= 512u && nseg <= 16u) ? 8u : 1u;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nseg * cn); cfg.blockDim = dim3(kSolveThreads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cn; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, solve_kernel, sp);
}

static void launch_iteration(icpsymm_ctx* ctx, const NNParams& np, SolveParams sp, uint32_t nseg, bool use_comm,
                             cudaEvent_t* ev /*4 events or null*/) {
    cudaStream_t st = ctx->stream;
    if (ev) cudaEventRecord(ev[0], st);
    if (np.tile_seed) probe_kernel<<<(np.total_work * 32u + kNNThreads - 1u) / kNNThreads, kNNThreads, 0, st>>>(np);
    nn_accum_kernel<<<np.grid, kNNThreads, 0, st>>>(np);
    if (ev) { cudaEventRecord(ev[1], st); cudaEventRecord(ev[2], st); }
    if (!use_comm) {
        launch_solve(sp, nseg, st);
    } else {
        sp.mode = 1;
        launch_solve(sp, nseg, st);
        icp_comm_allreduce(ctx, sp.gparts, (size_t)nseg * ICP_NPARTS * ICP_PART_STRIDE);
        sp.mode = 2;
        launch_solve(sp, nseg, st);
    }
    if (ev) cudaEventRecord(ev[3], st);
}

int icp_run_loop(icpsymm_ctx* ctx, int max_iters, float thr, const float* T_init, const double* origin,
                 bool single_search, float* T_out, icpsymm_stats* stats) {
    int rc = icp_loop_enqueue(ctx, max_iters, thr, T_init, origin, single_search);
    if (rc < 0) return rc;
    return icp_loop_finish(ctx, T_out, stats);
}

// first half of a registration: everything is ENQUEUED on the context stream, nothing waits for the device
int icp_loop_enqueue(icpsymm_ctx* ctx, int max_iters, float thr, const float* T_init, const double* origin,
                     bool single_search) {
    if (ctx->pending.active) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "a registration is already in flight: call icpsymm_wait first");
    CloudSet& S = ctx->src;
    CloudSet& T = ctx->tgt;
    { int jr = icp_join_clouds(ctx, true); if (jr < 0) return jr; }   // asynchronous setup: join + deferred checks
    if (!S.valid || !T.valid) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "set_source / set_target first");
    if (S.nseg != T.nseg) return icp_fail_msg(ctx, ICPSYMM_ERR_SIZE, "source and target hold a different number of pairs");
    if (max_iters < 0) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "max_iters < 0");
    if (ctx->opt_identity)
        for (uint32_t s = 0; s < S.nseg; s++)
            if (S.h_counts[s] != T.h_counts[s])
                return icp_fail_msg(ctx, ICPSYMM_ERR_SIZE, "identity pairing needs equal cloud sizes (func.cpp:21)");
    if (!single_search && (!S.has_normals || !T.has_normals))
        return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "normals missing: call icpsymm_estimate_normals or icpsymm_set_normals");
    const uint32_t nseg = S.nseg;
    cudaStream_t st = ctx->stream;
    int rc;

    // canonical grid (common.cuh): rows per part from the largest pair; a query shard is a run of whole parts
    if (ctx->shard_world > ICP_NPARTS)
        return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "query sharding supports at most 8 ranks (one NVSwitch domain)");
    const uint32_t bps = icp_tiles(S.max_count);   // rows per pair: one per tile
    const uint32_t part0 = (uint32_t)(ICP_NPARTS * ctx->shard_rank / ctx->shard_world);
    const uint32_t nparts = (uint32_t)(ICP_NPARTS * (ctx->shard_rank + 1) / ctx->shard_world) - part0;

    if ((rc = icp_reserve(ctx, &ctx->d_state, &ctx->state_cap, nseg))) return rc;
    if (ctx->h_state_cap < nseg) {
        if (ctx->h_state) ICP_CUDA_CHECK(cudaFreeHost(ctx->h_state));
        ctx->h_state = nullptr;
        ICP_CUDA_CHECK(cudaMallocHost((void**)&ctx->h_state, sizeof(SegState) * nseg));
        ctx->h_state_cap = nseg;
    }
    if ((rc = icp_reserve(ctx, &ctx->partial, &ctx->partial_cap, (size_t)nseg * bps * ICP_PART_STRIDE))) return rc;
    if ((rc = icp_reserve(ctx, &ctx->gsum, &ctx->gsum_cap, (size_t)nseg * ICP_PART_STRIDE))) return rc;
    if ((rc = icp_reserve(ctx, &ctx->gparts, &ctx->gparts_cap, (size_t)nseg * ICP_NPARTS * ICP_PART_STRIDE))) return rc;
    if ((rc = icp_reserve(ctx, &ctx->corr_pos, &ctx->corr_cap, (size_t)nseg * S.shape.pitch))) return rc;
    // searches whose transform is kept per pair; a query's reference search is stored in 8 bits (K1 records margins only
    // while the search index is below 255), so more history than that is never read
    const uint32_t hist_stride = (uint32_t)(single_search ? 1 : (max_iters + 2 < 256 ? max_iters + 2 : 256));
    if ((rc = icp_reserve(ctx, &ctx->margin2, &ctx->margin_cap, (size_t)nseg * S.shape.pitch))) return rc;
    if ((rc = icp_reserve(ctx, &ctx->tile_seed, &ctx->tile_seed_cap, (size_t)nseg * bps))) return rc;
    if ((rc = icp_reserve(ctx, &ctx->ref_search, &ctx->ref_cap, (size_t)nseg * S.shape.pitch))) return rc;
    if ((rc = icp_reserve(ctx, &ctx->hist, &ctx->hist_cap, (size_t)nseg * hist_stride * 12))) return rc;
    if (!ctx->d_lp) ICP_CUDA_CHECK(cudaMalloc((void**)&ctx->d_lp, sizeof(LoopParams)));
    const int n_search = single_search ? 1 : max_iters + 1;
    if (ctx->opt_trace) {
        size_t need = (size_t)nseg * n_search;
        if ((rc = icp_reserve(ctx, &ctx->trace_T, &ctx->trace_T_cap, need * 12))) return rc;
        if ((rc = icp_reserve(ctx, &ctx->trace_diff, &ctx->trace_diff_cap, need))) return rc;
        ctx->trace_stride = n_search;
    }

    if (!ctx->ev_begin) {
        ICP_CUDA_CHECK(cudaEventCreate(&ctx->ev_begin));
        ICP_CUDA_CHECK(cudaEventCreate(&ctx->ev_end));
    }
    ICP_CUDA_CHECK(cudaEventRecord(ctx->ev_begin, st));   // loop_ms covers everything from here: state set-up, tile order, loop

    LoopParams lp;
    lp.max_iters = max_iters;
    lp.diff_threshold = thr;
    lp.rel_diff_tol = ctx->opt_rel_diff_tol;
    lp.max_pair_d2 = ctx->opt_max_pair_dist > 0.f ? ctx->opt_max_pair_dist * ctx->opt_max_pair_dist : INFINITY;
    lp.min_normal_dot = ctx->opt_min_normal_dot;
    lp.identity_pairs = ctx->opt_identity;
    lp.warm_start = ctx->opt_warm_start;
    lp.trace = ctx->opt_trace;
    lp.skip_walk = ctx->opt_skip_walk;
    lp.probe = (ctx->opt_probe && !ctx->opt_identity) ? 1 : 0;
    set_lp_kernel<<<1, 1, 0, st>>>(ctx->d_lp, lp);

    InitArgs ia;
    memset(&ia, 0, sizeof ia);
    if (T_init) {
        ia.has_T = 1;
        for (int i = 0; i < 12; i++) ia.T[i] = T_init[i];
    }
    ia.probe = (ctx->opt_probe && !ctx->opt_identity) ? 1 : 0;
    if (origin) {
        ia.has_origin = 1;
        for (int i = 0; i < 3; i++) ia.origin[i] = origin[i];
    }
    if (!ctx->d_ctrl) {
        ICP_CUDA_CHECK(cudaMalloc((void**)&ctx->d_ctrl, sizeof(LoopCtrl)));
        ICP_CUDA_CHECK(cudaMemsetAsync(ctx->d_ctrl, 0, sizeof(LoopCtrl), st));
    }
    uint32_t k1_grid;
    {
        const uint32_t wps = icp_shard_work(bps, nparts);
        const uint32_t resident = (uint32_t)ctx->sm_count * (uint32_t)kNNBlocksPerSM;
        k1_grid = nseg * wps < resident ? nseg * wps : resident;
    }
    init_state_kernel<<<(nseg + 127) / 128, 128, 0, st>>>(ctx->d_state, nseg, ia, T.bb_lo, T.bb_hi, T.counts, (LoopCtrl*)ctx->d_ctrl, k1_grid, ctx->hist,
                                                           hist_stride);
    // no unit has walked yet in this registration: nothing can be skipped in the first search
    ICP_CUDA_CHECK(cudaMemsetAsync(ctx->ref_search, 0xff, (size_t)nseg * S.shape.pitch, st));
    // ICPSYMM_OPT_WARM_START = 2 keeps the correspondences of the previous call as seeds of the first search
    // (frame-to-frame tracking); a seed is only ever a candidate, so the result stays exact
    if (!(ctx->opt_warm_start == 2 && ctx->have_corr))
        ICP_CUDA_CHECK(cudaMemsetAsync(ctx->corr_pos, 0xff, sizeof(uint32_t) * (size_t)nseg * S.shape.pitch, st));

    NNParams np;
    np.src = S.view(); np.tgt = T.view();
    np.state = ctx->d_state; np.lp = ctx->d_lp; np.corr_pos = ctx->corr_pos; np.partial = ctx->partial;
    np.tiles_max = bps; np.part0 = part0; np.nparts = nparts;
    np.work_per_seg = icp_shard_work(bps, nparts);
    np.total_work = nseg * np.work_per_seg;
    {
        const uint32_t resident = (uint32_t)ctx->sm_count * (uint32_t)kNNBlocksPerSM;   // persistent blocks
        np.grid = np.total_work < resident ? np.total_work : resident;
    }
    SolveParams sp;
    sp.state = ctx->d_state; sp.lp = ctx->d_lp; sp.partial = ctx->partial; sp.gsum = ctx->gsum; sp.gparts = ctx->gparts;
    sp.hist = ctx->hist; sp.hist_stride = hist_stride;
    sp.src_counts = S.counts; sp.tiles_max = bps; sp.part0 = part0; sp.nparts = nparts; sp.k1_grid = np.grid;
    sp.trace_T = ctx->opt_trace ? ctx->trace_T : nullptr;
    sp.trace_diff = ctx->opt_trace ? ctx->trace_diff : nullptr;
    sp.trace_stride = ctx->trace_stride; sp.mode = single_search ? 3 : 0;
    np.ctrl = (LoopCtrl*)ctx->d_ctrl; np.order = nullptr;
    np.tile_seed = (ctx->opt_probe && !ctx->opt_identity) ? ctx->tile_seed : nullptr;
    np.margin2 = ctx->margin2; np.ref_search = ctx->ref_search; np.hist = ctx->hist; np.hist_stride = hist_stride;
    // most-expensive-first order of the first search (single pairs that fill the machine more than twice)
    if (nseg == 1 && !ctx->opt_identity && ctx->opt_tile_order && np.total_work >= 2u * np.grid && T.shape.n_levels >= 4) {
        const uint32_t work = np.total_work;
        int rec_level = (int)T.shape.n_levels - 1;   // records of level l hold the boxes of level l-1: the finest level
        for (uint32_t l = 1; l + 1 <= T.shape.n_levels - 1; l++)   // whose boxes fit the kernel's shared memory
            if (T.shape.level_cnt[l] * ICP_FAN <= (uint32_t)kCostBoxesMax) { rec_level = (int)l; break; }
        size_t cap = ctx->order_cap;
        if ((rc = icp_reserve(ctx, &ctx->order_buf, &cap, (size_t)4 * work))) return rc;
        ctx->order_cap = cap;
        uint32_t *ka = ctx->order_buf, *kb = ka + work, *va = kb + work, *vb = va + work;
        size_t temp_bytes = 0;
        ICP_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, ka, kb, va, vb, (int)work, 0, 16, st));
        {
            char* tmp = (char*)ctx->order_temp;
            if ((rc = icp_reserve(ctx, &tmp, &ctx->order_temp_bytes, temp_bytes))) return rc;
            ctx->order_temp = tmp;
        }
        {
            // one warp per 2 tiles (16 for the largest pairs): enough blocks to fill the machine, few enough to amortise the
            // staging of the boxes (1M pair: 45 us with 4 tiles per warp, 32 us with 2)
            const uint32_t tpw = work >= 16384u ? 16u : 2u, warps = (work + tpw - 1u) / tpw;
            const size_t smem = (size_t)kCostBoxesMax * 6 * sizeof(float);
            if (!ctx->cost_attr_set) {   // per device
                ICP_CUDA_CHECK(cudaFuncSetAttribute(tile_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ctx->cost_attr_set = true;
            }
            tile_cost_kernel<<<(warps * 32u + 255u) / 256u, 256, smem, st>>>(S.view(), T.view(), ctx->d_state, part0, nparts,
                                                                             work, rec_level, tpw, ka, va);
        }
        ICP_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ctx->order_temp, temp_bytes, ka, kb, va, vb, (int)work, 0, 16, st));
        np.order = vb;
    }
    sp.ctrl = (LoopCtrl*)ctx->d_ctrl; sp.use_cond = 0; sp.cond = 0;
    // single searches (find_correspondences, eval_diff, normal_equations) are LOCAL to the rank's shard: only
    // the registration loop is collective, so one rank may call them without stalling on its peers
    sp.peer_world = single_search ? 0 : ctx->peer_world; sp.peer_rank = ctx->shard_rank;
    sp.peer_timeout_ns = (unsigned long long)(ctx->opt_peer_timeout_ms * 1e6);
    for (int r = 0; r < ICP_MAX_PEERS; r++) sp.peers[r] = (PeerBuf*)ctx->peer_ptrs[r];
    if (ctx->peer_world > 1 && nseg != 1)
        return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "the NVLink peer exchange supports single pairs only");

    const bool use_comm = ctx->nccl_comm != nullptr && ctx->shard_world > 1;
    const bool timing = ctx->opt_kernel_timing != 0;
    const bool use_graph = ctx->opt_use_graph && !timing && !single_search && !use_comm && stream_capturable(st);

    if (timing && ctx->ev_pool_n < 4 * n_search) {
        for (int i = 0; i < ctx->ev_pool_n; i++) cudaEventDestroy(ctx->ev_pool[i]);
        free(ctx->ev_pool);
        ctx->ev_pool_n = 4 * n_search;
        ctx->ev_pool = (cudaEvent_t*)malloc(sizeof(cudaEvent_t) * ctx->ev_pool_n);
        for (int i = 0; i < ctx->ev_pool_n; i++) ICP_CUDA_CHECK(cudaEventCreate(&ctx->ev_pool[i]));
    }

    if (use_graph) {
        GraphKey key;
        key.max_iters = max_iters; key.nseg = nseg; key.tiles_max = bps;   // the parts follow from the shard
        key.src_pitch = S.shape.pitch; key.tgt_pitch = T.shape.pitch;
        key.generation = ctx->generation;
        key.shard_rank = ctx->shard_rank; key.shard_world = ctx->shard_world; key.trace = ctx->opt_trace;
        key.trace_stride = ctx->opt_trace ? ctx->trace_stride : 0;
        key.mode = ctx->opt_use_graph + 16 * ctx->peer_world + (np.order ? 1024 : 0) + (np.tile_seed ? 2048 : 0);
        if (!ctx->graph_exec || !(ctx->graph_key == key)) {
            if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
            cudaGraph_t g = nullptr;
            if (ctx->opt_use_graph == 2) {
                // unrolled graph: (max_iters+1) x [K1,K2], finished pairs turn launches into early exits
                ICP_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                for (int k = 0; k < n_search; k++) launch_iteration(ctx, np, sp, nseg, false, nullptr);
                ICP_CUDA_CHECK(cudaStreamEndCapture(st, &g));
            } else {
                // WHILE node: body [K1,K2] repeats until K2 clears the condition on the device
                ICP_CUDA_CHECK(cudaGraphCreate(&g, 0));
                cudaGraphConditionalHandle handle;
                ICP_CUDA_CHECK(cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault));
                cudaGraphNodeParams cp = {};
                cp.type = cudaGraphNodeTypeConditional;
                cp.conditional.handle = handle;
                cp.conditional.type = cudaGraphCondTypeWhile;
                cp.conditional.size = 1;
                cudaGraphNode_t wnode;
                ICP_CUDA_CHECK(cudaGraphAddNode(&wnode, g, nullptr, 0, &cp));
                cudaGraph_t body = cp.conditional.phGraph_out[0];
                SolveParams spc = sp;
                spc.use_cond = 1; spc.cond = handle;
                ICP_CUDA_CHECK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
                launch_iteration(ctx, np, spc, nseg, false, nullptr);
                ICP_CUDA_CHECK(cudaStreamEndCapture(st, nullptr));
            }
            ICP_CUDA_CHECK(cudaGraphInstantiate(&ctx->graph_exec, g, 0));
            cudaGraphDestroy(g);
            ctx->graph_key = key;
        }
        ICP_CUDA_CHECK(cudaGraphLaunch(ctx->graph_exec, st));
        ICP_CUDA_CHECK(cudaEventRecord(ctx->ev_end, st));
    } else {
        for (int k = 0; k < n_search; k++)
            launch_iteration(ctx, np, sp, nseg, use_comm, timing ? ctx->ev_pool + 4 * k : nullptr);
        ICP_CUDA_CHECK(cudaEventRecord(ctx->ev_end, st));
    }
    ICP_CUDA_CHECK(cudaGetLastError());
    ICP_CUDA_CHECK(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(SegState) * nseg, cudaMemcpyDeviceToHost, st));
    if (!ctx->done_ev) ICP_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->done_ev, cudaEventDisableTiming));
    ICP_CUDA_CHECK(cudaEventRecord(ctx->done_ev, st));
    ctx->have_corr = true;
    // everything the second half needs to turn the pinned state into the caller's outputs
    ctx->pending.active = true; ctx->pending.nseg = nseg; ctx->pending.n_search = n_search; ctx->pending.timing = timing;
    ctx->pending.use_comm = use_comm; ctx->pending.single_search = single_search; ctx->pending.part0 = part0;
    ctx->pending.nparts = nparts;
    return ICPSYMM_OK;
}

// second half of a registration: wait for the loop enqueued by icp_loop_enqueue and hand out its result
int icp_loop_finish(icpsymm_ctx* ctx, float* T_out, icpsymm_stats* stats) {
    if (!ctx->pending.active) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "no registration in flight");
    ctx->pending.active = false;
    CloudSet& S = ctx->src;
    cudaStream_t st = ctx->stream;
    const uint32_t nseg = ctx->pending.nseg, part0 = ctx->pending.part0, nparts = ctx->pending.nparts;
    const int n_search = ctx->pending.n_search;
    const bool timing = ctx->pending.timing, use_comm = ctx->pending.use_comm, single_search = ctx->pending.single_search;
    ICP_CUDA_CHECK(cudaEventSynchronize(ctx->done_ev));
    if (ctx->peer_world > 1) {
        LoopCtrl hc;
        ICP_CUDA_CHECK(cudaMemcpy(&hc, ctx->d_ctrl, sizeof hc, cudaMemcpyDeviceToHost));
        if (hc.comm_error) {
            // report once and clear, so that the context is usable again after the peers have re-attached
            ICP_CUDA_CHECK(cudaMemsetAsync(&((LoopCtrl*)ctx->d_ctrl)->comm_error, 0, sizeof(int), st));
            return icp_fail_msg(ctx, ICPSYMM_ERR_COMM, "peer exchange timed out: a rank did not publish its sums "
                                                      "(all ranks must detach and attach again before the next sharded registration)");
        }
    }

    float loop_ms = 0.f;
    cudaEventElapsedTime(&loop_ms, ctx->ev_begin, ctx->ev_end);
    int max_searches = 0;
    for (uint32_t s = 0; s < nseg; s++) max_searches = ctx->h_state[s].searches > max_searches ? ctx->h_state[s].searches : max_searches;
    float search_ms = 0.f, solve_ms = 0.f;
    int search_launches = 0, solve_launches = 0;
    ctx->timed_searches = timing ? (max_searches < n_search ? max_searches : n_search) : 0;
    if (timing) {
        for (int k = 0; k < max_searches && k < n_search; k++) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, ctx->ev_pool[4 * k], ctx->ev_pool[4 * k + 1]);
            cudaEventElapsedTime(&b, ctx->ev_pool[4 * k + 2], ctx->ev_pool[4 * k + 3]);
            search_ms += a; solve_ms += b;
            search_launches++; solve_launches++;
        }
    }
    int any_deg = 0;
    for (uint32_t s = 0; s < nseg; s++) {
        const SegState& hs = ctx->h_state[s];
        if (T_out) {
            float* o = T_out + 16 * (size_t)s;
            for (int i = 0; i < 12; i++) o[i] = hs.Tf[i];
            o[12] = o[13] = o[14] = 0.f; o[15] = 1.f;
        }
        if (hs.degenerate) any_deg = 1;
        if (stats) {
            icpsymm_stats& o = stats[s];
            memset(&o, 0, sizeof o);
            o.iters = hs.iter; o.converged = hs.converged; o.degenerate = hs.degenerate; o.searches = hs.searches;
            o.diff = hs.diff;
            uint32_t nq = S.h_counts[s];
            if (ctx->shard_world > 1) {
                const uint32_t tiles = icp_tiles(nq);
                uint32_t mine = 0;   // queries of this rank's parts
                for (uint32_t o = part0; o < part0 + nparts; o++)
                    for (uint32_t j = 0, n_o = icp_part_tiles(tiles, o); j < n_o; j++) {
                        const uint32_t first = icp_part_tile(tiles, o, j) * ICP_TILE;
                        mine += nq - first < ICP_TILE ? nq - first : ICP_TILE;
                    }
                nq = (use_comm || (ctx->peer_world > 1 && !single_search)) ? nq : mine;   // global sums: all source points
            }
            o.rmse = nq ? sqrt(hs.sumd2 / (double)nq) : 0.0;
            o.min_pivot_ratio = hs.min_ratio; o.pairs = hs.pairs;
            o.loop_ms = loop_ms; o.search_ms = search_ms; o.solve_ms = solve_ms;
            o.search_launches = search_launches; o.solve_launches = solve_launches;
        }
    }
    return any_deg ? ICPSYMM_WARN_DEGENERATE : ICPSYMM_OK;
}

int icp_run_single(icpsymm_ctx* ctx, const float* T16, const double* origin, icpsymm_stats* stats) {
    float T12v[12];
    if (T16) for (int i = 0; i < 12; i++) T12v[i] = T16[i];
    return icp_run_loop(ctx, 0, 0.f, T16 ? T12v : nullptr, origin, true, nullptr, stats);
}

int icp_export_correspondences(icpsymm_ctx* ctx, int32_t* idx_out, float* d2_out) {
    CloudSet& S = ctx->src;
    if (!ctx->have_corr || !S.valid || !ctx->tgt.valid) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "no search has run");
    size_t n = (size_t)S.nseg * S.max_count;
    int rc = ensure_io(ctx, n * 8);
    if (rc) return rc;
    int32_t* d_idx = (int32_t*)ctx->io_buf;
    float* d_d2 = (float*)((char*)ctx->io_buf + n * 4);
    cudaStream_t st = ctx->stream;
    ICP_CUDA_CHECK(cudaMemsetAsync(d_idx, 0xff, n * 4, st));
    ICP_CUDA_CHECK(cudaMemsetAsync(d_d2, 0, n * 4, st));
    size_t slots = (size_t)S.nseg * S.shape.pitch;
    export_corr_kernel<<<(uint32_t)((slots + 255) / 256), 256, 0, st>>>(S.view(), ctx->tgt.view(), ctx->d_state,
                                                                         ctx->corr_pos, S.max_count, d_idx, d_d2);
    ICP_CUDA_CHECK(cudaGetLastError());
    if (idx_out) ICP_CUDA_CHECK(cudaMemcpyAsync(idx_out, d_idx, n * 4, cudaMemcpyDeviceToHost, st));
    if (d2_out) ICP_CUDA_CHECK(cudaMemcpyAsync(d2_out, d_d2, n * 4, cudaMemcpyDeviceToHost, st));
    ICP_CUDA_CHECK(cudaStreamSynchronize(st));
    return ICPSYMM_OK;
}

int icp_export_transformed(icpsymm_ctx* ctx, const float* T, float* xyz_out, size_t stride_bytes) {
    CloudSet& S = ctx->src;
    { int jr = icp_join_clouds(ctx, true); if (jr < 0) return jr; }
    if (!S.valid) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "set_source first");
    if (!T || !xyz_out || stride_bytes < 12 || (stride_bytes & 3)) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "bad argument");
    size_t n = (size_t)S.nseg * S.max_count;
    int rc = ensure_io(ctx, n * stride_bytes);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    ICP_CUDA_CHECK(cudaMemsetAsync(ctx->io_buf, 0, n * stride_bytes, st));
    T12 t;
    for (int i = 0; i < 12; i++) t.m[i] = T[i];
    size_t slots = (size_t)S.nseg * S.shape.pitch;
    export_transformed_kernel<<<(uint32_t)((slots + 255) / 256), 256, 0, st>>>(S.view(), t, S.max_count,
                                                                                (float*)ctx->io_buf,
                                                                                (uint32_t)(stride_bytes / 4));
    ICP_CUDA_CHECK(cudaGetLastError());
    ICP_CUDA_CHECK(cudaMemcpyAsync(xyz_out, ctx->io_buf, n * stride_bytes, cudaMemcpyDeviceToHost, st));
    ICP_CUDA_CHECK(cudaStreamSynchronize(st));
    return ICPSYMM_OK;
}

int icp_read_sums(icpsymm_ctx* ctx, double* sums, double* origin_out) {
    double tmp[ICP_PART_STRIDE];
    ICP_CUDA_CHECK(cudaMemcpyAsync(tmp, ctx->gsum, sizeof(double) * ICP_PART_STRIDE, cudaMemcpyDeviceToHost, ctx->stream));
    ICP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < ICP_NSUMS; i++) sums[i] = tmp[i];
    if (origin_out) for (int i = 0; i < 3; i++) origin_out[i] = ctx->h_state[0].origin[i];
    return ICPSYMM_OK;
}

int icp_solve_compose_device(icpsymm_ctx* ctx, const double* sums, const double* origin, double* Tinc, double* x,
                             double* ratio) {
    int rc = ensure_io(ctx, sizeof(double) * 64);
    if (rc) return rc;
    double in[ICP_NSUMS + 3], out[20];
    for (int i = 0; i < ICP_NSUMS; i++) in[i] = sums[i];
    for (int i = 0; i < 3; i++) in[ICP_NSUMS + i] = origin ? origin[i] : 0.0;
    double* d_in = (double*)ctx->io_buf;
    double* d_out = d_in + 40;
    cudaStream_t st = ctx->stream;
    ICP_CUDA_CHECK(cudaMemcpyAsync(d_in, in, sizeof in, cudaMemcpyHostToDevice, st));
    solve_compose_kernel<<<1, 1, 0, st>>>(d_in, d_out);
    ICP_CUDA_CHECK(cudaGetLastError());
    ICP_CUDA_CHECK(cudaMemcpyAsync(out, d_out, sizeof out, cudaMemcpyDeviceToHost, st));
    ICP_CUDA_CHECK(cudaStreamSynchronize(st));
    if (Tinc) {
        for (int i = 0; i < 12; i++) Tinc[i] = out[i];
        Tinc[12] = Tinc[13] = Tinc[14] = 0.0; Tinc[15] = 1.0;
    }
    if (x) for (int i = 0; i < 6; i++) x[i] = out[12 + i];
    if (ratio) *ratio = out[18];
    return out[19] != 0.0 ? ICPSYMM_WARN_DEGENERATE : ICPSYMM_OK;
}

int icp_search_times(icpsymm_ctx* ctx, float* search_ms, float* solve_ms, int n) {
    if (!ctx->h_state || ctx->ev_pool_n == 0) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "no registration with ICPSYMM_OPT_KERNEL_TIMING has run");
    int k = 0;
    for (; k < n && 4 * k + 3 < ctx->ev_pool_n && k < ctx->timed_searches; k++) {
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, ctx->ev_pool[4 * k], ctx->ev_pool[4 * k + 1]) != cudaSuccess ||
            cudaEventElapsedTime(&b, ctx->ev_pool[4 * k + 2], ctx->ev_pool[4 * k + 3]) != cudaSuccess) { cudaGetLastError(); break; }
        if (search_ms) search_ms[k] = a;
        if (solve_ms) solve_ms[k] = b;
    }
    return k;
}

#ifdef ICP_EXPERIMENT   // developer build only (make exp): seed-quality experiments, never in the shipped library
namespace {
__global__ void seed_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, int mode) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t j = i;
    switch (mode) {
        case 1: j = i; break;                                   // own neighbour: perfect bound
        case 2: j = (i & ~31u) + 16u; break;                    // the unit centre's neighbour
        case 3: j = i >= 32u ? i - 32u : i; break;              // same lane of the previous unit
        case 4: j = i >= 256u ? i - 256u : i; break;            // same lane, previous 256-point tile
        case 5: j = i >= 32u ? ((i - 32u) & ~31u) + 16u : i; break;   // centre of the previous unit
        case 6: j = (i & ~255u) + 128u; break;                  // one probe per 256 points
        case 7: j = (i & ~1023u) + 512u; break;                 // one probe per 1024 points
        case 8: j = (i & ~7u) + 4u; break;                      // one probe per leaf group
    }
    if (j >= n) j = i;
    out[i] = mode == 0 ? 0xffffffffu : in[j];
}
}
extern "C" __attribute__((visibility("default"))) int icpsymm_debug_seed(icpsymm_ctx* ctx, int mode) {
    // mode -1: back up the current correspondences; mode >= 0: corr = f_mode(backup)
    static uint32_t* backup = nullptr; static size_t cap = 0;
    const size_t n = (size_t)ctx->src.shape.pitch;
    if (n > cap) { cudaFree(backup); cudaMalloc(&backup, n * 4); cap = n; }
    if (mode < 0) return (int)cudaMemcpyAsync(backup, ctx->corr_pos, n * 4, cudaMemcpyDeviceToDevice, ctx->stream);
    seed_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, ctx->stream>>>(backup, ctx->corr_pos, ctx->src.h_counts[0], mode);
    ctx->have_corr = true;
    return (int)cudaGetLastError();
}
extern "C" __attribute__((visibility("default"))) int icpsymm_debug_single_search(icpsymm_ctx* ctx, const float* T16, icpsymm_stats* stats) {
    return icp_run_single(ctx, T16, nullptr, stats);
}
#endif
