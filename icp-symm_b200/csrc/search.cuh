// search.cuh -- exact nearest-neighbour traversal of the implicit 8-ary wide-node hierarchy.
//
// Replaces the correspondence search the reference leaves as a todo (ICP/myicp.cpp:128-131)
// and the kd-tree PCL's NormalEstimation uses for its k=10 neighbourhoods (myicp.cpp:157-160).
//
// The hierarchy is implicit: node i of level l covers nodes [8i, 8i+8) of level l-1, level 0
// node g covers sorted points [8g, 8g+8).  One WideNode record holds the tight boxes of the 8
// children of an internal node.  The pending children of every open level are an 8-bit mask and
// the masks of all levels live in one 64-bit register, so the walk needs no stack in memory:
//   pop a pending child of the current node (nearest first, then lowest index); if it is a leaf
//   group scan its 8 points, otherwise test its 8 children and make it the current node; when a
//   mask runs empty go up (parent index = i >> 3).
// Padding nodes carry empty boxes (lo=+inf, hi=-inf: bound=+inf) and are never entered.
// Candidates are ranked by (d2, original index): ties go to the lowest original index, and a
// subtree is skipped only when its bound is strictly greater than the current best d2.  The
// bound uses the same separately rounded operations as the distance, so it can never exceed the
// distance of a member point (monotone rounding): the search is exact, not approximate.
#pragma once
#include "common.cuh"

#define ICP_INF_BITS 0x7f800000u
#define ICP_KNN_BUF 32                   // parked kNN candidates per lane (shared memory, 8 B each)

// best candidate so far as ONE 64-bit key, (d2 bits << 32) | original index: d2 >= 0, so unsigned
// integer order on the key is exactly the (d2, lowest original index) order the oracle defines, and
// an update is a 64-bit min (two compares, two selects) instead of a three-way predicate.
struct Best {
    unsigned long long key;
    __device__ __forceinline__ float d2() const { return __uint_as_float((uint32_t)(key >> 32)); }
    __device__ __forceinline__ uint32_t id() const { return (uint32_t)key; }
};

__device__ __forceinline__ void best_init(Best& b) {
    b.key = ((unsigned long long)ICP_INF_BITS << 32) | (unsigned long long)ICP_INVALID_ID;
}

__device__ __forceinline__ void best_offer(Best& b, const float4& c, float qx, float qy, float qz) {
    const float d = dist2_rn(qx, qy, qz, c.x, c.y, c.z);
    const unsigned long long k = ((unsigned long long)__float_as_uint(d) << 32) | __float_as_uint(c.w);
    b.key = k < b.key ? k : b.key;
}
__device__ __forceinline__ void best_offer_packed(Best& b, const float4& c, const NegQuery& nq) {
    const float d = dist2_packed(nq, c);
    const unsigned long long k = ((unsigned long long)__float_as_uint(d) << 32) | __float_as_uint(c.w);
    b.key = k < b.key ? k : b.key;
}

__device__ __forceinline__ void scan_leaf(const float4* __restrict__ seg_pts, uint32_t g, float qx, float qy,
                                          float qz, Best& b) {
    const float4* lp = seg_pts + (size_t)g * ICP_LEAF;
    float4 c[ICP_LEAF];
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) c[j] = ldg4(lp + j);
    const NegQuery nq = make_neg_query(qx, qy, qz);
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) best_offer_packed(b, c[j], nq);
}

// Warp-synchronous variant: the 32 lanes hold 32 neighbouring queries (consecutive points of the
// Hilbert-sorted source, a compact patch after any rigid transform) and walk the hierarchy
// TOGETHER, so node records and leaf points are warp-uniform loads and nothing diverges: the warp
// pays once for the union of the 32 search regions instead of for 32 diverging walks.
//
// A node visit is two-stage.  Stage 1: lane c (mod 8) tests child c against the bounding box of
// the whole patch and the largest bound of any lane -- one box-box test per lane, and a
// conservative superset because every step of the fp32 bound is monotone.  Stage 2: only the
// children that survive are tested exactly, every lane against its own query and its own bound
// (box broadcast by shuffles); a child stays pending when ANY lane still needs it.  Near
// convergence one or two children survive stage 1, which halves the cost of a visit.
//
// Pending masks are computed under the bounds of that moment.  When a leaf scan improves a bound
// the masks of all open levels become stale; a stale child is re-tested exactly (one box per lane)
// when it is popped, before paying for a full visit.
struct WarpPatch {
    float lox, loy, loz, hix, hiy, hiz;  // bounding box of the active lanes' queries (warp-uniform)
};

__device__ __forceinline__ float warp_min_f(float x) {
    return ordered_to_float(__reduce_min_sync(0xffffffffu, float_to_ordered(x)));
}
__device__ __forceinline__ float warp_max_f(float x) {
    return ordered_to_float(__reduce_max_sync(0xffffffffu, float_to_ordered(x)));
}

__device__ __forceinline__ uint32_t point_box_bits(float lx, float ly, float lz, float hx, float hy, float hz,
                                                   float qx, float qy, float qz) {
    float ex = fmaxf(fmaxf(__fsub_rn(lx, qx), __fsub_rn(qx, hx)), 0.f);
    float ey = fmaxf(fmaxf(__fsub_rn(ly, qy), __fsub_rn(qy, hy)), 0.f);
    float ez = fmaxf(fmaxf(__fsub_rn(lz, qz), __fsub_rn(qz, hz)), 0.f);
    return __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
}

// exact test of all 8 children of one record, every lane against its own query and bound: 12
// warp-uniform 16-byte loads, packed arithmetic, no shuffles and no loop.  Returns the lane's
// pass mask; kmin receives the lane's nearest passing child as (bound & ~7) | child.
__device__ __forceinline__ uint32_t dense_children_test(const WideNode* __restrict__ nd, float qx, float qy, float qz,
                                                        uint32_t my_bound, uint32_t& kmin) {
    const float4* p4 = reinterpret_cast<const float4*>(nd);
    float4 r[12];
#pragma unroll
    for (int j = 0; j < 12; j++) r[j] = ldg4(p4 + j);
    uint32_t mine = 0;
    kmin = 0xffffffffu;
    const NegQuery nq = make_neg_query(qx, qy, qz);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t b[4];   // children 4h .. 4h+3, two per packed evaluation
        box_bounds2(r[h].x, r[h].y, r[2 + h].x, r[2 + h].y, r[4 + h].x, r[4 + h].y, r[6 + h].x, r[6 + h].y,
                    r[8 + h].x, r[8 + h].y, r[10 + h].x, r[10 + h].y, nq, b[0], b[1]);
        box_bounds2(r[h].z, r[h].w, r[2 + h].z, r[2 + h].w, r[4 + h].z, r[4 + h].w, r[6 + h].z, r[6 + h].w,
                    r[8 + h].z, r[8 + h].w, r[10 + h].z, r[10 + h].w, nq, b[2], b[3]);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const bool pass = b[c] <= my_bound && b[c] != ICP_INF_BITS;
            mine |= pass ? (1u << (4 * h + c)) : 0u;
            kmin = min(kmin, pass ? ((b[c] & ~7u) | (uint32_t)(4 * h + c)) : 0xffffffffu);
        }
    }
    return mine;
}

// far_mode: the bounds are much larger than the patch (first iterations of a badly aligned pair), so
// the patch box would prune next to nothing: go straight to the exact test of all 8 children.
__device__ __forceinline__ uint32_t warp_node_test(const WideNode* __restrict__ nd, float qx, float qy, float qz,
                                                   uint32_t my_bound, bool active, const WarpPatch& pb,
                                                   uint32_t max_bound, bool far_mode, int& pref) {
    if (far_mode) {
        uint32_t kmin;
        const uint32_t mine = dense_children_test(nd, qx, qy, qz, my_bound, kmin);
        pref = (int)(__reduce_min_sync(0xffffffffu, active ? kmin : 0xffffffffu) & 7u);
        return __reduce_or_sync(0xffffffffu, active ? mine : 0u);
    }
    const int cl = threadIdx.x & 7;
    const float* f = reinterpret_cast<const float*>(nd);
    const float blx = __ldg(f + cl), bly = __ldg(f + 8 + cl), blz = __ldg(f + 16 + cl);
    const float bhx = __ldg(f + 24 + cl), bhy = __ldg(f + 32 + cl), bhz = __ldg(f + 40 + cl);
    // stage 1: box(patch) vs box(child cl).  For every query q of the patch lo-q >= lo-patch.hi and
    // q-hi >= patch.lo-hi, and rounding is monotone, so this bound never exceeds a lane's own bound.
    float ex = fmaxf(fmaxf(__fsub_rn(blx, pb.hix), __fsub_rn(pb.lox, bhx)), 0.f);
    float ey = fmaxf(fmaxf(__fsub_rn(bly, pb.hiy), __fsub_rn(pb.loy, bhy)), 0.f);
    float ez = fmaxf(fmaxf(__fsub_rn(blz, pb.hiz), __fsub_rn(pb.loz, bhz)), 0.f);
    const uint32_t lbp = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
    const bool pass1 = lbp <= max_bound && lbp != ICP_INF_BITS;
    const uint32_t mask1 = __ballot_sync(0xffffffffu, pass1) & 0xffu;
    const uint32_t kmin1 = __reduce_min_sync(0xffffffffu, pass1 ? ((lbp & ~7u) | (uint32_t)cl) : 0xffffffffu);
    pref = (int)(kmin1 & 7u);
    // stage 2: exact, per lane, only for the survivors
    uint32_t mask2 = 0;
    if (__popc(mask1) > 3) {
        uint32_t kmin;
        const uint32_t mine = dense_children_test(nd, qx, qy, qz, my_bound, kmin);
        mask2 = __reduce_or_sync(0xffffffffu, active ? mine : 0u) & mask1;
    } else {
        for (uint32_t mm = mask1; mm; mm &= mm - 1) {
            const int c = __ffs((int)mm) - 1;
            const float lx = __shfl_sync(0xffffffffu, blx, c), ly = __shfl_sync(0xffffffffu, bly, c);
            const float lz = __shfl_sync(0xffffffffu, blz, c), hx = __shfl_sync(0xffffffffu, bhx, c);
            const float hy = __shfl_sync(0xffffffffu, bhy, c), hz = __shfl_sync(0xffffffffu, bhz, c);
            const uint32_t lb = point_box_bits(lx, ly, lz, hx, hy, hz, qx, qy, qz);
            if (__any_sync(0xffffffffu, active && lb <= my_bound && lb != ICP_INF_BITS)) mask2 |= 1u << c;
        }
    }
    return mask2;
}

// exact re-test of one child (slot c of record nd) under the current bounds
__device__ __forceinline__ bool warp_child_recheck(const WideNode* __restrict__ nd, int c, float qx, float qy, float qz,
                                                   uint32_t my_bound, bool active) {
    const float* f = reinterpret_cast<const float*>(nd) + c;
    const uint32_t lb = point_box_bits(__ldg(f), __ldg(f + 8), __ldg(f + 16), __ldg(f + 24), __ldg(f + 32),
                                       __ldg(f + 40), qx, qy, qz);
    return __any_sync(0xffffffffu, active && lb <= my_bound && lb != ICP_INF_BITS);
}

template <typename V>
__device__ __forceinline__ void wide_traverse_warp(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                                   float qx, float qy, float qz, bool active, V& v) {
    const int top = (int)sh.n_levels - 1;
    if (top == 0) { v.leaf(0); return; }
    if (!__any_sync(0xffffffffu, active)) return;
    const float inf = __uint_as_float(ICP_INF_BITS);
    WarpPatch pb;
    pb.lox = warp_min_f(active ? qx : inf);  pb.loy = warp_min_f(active ? qy : inf);  pb.loz = warp_min_f(active ? qz : inf);
    pb.hix = warp_max_f(active ? qx : -inf); pb.hiy = warp_max_f(active ? qy : -inf); pb.hiz = warp_max_f(active ? qz : -inf);
    uint32_t max_bound = __reduce_max_sync(0xffffffffu, active ? v.bound_bits() : 0u);
    // far mode while the largest bound exceeds 4x the squared patch diagonal (both warp-uniform)
    const float pdx = pb.hix - pb.lox, pdy = pb.hiy - pb.loy, pdz = pb.hiz - pb.loz;
    const uint32_t far_thr = __float_as_uint(4.f * (pdx * pdx + pdy * pdy + pdz * pdz));
    bool far_mode = max_bound > far_thr;
    uint32_t stale = 0;   // bit l: the pending mask of level l predates the latest bound improvement
    int l = top, pref;
    uint32_t i = 0;
    unsigned long long masks =
        (unsigned long long)warp_node_test(seg_nodes + sh.level_off[top], qx, qy, qz, v.bound_bits(), active, pb,
                                           max_bound, far_mode, pref) << (8 * (top - 1));
    for (;;) {
        const int sh8 = 8 * (l - 1);
        const uint32_t m = (uint32_t)(masks >> sh8) & 0xffu;
        if (m == 0) {
            if (l == top) break;
            ++l; i >>= ICP_LOG_FAN;
            pref = -1;
            continue;
        }
        const int c = (pref >= 0 && ((m >> pref) & 1u)) ? pref : (__ffs((int)m) - 1);
        pref = -1;
        masks &= ~(1ull << (sh8 + c));
        const WideNode* rec = seg_nodes + sh.level_off[l] + i;
        if (((stale >> l) & 1u) && !warp_child_recheck(rec, c, qx, qy, qz, v.bound_bits(), active)) continue;
        const uint32_t child = (i << ICP_LOG_FAN) + (uint32_t)c;
        if (l == 1) {
            const uint32_t before = v.bound_bits();
            v.leaf(child);
            const uint32_t after = v.bound_bits();
            if (__any_sync(0xffffffffu, active && after != before)) {
                stale = 0xffffffffu;
                max_bound = __reduce_max_sync(0xffffffffu, active ? after : 0u);
                far_mode = max_bound > far_thr;
            }
        } else {
            --l; i = child;
            const uint32_t cm = warp_node_test(seg_nodes + sh.level_off[l] + i, qx, qy, qz, v.bound_bits(), active, pb,
                                               max_bound, far_mode, pref);
            masks |= (unsigned long long)cm << (8 * (l - 1));
            stale &= ~(1u << l);
        }
    }
}

struct NNVisitor {
    Best& b;
    const float4* __restrict__ pts;
    float qx, qy, qz;
    __device__ __forceinline__ uint32_t bound_bits() const { return (uint32_t)(b.key >> 32); }
    __device__ __forceinline__ void leaf(uint32_t g) { scan_leaf(pts, g, qx, qy, qz, b); }
};

// full exact 1-NN search of all lanes of a warp; each `b` may already hold a candidate (warm start).
// Must be called by all 32 lanes; lanes without a query pass active = false.
__device__ __forceinline__ void nn_traverse_warp(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                                 const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                                 bool active, Best& b) {
    NNVisitor v{b, seg_pts, qx, qy, qz};
    wide_traverse_warp(sh, seg_nodes, qx, qy, qz, active, v);
}

// ---- k nearest: ascending list of keys (d2 bits << 32 | original index), k <= KMAX at run time ----
// d2 >= 0 or +inf, so its bit pattern orders like an unsigned integer; an unfilled slot is ~0.
// EXACT = the list is used with k == KMAX (the reference's k = 10, myicp.cpp:160): no per-slot
// `j < k` predicate and the k-th best is simply the last slot.
template <int KMAX, bool EXACT>
struct KList {
    unsigned long long key[KMAX];
    unsigned long long kth;  // key[k-1]: the current k-th best
    // candidates that beat kth are first parked in a per-lane shared-memory column and merged into
    // the list in bursts (flush): the long insertion body then runs max-over-lanes(count) times per
    // burst with most lanes busy, instead of once per candidate with one or two lanes busy
    unsigned long long* buf;   // this lane's column: buf[slot * stride]
    int stride, cnt;

    __device__ __forceinline__ void init(unsigned long long* column, int column_stride) {
#pragma unroll
        for (int j = 0; j < KMAX; j++) key[j] = ~0ull;
        kth = ~0ull;
        buf = column; stride = column_stride; cnt = 0;
    }
    __device__ __forceinline__ void flush(int k) {
        const int rounds = __reduce_max_sync(0xffffffffu, cnt);
#pragma unroll 1
        for (int r = 0; r < rounds; r++) {
            const unsigned long long cand = (r < cnt) ? buf[r * stride] : ~0ull;
            if (cand < kth) insert(k, cand);
        }
        cnt = 0;
    }
    // merge when some lane has no room left for another leaf group
    __device__ __forceinline__ void maybe_flush(int k) {
        if (__any_sync(0xffffffffu, cnt > ICP_KNN_BUF - ICP_LEAF)) flush(k);
    }
    // sorted insert of a key that is known to be < kth: slot j keeps its key while it is smaller
    // than the new one, otherwise it takes the new key (first such slot) or its left neighbour
    __device__ __forceinline__ void insert(int k, unsigned long long kk) {
        bool prev_less = true;   // "key[j-1] < kk" for j = 0
        unsigned long long prev_key = 0ull;
#pragma unroll
        for (int j = 0; j < KMAX; j++) {
            const unsigned long long cur = key[j];
            const bool less = cur < kk;
            if (EXACT || j < k) key[j] = less ? cur : (prev_less ? kk : prev_key);
            prev_less = less;
            prev_key = cur;
        }
        if (EXACT) {
            kth = key[KMAX - 1];
        } else {
            // k-th best = max of the first k slots (written as a max so the compiler keeps key[] in
            // registers instead of turning a select chain into a dynamic local-memory index)
            unsigned long long t = 0ull;
#pragma unroll
            for (int j = 0; j < KMAX; j++) {
                unsigned long long c = (j < k) ? key[j] : 0ull;
                t = c > t ? c : t;
            }
            kth = t;
        }
    }
    __device__ __forceinline__ void offer(int k, float d2, uint32_t id) {
        const unsigned long long kk = ((unsigned long long)__float_as_uint(d2) << 32) | id;
        if (kk < kth) insert(k, kk);
    }
    __device__ __forceinline__ uint32_t bound_bits() const { return (uint32_t)(kth >> 32); }
};

// offer the 8 points of one leaf group: form the 8 keys and park the ones that beat the lane's current
// k-th best; the caller merges (maybe_flush) when a lane's column is nearly full.  Until the merge the pruning bound is the
// (slightly stale, hence only looser) k-th best of the last merge: exactness is unaffected.
template <int KMAX, bool EXACT>
__device__ __forceinline__ void knn_offer_leaf(KList<KMAX, EXACT>& L, int k, const float4* __restrict__ lp,
                                               float qx, float qy, float qz) {
    unsigned long long kk[ICP_LEAF];
    const NegQuery nq = make_neg_query(qx, qy, qz);
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) {
        const float4 c = ldg4(lp + j);
        const uint32_t id = __float_as_uint(c.w);
        const float d = dist2_packed(nq, c);
        kk[j] = (id != (uint32_t)ICP_INVALID_ID) ? (((unsigned long long)__float_as_uint(d) << 32) | id) : ~0ull;
    }
    // park the candidates that beat the current k-th best (warp-uniform call: every lane runs this)
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) {
        if (kk[j] < L.kth) { L.buf[L.cnt * L.stride] = kk[j]; L.cnt++; }
    }
}

// Leaves [g_lo, g_hi] were already offered by the caller (the query's own neighbourhood along the
// curve, which gives a finite first bound) and are skipped so that no candidate enters the list
// twice.  Sentinel padding points are ignored.
template <int KMAX, bool EXACT>
struct KnnVisitor {
    KList<KMAX, EXACT>& L;
    const float4* __restrict__ pts;
    float qx, qy, qz;
    int k;
    uint32_t g_lo, g_hi;
    bool active;
    __device__ __forceinline__ uint32_t bound_bits() const { return L.bound_bits(); }
    __device__ __forceinline__ void leaf(uint32_t g) {
        // warp-uniform call; a lane skips the leaves it has already seen by making its keys inert
        const bool skip = !active || (g >= g_lo && g <= g_hi);
        const unsigned long long saved = L.kth;
        if (skip) L.kth = 0ull;                 // nothing beats 0: no insertion for this lane
        knn_offer_leaf(L, k, pts + (size_t)g * ICP_LEAF, qx, qy, qz);
        if (skip) L.kth = saved;
        L.maybe_flush(k);
    }
};

template <int KMAX, bool EXACT>
__device__ __forceinline__ void knn_traverse_warp(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                                  const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                                  int k, uint32_t g_lo, uint32_t g_hi, bool active,
                                                  KList<KMAX, EXACT>& L) {
    KnnVisitor<KMAX, EXACT> v{L, seg_pts, qx, qy, qz, k, g_lo, g_hi, active};
    L.flush(k);                                   // the caller's window scan
    wide_traverse_warp(sh, seg_nodes, qx, qy, qz, active, v);
    L.flush(k);
}
