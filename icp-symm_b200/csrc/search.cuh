// search.cuh -- exact nearest-neighbour traversal of the implicit 8-ary wide-node hierarchy.
//
// Replaces the correspondence search the reference leaves as a todo (ICP/myicp.cpp:128-131)
// and the kd-tree PCL's NormalEstimation uses for its k=10 neighbourhoods (myicp.cpp:157-160).
//
// The hierarchy is implicit: node i of level l covers nodes [8i, 8i+8) of level l-1, level 0
// node g covers sorted points [8g, 8g+8).  One WideNode record holds the tight boxes of the 8
// children of an internal node, so a visit is 12 independent 16-byte loads followed by 8 bound
// tests, and the children that survive become an 8-bit pending mask.  The masks of all levels
// live in one 64-bit register (one byte per level), so the walk needs no stack in memory:
//   pop the lowest pending child of the current node; if it is a leaf group scan its 8 points,
//   otherwise test its 8 children and make it the current node; when a mask runs empty go up.
// Padding nodes carry empty boxes (lo=+inf, hi=-inf: bound=+inf) and are never entered.
// Candidates are ranked by (d2, original index): ties go to the lowest original index, and a
// subtree is skipped only when its bound is strictly greater than the current best d2.  The
// bound uses the same separately rounded operations as the distance, so it can never exceed the
// distance of a member point (monotone rounding): the search is exact, not approximate.
#pragma once
#include "common.cuh"

#define ICP_INF_BITS 0x7f800000u

struct Best {
    float d2;
    uint32_t id;   // original index of the best candidate
    uint32_t pos;  // its sorted position inside the segment
};

__device__ __forceinline__ void best_init(Best& b) {
    b.d2 = __uint_as_float(ICP_INF_BITS);
    b.id = ICP_INVALID_ID;
    b.pos = 0xffffffffu;
}

__device__ __forceinline__ void best_offer(Best& b, const float4& c, uint32_t pos, float qx, float qy, float qz) {
    float d = dist2_rn(qx, qy, qz, c.x, c.y, c.z);
    uint32_t id = __float_as_uint(c.w);
    if (d < b.d2 || (d == b.d2 && id < b.id)) { b.d2 = d; b.id = id; b.pos = pos; }
}

__device__ __forceinline__ void scan_leaf(const float4* __restrict__ seg_pts, uint32_t g, float qx, float qy,
                                          float qz, Best& b) {
    const float4* lp = seg_pts + (size_t)g * ICP_LEAF;
    float4 c[ICP_LEAF];
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) c[j] = ldg4(lp + j);
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) best_offer(b, c[j], g * ICP_LEAF + j, qx, qy, qz);
}

// bounds (as ordered bit patterns; every bound is >= +0) of the 8 children stored in one record
__device__ __forceinline__ void child_bounds(const WideNode* __restrict__ nd, float qx, float qy, float qz,
                                             uint32_t lbb[ICP_FAN]) {
    const float4* p = reinterpret_cast<const float4*>(nd);
    float4 r[12];
#pragma unroll
    for (int j = 0; j < 12; j++) r[j] = ldg4(p + j);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const float lox[4] = {r[h].x, r[h].y, r[h].z, r[h].w};
        const float loy[4] = {r[2 + h].x, r[2 + h].y, r[2 + h].z, r[2 + h].w};
        const float loz[4] = {r[4 + h].x, r[4 + h].y, r[4 + h].z, r[4 + h].w};
        const float hix[4] = {r[6 + h].x, r[6 + h].y, r[6 + h].z, r[6 + h].w};
        const float hiy[4] = {r[8 + h].x, r[8 + h].y, r[8 + h].z, r[8 + h].w};
        const float hiz[4] = {r[10 + h].x, r[10 + h].y, r[10 + h].z, r[10 + h].w};
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float ex = fmaxf(fmaxf(__fsub_rn(lox[c], qx), __fsub_rn(qx, hix[c])), 0.f);
            float ey = fmaxf(fmaxf(__fsub_rn(loy[c], qy), __fsub_rn(qy, hiy[c])), 0.f);
            float ez = fmaxf(fmaxf(__fsub_rn(loz[c], qz), __fsub_rn(qz, hiz[c])), 0.f);
            lbb[4 * h + c] =
                __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
        }
    }
}

// pending mask of the children whose bound does not exceed `bound_bits`, plus the nearest of them
// (pref, valid when the mask is non-zero): the walk enters the nearest child first so that the
// bound tightens as early as possible, the others follow in index order.
__device__ __forceinline__ uint32_t child_mask(const WideNode* __restrict__ nd, float qx, float qy, float qz,
                                               uint32_t bound_bits, int& pref) {
    uint32_t lbb[ICP_FAN];
    child_bounds(nd, qx, qy, qz, lbb);
    uint32_t m = 0, kmin = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < ICP_FAN; c++) {
        m |= (lbb[c] <= bound_bits && lbb[c] != ICP_INF_BITS) ? (1u << c) : 0u;
        kmin = min(kmin, (lbb[c] & ~7u) | (uint32_t)c);   // child index rides in the 3 low mantissa bits
    }
    pref = (int)(kmin & 7u);
    return m;
}

// Warp-synchronous variant: the 32 lanes hold 32 neighbouring queries (consecutive points of the
// Hilbert-sorted source, a compact patch after any rigid transform) and walk the hierarchy
// TOGETHER.  Every lane tests the same record against its own query and its own bound; a child
// stays pending when ANY lane still needs it (one REDUX.OR), the nearest child over all lanes is
// entered first (one REDUX.MIN).  Node records and leaf points are then warp-uniform loads, and
// there is no divergence: the warp pays for the union of the 32 search regions once instead of
// for 32 diverging walks.
__device__ __forceinline__ uint32_t warp_child_mask(const WideNode* __restrict__ nd, float qx, float qy, float qz,
                                                    uint32_t bound_bits, bool active, int& pref) {
    uint32_t lbb[ICP_FAN];
    child_bounds(nd, qx, qy, qz, lbb);
    uint32_t m = 0, kmin = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < ICP_FAN; c++) {
        const bool pass = lbb[c] <= bound_bits && lbb[c] != ICP_INF_BITS;
        m |= pass ? (1u << c) : 0u;
        kmin = min(kmin, pass ? ((lbb[c] & ~7u) | (uint32_t)c) : 0xffffffffu);
    }
    m = active ? m : 0u;
    kmin = active ? kmin : 0xffffffffu;
    m = __reduce_or_sync(0xffffffffu, m);
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    pref = (int)(kmin & 7u);   // only used when m != 0, i.e. when some lane passed a child
    return m;
}

template <typename V>
__device__ __forceinline__ void wide_traverse_warp(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                                   float qx, float qy, float qz, bool active, V& v) {
    const int top = (int)sh.n_levels - 1;
    if (top == 0) { v.leaf(0); return; }
    int l = top, pref;
    uint32_t i = 0;
    unsigned long long masks =
        (unsigned long long)warp_child_mask(seg_nodes + sh.level_off[top], qx, qy, qz, v.bound_bits(), active, pref)
        << (8 * (top - 1));
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
        const uint32_t child = (i << ICP_LOG_FAN) + (uint32_t)c;
        if (l == 1) {
            v.leaf(child);
        } else {
            --l; i = child;
            const uint32_t cm =
                warp_child_mask(seg_nodes + sh.level_off[l] + i, qx, qy, qz, v.bound_bits(), active, pref);
            masks |= (unsigned long long)cm << (8 * (l - 1));
        }
    }
}

// Generic exact walk.  V supplies bound_bits() (current pruning bound, ordered bits) and
// leaf(g) (offer the 8 points of leaf group g).
template <typename V>
__device__ __forceinline__ void wide_traverse(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                              float qx, float qy, float qz, V& v) {
    const int top = (int)sh.n_levels - 1;
    if (top == 0) { v.leaf(0); return; }       // a single leaf group, no internal node
    int l = top, pref;
    uint32_t i = 0;
    unsigned long long masks =
        (unsigned long long)child_mask(seg_nodes + sh.level_off[top], qx, qy, qz, v.bound_bits(), pref) << (8 * (top - 1));
    for (;;) {
        const int sh8 = 8 * (l - 1);
        const uint32_t m = (uint32_t)(masks >> sh8) & 0xffu;
        if (m == 0) {
            if (l == top) break;
            ++l; i >>= ICP_LOG_FAN;
            pref = -1;
            continue;
        }
        // first pop after a node was opened: its nearest child if it is pending (the index rides in
        // the low bound bits, so a near-tie may name a child that just missed the mask); afterwards
        // lowest index first
        const int c = (pref >= 0 && ((m >> pref) & 1u)) ? pref : (__ffs((int)m) - 1);
        pref = -1;
        masks &= ~(1ull << (sh8 + c));
        const uint32_t child = (i << ICP_LOG_FAN) + (uint32_t)c;
        if (l == 1) {
            v.leaf(child);
        } else {
            --l; i = child;
            const uint32_t cm = child_mask(seg_nodes + sh.level_off[l] + i, qx, qy, qz, v.bound_bits(), pref);
            masks |= (unsigned long long)cm << (8 * (l - 1));   // that byte is empty: its node was exhausted
        }
    }
}

struct NNVisitor {
    Best& b;
    const float4* __restrict__ pts;
    float qx, qy, qz;
    __device__ __forceinline__ uint32_t bound_bits() const { return __float_as_uint(b.d2); }
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

// per-thread variant (independent walks)
__device__ __forceinline__ void nn_traverse(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                            const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                            Best& b) {
    NNVisitor v{b, seg_pts, qx, qy, qz};
    wide_traverse(sh, seg_nodes, qx, qy, qz, v);
}

// ---- k nearest: ascending list of keys (d2 bits << 32 | original index), k <= KMAX at run time ----
// d2 >= 0 or +inf, so its bit pattern orders like an unsigned integer; an unfilled slot is ~0.
template <int KMAX>
struct KList {
    unsigned long long key[KMAX];
    unsigned long long kth;  // key[k-1]: the current k-th best

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < KMAX; j++) key[j] = ~0ull;
        kth = ~0ull;
    }
    __device__ __forceinline__ void offer(int k, float d2, uint32_t id) {
        unsigned long long kk = ((unsigned long long)__float_as_uint(d2) << 32) | id;
        if (kk >= kth) return;
        // carry-insert: every slot keeps the smaller of (carry, slot) and passes the larger on
        unsigned long long cur = kk;
#pragma unroll
        for (int j = 0; j < KMAX; j++) {
            if (j < k) {
                unsigned long long a = key[j];
                bool sw = cur < a;
                key[j] = sw ? cur : a;
                cur = sw ? a : cur;
            }
        }
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
    __device__ __forceinline__ uint32_t bound_bits() const { return (uint32_t)(kth >> 32); }
};

// Leaves [g_lo, g_hi] were already offered by the caller (the query's own neighbourhood along the curve,
// which gives a finite first bound) and are skipped so that no candidate enters the list twice.
// Sentinel padding points are ignored.
template <int KMAX>
struct KnnVisitor {
    KList<KMAX>& L;
    const float4* __restrict__ pts;
    float qx, qy, qz;
    int k;
    uint32_t g_lo, g_hi;
    __device__ __forceinline__ uint32_t bound_bits() const { return L.bound_bits(); }
    __device__ __forceinline__ void leaf(uint32_t g) {
        if (g >= g_lo && g <= g_hi) return;
        const float4* lp = pts + (size_t)g * ICP_LEAF;
        float4 c[ICP_LEAF];
#pragma unroll
        for (int j = 0; j < ICP_LEAF; j++) c[j] = ldg4(lp + j);
#pragma unroll
        for (int j = 0; j < ICP_LEAF; j++) {
            uint32_t id = __float_as_uint(c[j].w);
            if (id != (uint32_t)ICP_INVALID_ID) L.offer(k, dist2_rn(qx, qy, qz, c[j].x, c[j].y, c[j].z), id);
        }
    }
};

template <int KMAX>
__device__ __forceinline__ void knn_traverse_warp(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                                  const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                                  int k, uint32_t g_lo, uint32_t g_hi, bool active, KList<KMAX>& L) {
    KnnVisitor<KMAX> v{L, seg_pts, qx, qy, qz, k, g_lo, g_hi};
    wide_traverse_warp(sh, seg_nodes, qx, qy, qz, active, v);
}

template <int KMAX>
__device__ __forceinline__ void knn_traverse(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                             const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                             int k, uint32_t g_lo, uint32_t g_hi, KList<KMAX>& L) {
    KnnVisitor<KMAX> v{L, seg_pts, qx, qy, qz, k, g_lo, g_hi};
    wide_traverse(sh, seg_nodes, qx, qy, qz, v);
}
