// search.cuh -- exact nearest-neighbour traversal of the implicit F-ary box hierarchy.
//
// Replaces the correspondence search the reference leaves as a todo (ICP/myicp.cpp:128-131)
// and the kd-tree PCL's NormalEstimation uses for its k=10 neighbourhoods (myicp.cpp:157-160).
//
// The hierarchy is implicit: node i of level l covers nodes [i*F, i*F+F) of level l-1, level 0
// node g covers sorted points [g*LEAF, g*LEAF+LEAF).  Every level is padded to a multiple of F
// with empty boxes (lo=+inf, hi=-inf, bound=+inf), so traversal needs no stack and no per-level
// counts: "next node" is i+1, and when i+1 crosses a multiple of F the walk pops to the parent.
// Candidates are ranked by (d2, original index) so ties go to the lowest original index, and a
// subtree is skipped only when its bound is strictly greater than the current best d2.
#pragma once
#include "common.cuh"

struct Best {
    float d2;
    uint32_t id;   // original index of the best candidate
    uint32_t pos;  // its sorted position inside the segment
};

__device__ __forceinline__ void best_init(Best& b) {
    b.d2 = __int_as_float(0x7f800000);
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

// descend to the leaf whose chain of boxes is nearest at every level: a cheap first bound
__device__ __forceinline__ void greedy_descent(const TreeShape& sh, const Box* __restrict__ seg_boxes,
                                               const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                               Best& b) {
    uint32_t node = 0;
    for (int l = (int)sh.n_levels - 1; l > 0; --l) {
        const Box* cb = seg_boxes + sh.level_off[l - 1] + node * ICP_FAN;
        float blb = __int_as_float(0x7f800000);
        uint32_t bc = 0;
#pragma unroll
        for (int c = 0; c < ICP_FAN; c++) {
            float4 lo = ldg4(&cb[c].lo), hi = ldg4(&cb[c].hi);
            float lb = box_dist2_rn(lo, hi, qx, qy, qz);
            if (lb < blb) { blb = lb; bc = c; }
        }
        node = node * ICP_FAN + bc;
    }
    scan_leaf(seg_pts, node, qx, qy, qz, b);
}

// full exact search; `b` may already hold a candidate (warm start / greedy descent)
__device__ __forceinline__ void nn_traverse(const TreeShape& sh, const Box* __restrict__ seg_boxes,
                                            const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                            Best& b) {
    const int top = (int)sh.n_levels - 1;
    int l = top;
    uint32_t i = 0;
    for (;;) {
        const Box* nb = seg_boxes + sh.level_off[l] + i;
        float4 lo = ldg4(&nb->lo), hi = ldg4(&nb->hi);
        float lb = box_dist2_rn(lo, hi, qx, qy, qz);
        if (lb <= b.d2 && lb < __int_as_float(0x7f800000)) {  // empty (padding) boxes have bound +inf
            if (l > 0) { --l; i <<= ICP_LOG_FAN; continue; }
            scan_leaf(seg_pts, i, qx, qy, qz, b);
        }
        ++i;
        while (l < top && (i & (ICP_FAN - 1)) == 0) { i >>= ICP_LOG_FAN; ++l; }
        if (l == top) break;
    }
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

// exact k nearest of (qx,qy,qz) inside one segment.  Leaves [g_lo, g_hi] were already offered by
// the caller (the query's own Morton neighbourhood, which gives a finite first bound) and are
// skipped here so that no candidate enters the list twice.  Sentinel padding points are ignored.
template <int KMAX>
__device__ __forceinline__ void knn_traverse(const TreeShape& sh, const Box* __restrict__ seg_boxes,
                                             const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                             int k, uint32_t g_lo, uint32_t g_hi, KList<KMAX>& L) {
    const int top = (int)sh.n_levels - 1;
    int l = top;
    uint32_t i = 0;
    for (;;) {
        const Box* nb = seg_boxes + sh.level_off[l] + i;
        float4 lo = ldg4(&nb->lo), hi = ldg4(&nb->hi);
        uint32_t lbb = __float_as_uint(box_dist2_rn(lo, hi, qx, qy, qz));
        if (lbb <= L.bound_bits() && lbb != 0x7f800000u) {
            if (l > 0) { --l; i <<= ICP_LOG_FAN; continue; }
            if (i < g_lo || i > g_hi) {
                const float4* lp = seg_pts + (size_t)i * ICP_LEAF;
#pragma unroll
                for (int j = 0; j < ICP_LEAF; j++) {
                    float4 c = ldg4(lp + j);
                    uint32_t id = __float_as_uint(c.w);
                    if (id != ICP_INVALID_ID) L.offer(k, dist2_rn(qx, qy, qz, c.x, c.y, c.z), id);
                }
            }
        }
        ++i;
        while (l < top && (i & (ICP_FAN - 1)) == 0) { i >>= ICP_LOG_FAN; ++l; }
        if (l == top) break;
    }
}
