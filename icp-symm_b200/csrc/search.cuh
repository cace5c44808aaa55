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
#ifndef ICP_FAR_FACTOR
#define ICP_FAR_FACTOR 4.f               // far mode while the largest bound^2 exceeds this times the squared patch diagonal
#endif
#ifndef ICP_DENSE_MIN
#define ICP_DENSE_MIN 4                  // stage-1 survivors from which the all-children exact test is cheaper than per-child shuffles
#endif

// 256-bit loads (common.cuh: ldg8) for the leaf groups / the node records; a translation unit chooses before including
// this header (K1 gains 4-6 % per search; the normals kernel, at 72 registers, starts to spill with them)
#ifndef ICP_WIDE_LEAF
#define ICP_WIDE_LEAF 0
#endif
#ifndef ICP_WIDE_NODE
#define ICP_WIDE_NODE 0
#endif
__device__ __forceinline__ void load_leaf_group(const float4* __restrict__ lp, float4 (&c)[ICP_LEAF]) {
#if ICP_WIDE_LEAF
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j += 2) { const float4x2 t = ldg8(lp + j); c[j] = t.a; c[j + 1] = t.b; }
#else
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) c[j] = ldg4(lp + j);
#endif
}
__device__ __forceinline__ void load_node_record(const float4* __restrict__ p4, float4 (&r)[12]) {
#if ICP_WIDE_NODE
#pragma unroll
    for (int j = 0; j < 12; j += 2) { const float4x2 t = ldg8(p4 + j); r[j] = t.a; r[j + 1] = t.b; }
#else
#pragma unroll
    for (int j = 0; j < 12; j++) r[j] = ldg4(p4 + j);
#endif
}

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
    load_leaf_group(lp, c);
    const NegQuery nq = make_neg_query(qx, qy, qz);
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) best_offer_packed(b, c[j], nq);
}

// Warp-uniform leaf scan that builds the 64-bit (d2, index) keys only when some lane can use them: the 8
// squared distances and their minimum are plain fp32 work (FMNMX3); a point can replace a lane's best only
// when its d2 does not exceed the lane's bound (equal d2: the lower original index wins), so when no
// taking lane has min d2 <= bound the scan changes nothing.  Returns whether some taking lane's bound
// changed (warp-uniform).  `take` = false: the lane is a bystander and keeps its state.
//
// m2 (the "margin", see NNVisitor) collects the smallest squared distance of any point OTHER than the lane's
// best: in the common case every point of the group is farther than the bound, hence not the best, and the
// group's minimum is all that is needed; otherwise the loser of every key comparison is recorded.
template <bool TRACK>
__device__ __forceinline__ bool scan_leaf_lazy(const float4* __restrict__ seg_pts, uint32_t g, const NegQuery& nq,
                                               bool take, Best& b, uint32_t& m2) {
    const float4* lp = seg_pts + (size_t)g * ICP_LEAF;
    float4 c[ICP_LEAF];
    float d[ICP_LEAF];
    load_leaf_group(lp, c);
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) d[j] = dist2_packed(nq, c[j]);
    const float m = fminf(fminf(fminf(d[0], d[1]), fminf(d[2], d[3])), fminf(fminf(d[4], d[5]), fminf(d[6], d[7])));
    const uint32_t bound = (uint32_t)(b.key >> 32);
    if (!__any_sync(0xffffffffu, take && __float_as_uint(m) <= bound)) {
        if (TRACK) m2 = min(m2, __float_as_uint(m));   // all eight are beyond the bound: none of them is the best
        return false;
    }
    if (take) {
        const unsigned long long old = b.key;
#pragma unroll
        for (int j = 0; j < ICP_LEAF; j++) {
            const unsigned long long k = ((unsigned long long)__float_as_uint(d[j]) << 32) | __float_as_uint(c[j].w);
            b.key = k < b.key ? k : b.key;
        }
        // every point that is not the (new) best is an "other" point: the eight of this group, and the previous best
        // if it lost its place (the same point met twice -- a seed and its leaf -- has the best's own index)
        if (TRACK) {
            const uint32_t best_id = (uint32_t)b.key;
            if ((uint32_t)old != best_id && (uint32_t)old != (uint32_t)ICP_INVALID_ID) m2 = min(m2, (uint32_t)(old >> 32));
#pragma unroll
            for (int j = 0; j < ICP_LEAF; j++)
                m2 = min(m2, __float_as_uint(c[j].w) != best_id ? __float_as_uint(d[j]) : 0xffffffffu);
        }
    }
    return __any_sync(0xffffffffu, take && (uint32_t)(b.key >> 32) != bound);
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
    bool few;                            // at most ICP_FEW_LANES lanes are live
    bool single;                         // exactly one
};
#ifndef ICP_FEW_LANES
#define ICP_FEW_LANES 4
#endif

__device__ __forceinline__ float warp_min_f(float x) {
    return ordered_to_float(__reduce_min_sync(0xffffffffu, float_to_ordered(x)));
}
__device__ __forceinline__ float warp_max_f(float x) {
    return ordered_to_float(__reduce_max_sync(0xffffffffu, float_to_ordered(x)));
}

// A box is entered when its bound does not exceed the lane's bound AND is finite (padding boxes have bound +inf and
// a lane without a candidate yet has bound +inf).  Capping the lane's bound at the largest finite float turns the
// two comparisons per box into one.
__device__ __forceinline__ uint32_t bound_cap(uint32_t bound_bits) { return min(bound_bits, ICP_INF_BITS - 1u); }

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
                                                        uint32_t my_bound, uint32_t& kmin, uint32_t (&bb)[ICP_FAN]) {
    const float4* p4 = reinterpret_cast<const float4*>(nd);
    float4 r[12];
    load_node_record(p4, r);
    uint32_t mine = 0;
    kmin = 0xffffffffu;
    my_bound = bound_cap(my_bound);
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
            const bool pass = b[c] <= my_bound;
            bb[4 * h + c] = b[c];
            mine |= pass ? (1u << (4 * h + c)) : 0u;
            kmin = min(kmin, pass ? ((b[c] & ~7u) | (uint32_t)(4 * h + c)) : 0xffffffffu);
        }
    }
    return mine;
}

// far_mode: the bounds are much larger than the patch (first iterations of a badly aligned pair), so
// the patch box would prune next to nothing: go straight to the exact test of all 8 children.
// Every child the WARP does not enter is reported to the visitor with the lane's own lower bound of it
// (v.prune): that is how a lane learns how far away everything it did not look at must be.
template <typename V>
__device__ __forceinline__ uint32_t warp_node_test(const WideNode* __restrict__ nd, float qx, float qy, float qz,
                                                   uint32_t my_bound, bool active, const WarpPatch& pb,
                                                   uint32_t max_bound, bool far_mode, int& pref, V& v,
                                                   bool children_are_leaves = true) {
    // A walk with only a few live lanes (the lanes of a unit that could not keep their winners, K1 skip test) is "far"
    // by the patch criterion -- its patch is tiny -- but 32 exact tests per child are wasted on it: above the leaf
    // parents the patch test alone selects nearly the same children (a superset; the extra ones are cut one level
    // down) for a quarter of the instructions.  (For full packets in a far search the same shortcut costs 18 %: the
    // largest bound of 32 lanes is a poor stand-in for each of them.)
    // With ONE live lane (the probe) the patch is that lane's query and the patch test is its exact test, at every level.
    const bool patch_only = pb.single || (far_mode && !children_are_leaves && pb.few);
    if (far_mode && !patch_only) {
        uint32_t kmin, bb[ICP_FAN];
        const uint32_t mine = dense_children_test(nd, qx, qy, qz, my_bound, kmin, bb);
        pref = (int)(__reduce_min_sync(0xffffffffu, active ? kmin : 0xffffffffu) & 7u);
        const uint32_t mask = __reduce_or_sync(0xffffffffu, active ? mine : 0u);
        if (V::kTrackMargin) {
#pragma unroll
            for (int c = 0; c < ICP_FAN; c++) v.prune(((mask >> c) & 1u) ? 0xffffffffu : bb[c]);
        }
        return mask;
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
    if (patch_only) {
        if (V::kTrackMargin) v.prune(__reduce_min_sync(0xffffffffu, pass1 ? 0xffffffffu : lbp));
        return mask1;
    }
    // stage 2: exact, per lane, only for the survivors
    uint32_t mask2 = 0;
    if (__popc(mask1) >= ICP_DENSE_MIN) {
        uint32_t kmin, bb[ICP_FAN];
        const uint32_t mine = dense_children_test(nd, qx, qy, qz, my_bound, kmin, bb);
        mask2 = __reduce_or_sync(0xffffffffu, active ? mine : 0u) & mask1;
        if (V::kTrackMargin) {
#pragma unroll
            for (int c = 0; c < ICP_FAN; c++) v.prune(((mask2 >> c) & 1u) ? 0xffffffffu : bb[c]);
        }
    } else {
        // children cut by the patch test: the patch bound is a lower bound for every lane of the patch
        if (V::kTrackMargin) v.prune(__reduce_min_sync(0xffffffffu, pass1 ? 0xffffffffu : lbp));
        for (uint32_t mm = mask1; mm; mm &= mm - 1) {
            const int c = __ffs((int)mm) - 1;
            const float lx = __shfl_sync(0xffffffffu, blx, c), ly = __shfl_sync(0xffffffffu, bly, c);
            const float lz = __shfl_sync(0xffffffffu, blz, c), hx = __shfl_sync(0xffffffffu, bhx, c);
            const float hy = __shfl_sync(0xffffffffu, bhy, c), hz = __shfl_sync(0xffffffffu, bhz, c);
            const uint32_t lb = point_box_bits(lx, ly, lz, hx, hy, hz, qx, qy, qz);
            if (__any_sync(0xffffffffu, active && lb <= my_bound && lb != ICP_INF_BITS)) mask2 |= 1u << c;
            else if (V::kTrackMargin) v.prune(lb);
        }
    }
    return mask2;
}

// exact re-test of one child (slot c of record nd) under the current bounds
template <typename V>
__device__ __forceinline__ bool warp_child_recheck(const WideNode* __restrict__ nd, int c, float qx, float qy, float qz,
                                                   uint32_t my_bound, bool active, V& v) {
    const float* f = reinterpret_cast<const float*>(nd) + c;
    const uint32_t lb = point_box_bits(__ldg(f), __ldg(f + 8), __ldg(f + 16), __ldg(f + 24), __ldg(f + 32),
                                       __ldg(f + 40), qx, qy, qz);
    const bool enter = __any_sync(0xffffffffu, active && lb <= my_bound && lb != ICP_INF_BITS);
    if (V::kTrackMargin && !enter) v.prune(lb);
    return enter;
}

// The pending mask and the record pointer of the CURRENT node live in registers: a pop touches only
// those; the 64-bit store of the ancestors' masks is read and written once per node entered or left.
template <typename V>
__device__ __forceinline__ void wide_traverse_group(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                                    float qx, float qy, float qz, bool active, V& v) {
    const int top = (int)sh.n_levels - 1;
    if (top == 0) { v.leaf(0, active); return; }
    if (!__any_sync(0xffffffffu, active)) return;
    const float inf = __uint_as_float(ICP_INF_BITS);
    WarpPatch pb;
    pb.lox = warp_min_f(active ? qx : inf);  pb.loy = warp_min_f(active ? qy : inf);  pb.loz = warp_min_f(active ? qz : inf);
    pb.hix = warp_max_f(active ? qx : -inf); pb.hiy = warp_max_f(active ? qy : -inf); pb.hiz = warp_max_f(active ? qz : -inf);
    {
        const int live = __popc(__ballot_sync(0xffffffffu, active));
        pb.few = live <= ICP_FEW_LANES;
        pb.single = live == 1;
    }
    uint32_t max_bound = __reduce_max_sync(0xffffffffu, active ? v.bound_bits() : 0u);
    // far mode while the largest bound exceeds 4x the squared patch diagonal (both warp-uniform)
    const float pdx = pb.hix - pb.lox, pdy = pb.hiy - pb.loy, pdz = pb.hiz - pb.loz;
    const uint32_t far_thr = __float_as_uint(ICP_FAR_FACTOR * (pdx * pdx + pdy * pdy + pdz * pdz));
    bool far_mode = max_bound > far_thr;
    const WideNode* rec = seg_nodes + sh.level_off[top];
    uint32_t stale = 0;   // bit l: the pending mask of level l predates the latest bound improvement
    int l = top, pref;
    uint32_t i = 0;
    uint32_t cur = warp_node_test(rec, qx, qy, qz, v.bound_bits(), active, pb, max_bound, far_mode, pref, v, top == 1);
    unsigned long long saved = 0ull;   // byte l-1: pending mask of the open ancestor at level l
    for (;;) {
        if (V::kBudget && v.exhausted()) break;
        if (cur == 0u) {
            if (l == top) break;
            ++l; i >>= ICP_LOG_FAN;
            const int sh8 = 8 * (l - 1);
            cur = (uint32_t)(saved >> sh8) & 0xffu;
            rec = seg_nodes + sh.level_off[l] + i;
            pref = -1;
            continue;
        }
        const int c = (pref >= 0 && ((cur >> pref) & 1u)) ? pref : (__ffs((int)cur) - 1);
        pref = -1;
        cur &= ~(1u << c);
        if (((stale >> l) & 1u) && !warp_child_recheck(rec, c, qx, qy, qz, v.bound_bits(), active, v)) continue;
        const uint32_t child = (i << ICP_LOG_FAN) + (uint32_t)c;
        ICP_ASSERT(l >= 1 && l <= top && i < sh.level_cnt[l] && child < sh.level_cnt[l - 1]);
        bool improved = false;
        if (l == 1) {
            ICP_ASSERT(child < sh.pitch / ICP_LEAF);
            improved = v.leaf(child, active);
        } else {
            const int sh8 = 8 * (l - 1);
            saved = (saved & ~(0xffull << sh8)) | ((unsigned long long)cur << sh8);
            --l; i = child;
            rec = seg_nodes + sh.level_off[l] + i;
            cur = warp_node_test(rec, qx, qy, qz, v.bound_bits(), active, pb, max_bound, far_mode, pref, v, l == 1);
            stale &= ~(1u << l);
        }
        if (improved) {   // some active lane's bound improved (warp-uniform)
            stale = 0xffffffffu;
            max_bound = __reduce_max_sync(0xffffffffu, active ? v.bound_bits() : 0u);
            far_mode = max_bound > far_thr;
        }
    }
}


// split_mask != 0 (warp-uniform): the unit holds two far-apart clusters (index build, finish_units_kernel);
// the lanes of each cluster walk the hierarchy on their own, so that each walk sees a compact patch.
template <typename V>
__device__ __forceinline__ void wide_traverse_warp(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                                   float qx, float qy, float qz, bool active, uint32_t split_mask, V& v) {
    const bool second = (split_mask >> (threadIdx.x & 31)) & 1u;
    const int passes = split_mask ? 2 : 1;
#pragma unroll 1
    for (int pass = 0; pass < passes; pass++) {
        const bool act = active && (passes == 1 || second == (pass == 1));
        v.set_active(act);
        wide_traverse_group(sh, seg_nodes, qx, qy, qz, act, v);
    }
}

// m2, the MARGIN of the search: a lower bound (fp32 bits) of the squared distance from the query to every target
// point other than the winner -- the smallest of (a) the distances of all points scanned that did not win and
// (b) the lane's own lower bounds of all boxes the warp never entered.  Both are exact statements about the
// pinned fp32 arithmetic (a box bound never exceeds the fp32 distance of a member point).  K1 stores it and uses it
// in later searches: while the query has moved by less than the gap between the winner's distance and the margin,
// the winner cannot have changed and the walk is skipped (icp_loop.cu).
template <bool TRACK, bool BUDGET = false>
struct NNVisitor {
    static constexpr bool kTrackMargin = TRACK;
    static constexpr bool kBudget = BUDGET;   // the walk may stop early (probe: any real point is a useful seed)
    int budget = 0;                           // leaf scans left
    __device__ __forceinline__ bool exhausted() const { return BUDGET && budget <= 0; }
    Best& b;
    const float4* __restrict__ pts;
    float qx, qy, qz;
    NegQuery nq;
    uint32_t m2;
    __device__ __forceinline__ uint32_t bound_bits() const { return (uint32_t)(b.key >> 32); }
    __device__ __forceinline__ void set_active(bool) {}
    __device__ __forceinline__ void prune(uint32_t lower_bound_bits) { if (TRACK) m2 = min(m2, lower_bound_bits); }
    // warp-uniform call; returns whether some active lane's bound changed
    __device__ __forceinline__ bool leaf(uint32_t g, bool active) {
        if (BUDGET) budget--;
        return scan_leaf_lazy<TRACK>(pts, g, nq, active, b, m2);
    }
};

// full exact 1-NN search of all lanes of a warp; each `b` may already hold a candidate (warm start).
// Must be called by all 32 lanes; lanes without a query pass active = false.
// Out-of-line entry of the walk: its own register allocation, so that the margin-tracking flavour (more live
// state) does not tax the lean one when a kernel holds both.
struct WalkResult { unsigned long long key; uint32_t m2; };
template <bool TRACK>
__device__ __forceinline__ uint32_t nn_traverse_warp(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                                     const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                                     bool active, uint32_t split_mask, Best& b);
template <bool TRACK>
__device__ __noinline__ WalkResult nn_walk(const TreeShape* sh, const WideNode* __restrict__ seg_nodes,
                                           const float4* __restrict__ seg_pts, float qx, float qy, float qz, bool active,
                                           unsigned long long key) {
    Best b; b.key = key;
    WalkResult r;
    r.m2 = nn_traverse_warp<TRACK>(*sh, seg_nodes, seg_pts, qx, qy, qz, active, 0u, b);
    r.key = b.key;
    return r;
}

// Returns the lane's margin (see NNVisitor; TRACK = false: no margin, 0); a candidate already in `b` (seed) counts as seen.
template <bool TRACK>
__device__ __forceinline__ uint32_t nn_traverse_warp(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                                     const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                                     bool active, uint32_t split_mask, Best& b) {
    NNVisitor<TRACK> v{0, b, seg_pts, qx, qy, qz, make_neg_query(qx, qy, qz), TRACK ? ICP_INF_BITS : 0u};
    wide_traverse_warp(sh, seg_nodes, qx, qy, qz, active, split_mask, v);
    return v.m2;
}

// the probe's walk: exact unless it runs out of its budget of leaf scans, in which case it returns the best point
// met so far (a query next to the medial axis of the target would otherwise walk for a millisecond on its own)
static __device__ __noinline__ unsigned long long nn_walk_budget(const TreeShape* sh, const WideNode* __restrict__ seg_nodes,
                                                          const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                                          bool active, unsigned long long key, int budget) {
    Best b; b.key = key;
    NNVisitor<false, true> v{budget, b, seg_pts, qx, qy, qz, make_neg_query(qx, qy, qz), 0u};
    wide_traverse_warp(*sh, seg_nodes, qx, qy, qz, active, 0u, v);
    return b.key;
}

// ---- k nearest ------------------------------------------------------------------------------
// Per lane: (1) the ascending bit patterns of the k smallest squared distances seen so far, in
// registers -- only the VALUES, so that merging a candidate is a branch-free min/max chain (two
// instructions per slot, nothing to carry along); (2) the candidates themselves, parked in a
// per-lane shared-memory column as (d2 bits << 32 | sorted position) whenever they do not exceed
// the current k-th distance.  d2 >= 0, so its bit pattern orders like an unsigned integer.
//   flush    merges the distances of the newly parked candidates into the register list (all lanes
//            run the chain together, max-over-lanes(count) rounds);
//   compact  drops the parked candidates that the tightened k-th distance has overtaken; if more than
//            k remain (several candidates share the k-th distance) the largest (d2, original index)
//            is removed until k are left -- the oracle's tie rule, lowest original index first.
// After a final flush + compact the column holds exactly the k nearest neighbours.
// EXACT = the list is used with k == KMAX (the reference's k = 10, myicp.cpp:160).
// out of line on purpose: it almost never runs, and inlined at every compaction it was a quarter of the normals kernel's code
static __device__ __noinline__ int knn_drop_ties(unsigned long long* buf, int stride, int w, int k, const float4* __restrict__ pts) {
    while (w > k) {
        int worst = 0;
        unsigned long long worst_key = 0ull;
        for (int r = 0; r < w; r++) {
            const unsigned long long e = buf[r * stride];
            const unsigned long long key = (e & 0xffffffff00000000ull) | __float_as_uint(ldg4(pts + (uint32_t)e).w);
            if (key >= worst_key) { worst_key = key; worst = r; }
        }
        w--;
        buf[worst * stride] = buf[w * stride];
    }
    return w;
}
#define ICP_KNN_UNFILLED 0xfffffffeu     // above every distance, below the 0xffffffff of padding points
template <int KMAX, bool EXACT>
struct KList {
#ifndef ICP_KNN_CAP
#define ICP_KNN_CAP 32
#endif
    static constexpr int CAP = (KMAX <= 16) ? ICP_KNN_CAP : 48;   // parked candidates per lane (8 B each)
    uint32_t d[KMAX];
    uint32_t kth;              // d[k-1]: the current k-th smallest squared distance (bits)
    unsigned long long* buf;   // this lane's column: buf[slot * stride]
    int stride, cnt, done;     // cnt parked, the first `done` of them already merged into d[]

    __device__ __forceinline__ void init(unsigned long long* column, int column_stride) {
#pragma unroll
        for (int j = 0; j < KMAX; j++) d[j] = ICP_KNN_UNFILLED;
        kth = ICP_KNN_UNFILLED;
        buf = column; stride = column_stride; cnt = 0; done = 0;
    }
    // sorted insert of one value: slot j keeps min(d[j], max(d[j-1], v))
    __device__ __forceinline__ void insert(int k, uint32_t v) {
        uint32_t prev = 0u;
#pragma unroll
        for (int j = 0; j < KMAX; j++) {
            const uint32_t cur = d[j];
            d[j] = min(cur, max(prev, v));
            prev = cur;
        }
        if (EXACT) {
            kth = d[KMAX - 1];
        } else {
            // d[k-1] as a max over the first k slots (a select chain would become a dynamic local-memory index)
            uint32_t t = 0u;
#pragma unroll
            for (int j = 0; j < KMAX; j++) t = max(t, (j < k) ? d[j] : 0u);
            kth = t;
        }
    }
    __device__ __forceinline__ void flush(int k) {
        const int rounds = __reduce_max_sync(0xffffffffu, cnt - done);
        const uint32_t* hi = reinterpret_cast<const uint32_t*>(buf) + 1;   // high words: the distances
#pragma unroll 2
        for (int r = 0; r < rounds; r++) {
            const int idx = done + r;
            insert(k, idx < cnt ? hi[2 * idx * stride] : 0xffffffffu);
        }
        done = cnt;
    }
    __device__ __forceinline__ void compact(int k, const float4* __restrict__ pts) {
        const int n = __reduce_max_sync(0xffffffffu, cnt);
        // branch-free: every lane reads slot r (r < n <= CAP, so the slot exists; what a lane reads beyond its own
        // count is ignored) and keeps it under a predicate; the survivors stay in their order
        unsigned long long* wp = buf;
        const unsigned long long* rp = buf;
        int w = 0;
#pragma unroll 2
        for (int r = 0; r < n; r++, rp += stride) {
            const unsigned long long e = *rp;
            if (r < cnt && (uint32_t)(e >> 32) <= kth) { *wp = e; wp += stride; w++; }
        }
        ICP_ASSERT(w <= CAP && w <= cnt);
        // surplus ties at the k-th distance (rare): drop the largest original index until k remain
        if (w > k) w = knn_drop_ties(buf, stride, w, k, pts);
        cnt = done = w;
    }
    // merge and prune when some lane has no room left for another leaf group
    __device__ __forceinline__ void maybe_flush(int k, const float4* __restrict__ pts) {
        if (__any_sync(0xffffffffu, cnt > CAP - ICP_LEAF)) { flush(k); compact(k, pts); }
    }
    __device__ __forceinline__ uint32_t bound_bits() const { return kth; }
};

// offer the 8 points of leaf group g: park the ones that do not exceed the lane's current k-th distance
// (`take` = false makes the lane a bystander).  Until the next flush the pruning bound is the slightly
// stale, hence only looser, k-th distance of the last merge: exactness is unaffected.
template <int KMAX, bool EXACT>
__device__ __forceinline__ void knn_offer_leaf(KList<KMAX, EXACT>& L, const float4* __restrict__ pts, uint32_t g,
                                               float qx, float qy, float qz, bool take) {
    uint32_t db[ICP_LEAF];
    const float4* lp = pts + (size_t)g * ICP_LEAF;
    const NegQuery nq = make_neg_query(qx, qy, qz);
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) db[j] = __float_as_uint(dist2_packed(nq, ldg4(lp + j)));
    // padding points sit at (FLT_MAX, FLT_MAX, FLT_MAX): their squared distance is +inf, above the cap, so they are
    // never parked -- not even while the list is unfilled (kth = ICP_KNN_UNFILLED)
#ifdef ICP_KNN_CHECK_ID
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) if (__float_as_uint(ldg4(lp + j).w) == (uint32_t)ICP_INVALID_ID) db[j] = 0xffffffffu;
#endif
    const uint32_t thr = take ? bound_cap(L.kth) : 0u;
#pragma unroll
    for (int j = 0; j < ICP_LEAF; j++) {
        if (db[j] <= thr && take) {
            ICP_ASSERT(L.cnt >= 0 && L.cnt < (KList<KMAX, EXACT>::CAP));
            L.buf[L.cnt * L.stride] = ((unsigned long long)db[j] << 32) | (g * ICP_LEAF + (uint32_t)j);
            L.cnt++;
        }
    }
}

// Leaves [g_lo, g_hi] were already offered by the caller (the query's own neighbourhood along the
// curve, which gives a finite first bound) and are skipped so that no candidate is parked twice.
template <int KMAX, bool EXACT>
struct KnnVisitor {
    static constexpr bool kTrackMargin = false;
    static constexpr bool kBudget = false;
    __device__ __forceinline__ bool exhausted() const { return false; }
    __device__ __forceinline__ void prune(uint32_t) {}
    KList<KMAX, EXACT>& L;
    const float4* __restrict__ pts;
    float qx, qy, qz;
    int k;
    uint32_t g_lo, g_hi;
    bool active;
    __device__ __forceinline__ uint32_t bound_bits() const { return L.bound_bits(); }
    __device__ __forceinline__ void set_active(bool a) { active = a; }   // a bystander must not park candidates
    __device__ __forceinline__ bool leaf(uint32_t g, bool act) {
        // warp-uniform call; a lane sits out the leaves it has already seen
        const uint32_t before = L.bound_bits();
        knn_offer_leaf(L, pts, g, qx, qy, qz, active && !(g >= g_lo && g <= g_hi));
        L.maybe_flush(k, pts);
        return __any_sync(0xffffffffu, act && L.bound_bits() != before);
    }
};

// on return every lane's column holds exactly its k nearest neighbours (L.cnt of them)
template <int KMAX, bool EXACT>
__device__ __forceinline__ void knn_traverse_warp(const TreeShape& sh, const WideNode* __restrict__ seg_nodes,
                                                  const float4* __restrict__ seg_pts, float qx, float qy, float qz,
                                                  int k, uint32_t g_lo, uint32_t g_hi, bool active, uint32_t split_mask,
                                                  KList<KMAX, EXACT>& L) {
    KnnVisitor<KMAX, EXACT> v{L, seg_pts, qx, qy, qz, k, g_lo, g_hi, active};
    L.flush(k);                                   // the caller's window scan
    wide_traverse_warp(sh, seg_nodes, qx, qy, qz, active, split_mask, v);
    L.flush(k);
    L.compact(k, seg_pts);
}
