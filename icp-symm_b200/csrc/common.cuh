// common.cuh -- shared types and device helpers of the B200 symmetric-ICP path.
//
// Data layout in HBM (DESIGN.md "Layout"): every cloud set is a batch of `nseg` independent
// clouds ("segments"; a single pair is nseg = 1) stored with a fixed pitch of P points:
//   pts[seg*P + j]  float4  Hilbert-sorted point j of the segment, .w = original index bits
//   nrm[seg*P + j]  float4  (nx, ny, nz, curvature) of the same sorted point
//   nodes[seg*nodes_per_seg + level_off[l] + i]  wide node i of level l >= 1 of the implicit 8-ary
//       hierarchy over the sorted points: the tight AABBs of its 8 children in SoA form (192 B);
//       level 0 = groups of LEAF consecutive points, whose boxes live in the level-1 records
// Rounding contract: distances and transforms use separately rounded fp32
// mul/add (no FMA) so that the CPU oracle computes bit-identical values.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// developer build (make EXTRA=-DICP_DEBUG_ASSERT): device-side bounds assertions on every index the
// traversals derive from data; the memory-checking tools are not available on the GPU pool
#ifdef ICP_DEBUG_ASSERT
#include <cassert>
#define ICP_ASSERT(c) assert(c)
#else
#define ICP_ASSERT(c) ((void)0)
#endif

#define ICP_LOG_LEAF 3
#define ICP_LEAF (1 << ICP_LOG_LEAF)     // points per leaf group (8 * 16 B = one 128 B line)
#define ICP_LOG_FAN 3
#define ICP_FAN (1 << ICP_LOG_FAN)       // children per node
#define ICP_MAX_LEVELS 9                 // level 0 + 8 internal levels (8 pending-mask bytes): 8^9 points
#define ICP_NSUMS 37
#define ICP_PART_STRIDE 40               // doubles per partial-sum row (37 padded)
#define ICP_INVALID_ID 0x7fffffff
#define ICP_MAX_PEERS 8                  // one NVSwitch domain

#include "parts.h"   // the canonical partition of a pair's queries into tiles and parts (plain C++, also compiled by the host self-test)

// exchange buffer of the NVLink peer mode (one per rank, mapped by every peer): the owner of part o
// stores the part's sums of exchange `seq` into sums[seq & 1][o] of EVERY rank's buffer and then
// publishes seq in flag[seq & 1][its rank]
struct __align__(128) PeerFlag {
    unsigned long long seq;
    unsigned long long pad[15];
};
struct __align__(128) PeerBuf {
    double sums[2][ICP_NPARTS][ICP_PART_STRIDE];
    PeerFlag flag[2][ICP_MAX_PEERS];
};

// control words of the device-side loop (one per context)
struct LoopCtrl {
    unsigned int ticket;      // blocks of K2 that have arrived in this iteration
    unsigned int done_count;  // pairs that are finished (now or earlier) in this iteration
    unsigned long long seq;   // number of peer exchanges done so far (identical on every rank)
    int comm_error;           // a peer did not answer in time
    unsigned int work;        // next tile ticket of the running K1 launch (rewound by K2)
};

// boxes of the 8 children of one internal node, structure-of-arrays so that one node is 12
// independent 16-byte loads and the 8 bounds are computed together
struct __align__(16) WideNode {
    float lox[ICP_FAN], loy[ICP_FAN], loz[ICP_FAN];
    float hix[ICP_FAN], hiy[ICP_FAN], hiz[ICP_FAN];
};

// shape of the implicit hierarchy, identical for every segment of a cloud set
struct TreeShape {
    uint32_t pitch;                       // P: points per segment (multiple of ICP_LEAF)
    uint32_t n_levels;                    // levels above and including the leaf-group level; root = n_levels-1
    uint32_t level_off[ICP_MAX_LEVELS];   // offset of level l (>= 1) inside a segment's record array
    uint32_t level_cnt[ICP_MAX_LEVELS];   // padded node count of level l (multiple of ICP_FAN, root = 1)
    uint32_t nodes_per_seg;               // records per segment = sum of level_cnt[1..]
    uint32_t units_per_seg;               // 32-point units per segment = ceil(pitch / 32)
};

// device view of a cloud set
struct CloudView {
    const float4* pts;
    const float4* nrm;
    const WideNode* nodes;
    const uint32_t* pos_of_orig;          // seg*P + original index -> sorted position
    const uint32_t* unit_split;           // seg*units_per_seg + unit: lane mask of the second cluster of a split unit, else 0
    const uint32_t* counts;               // points per segment
    TreeShape shape;
    uint32_t nseg;
};

// per-pair loop state, lives in device memory for the whole registration
struct SegState {
    double T[12];        // accumulated source->target transform, 3x4 row-major, fp64
    float Tf[12];        // fp32 copy actually applied to the points
    double origin[3];    // fixed origin of the raw sums (target bbox centre)
    double diff, prev_diff, sumd2, min_ratio;
    unsigned long long pairs;
    int iter;            // solves applied
    int searches;        // searches run
    int done;            // 0 = running
    int converged;       // 1 diff<=thr, 2 rel tol, 0 max iters
    int degenerate;
    int track;           // walks of the next search record margins (the last step was small, see solve_kernel)
    int have_margins;    // the search that ran last recorded margins
    int probe;           // the next search starts with one exact probe per 256-query tile (the last step was large)
    double radius;       // half diagonal of the target's bounding box
    double spacing;      // typical point spacing of the target taken as a surface: extent / sqrt(points)
};

struct LoopParams {
    int max_iters;
    float diff_threshold;
    double rel_diff_tol;
    float max_pair_d2;   // +inf = off
    float min_normal_dot;  // -2 = off: reject pairs whose normals disagree (n_p . n_q below this)
    int identity_pairs;
    int warm_start;
    int trace;
    int skip_walk;       // 1 = skip the walk of a unit whose winners provably stand (margin test in K1)
    int probe;           // 1 = searches after a long step start with one exact probe per tile
};

// ---- pinned fp32 arithmetic -------------------------------------------------------------
__device__ __forceinline__ float dist2_rn(float px, float py, float pz, float qx, float qy, float qz) {
    float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2) ----------------------------------------
// Two independent round-to-nearest operations per instruction: bit-identical to the scalar
// forms, half the issue slots.  Only add and mul are used and the sums of squares stay scalar:
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it does not do that to scalar .rn
// code), which would break the separately-rounded contract of the distances and bounds.
struct f32x2 {
    unsigned long long v;
};
__device__ __forceinline__ f32x2 pack2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 p, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v));
}
__device__ __forceinline__ f32x2 add2_rn(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f32x2 mul2_rn(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}

// negated query, packed once per query: (-qx,-qx), (-qy,-qy), (-qz,-qz) for box pairs and
// (-qx,-qy) for points.  c + (-q) = -(q - c) exactly, and only its square is used.
struct NegQuery {
    f32x2 xx, yy, zz, xy;
    float nz;
};
__device__ __forceinline__ NegQuery make_neg_query(float qx, float qy, float qz) {
    NegQuery n;
    n.xx = pack2(-qx, -qx); n.yy = pack2(-qy, -qy); n.zz = pack2(-qz, -qz);
    n.xy = pack2(-qx, -qy);
    n.nz = -qz;
    return n;
}

// dist2_rn(q, c) with the x,y lanes packed: same value bit for bit
__device__ __forceinline__ float dist2_packed(const NegQuery& nq, const float4& c) {
    float sx, sy;
    f32x2 d = add2_rn(pack2(c.x, c.y), nq.xy);
    unpack2(mul2_rn(d, d), sx, sy);
    const float dz = __fadd_rn(c.z, nq.nz);
    return __fadd_rn(__fadd_rn(sx, sy), __fmul_rn(dz, dz));
}

// bounds of two boxes (children c, c+1 of one record) against one query, packed
__device__ __forceinline__ void box_bounds2(float lx0, float lx1, float ly0, float ly1, float lz0, float lz1, float hx0,
                                            float hx1, float hy0, float hy1, float hz0, float hz1, const NegQuery& nq,
                                            uint32_t& b0, uint32_t& b1) {
    float ax0, ax1, ay0, ay1, az0, az1, cx0, cx1, cy0, cy1, cz0, cz1;
    unpack2(add2_rn(pack2(lx0, lx1), nq.xx), ax0, ax1);   // lo - q
    unpack2(add2_rn(pack2(ly0, ly1), nq.yy), ay0, ay1);
    unpack2(add2_rn(pack2(lz0, lz1), nq.zz), az0, az1);
    unpack2(add2_rn(pack2(hx0, hx1), nq.xx), cx0, cx1);   // hi - q = -(q - hi)
    unpack2(add2_rn(pack2(hy0, hy1), nq.yy), cy0, cy1);
    unpack2(add2_rn(pack2(hz0, hz1), nq.zz), cz0, cz1);
    const float ex0 = fmaxf(fmaxf(ax0, -cx0), 0.f), ex1 = fmaxf(fmaxf(ax1, -cx1), 0.f);
    const float ey0 = fmaxf(fmaxf(ay0, -cy0), 0.f), ey1 = fmaxf(fmaxf(ay1, -cy1), 0.f);
    const float ez0 = fmaxf(fmaxf(az0, -cz0), 0.f), ez1 = fmaxf(fmaxf(az1, -cz1), 0.f);
    float sx0, sx1, sy0, sy1, sz0, sz1;
    unpack2(mul2_rn(pack2(ex0, ex1), pack2(ex0, ex1)), sx0, sx1);
    unpack2(mul2_rn(pack2(ey0, ey1), pack2(ey0, ey1)), sy0, sy1);
    unpack2(mul2_rn(pack2(ez0, ez1), pack2(ez0, ez1)), sz0, sz1);
    b0 = __float_as_uint(__fadd_rn(__fadd_rn(sx0, sy0), sz0));
    b1 = __float_as_uint(__fadd_rn(__fadd_rn(sx1, sy1), sz1));
}

// x' = (((T0*x)+(T1*y))+(T2*z))+T3, no FMA.  T is 3x4 row-major.
__device__ __forceinline__ float3 apply_rn(const float* __restrict__ T, float x, float y, float z) {
    float3 r;
    r.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), __fmul_rn(T[2], z)), T[3]);
    r.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], x), __fmul_rn(T[5], y)), __fmul_rn(T[6], z)), T[7]);
    r.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], x), __fmul_rn(T[9], y)), __fmul_rn(T[10], z)), T[11]);
    return r;
}
__device__ __forceinline__ float3 rotate_rn(const float* __restrict__ T, float x, float y, float z) {
    float3 r;
    r.x = __fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), __fmul_rn(T[2], z));
    r.y = __fadd_rn(__fadd_rn(__fmul_rn(T[4], x), __fmul_rn(T[5], y)), __fmul_rn(T[6], z));
    r.z = __fadd_rn(__fadd_rn(__fmul_rn(T[8], x), __fmul_rn(T[9], y)), __fmul_rn(T[10], z));
    return r;
}

// order-preserving float <-> uint map for atomicMin/Max on floats
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// Two consecutive float4 (32-byte aligned) with ONE 256-bit read-only load (sm_100: LDG.E.ENL2.256): a leaf group is
// four loads instead of eight, a node record six instead of twelve -- the walks are issue-bound, and every load of a
// warp-uniform address also costs the two moves that bring the address out of the uniform registers.
struct float4x2 {
    float4 a, b;
};
__device__ __forceinline__ float4x2 ldg8(const float4* p) {
    float4x2 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
        : "l"(p));
    return r;
}

#define ICP_CUDA_CHECK(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return icp_fail(ctx, _e, #expr, __FILE__, __LINE__); \
    } while (0)
