// index_build.cu -- upload, space-filling-curve sort and implicit box hierarchy of a cloud set.
//
// Replaces pasteInMatrix (ICP/func.cpp:5-15: AoS cloud -> dense matrices) and the kd-tree
// construction hidden inside pcl::search::KdTree (ICP/myicp.cpp:158-159).  One-off per cloud;
// the hot loop only reads the result.  Steps, all on the context stream:
//   bbox reduce -> 30-bit Hilbert key per point (uniform cell size) -> radix sort (CUB, keys are
//   segment<<30 | key so a whole batch sorts in one pass) -> gather into the float4 layout
//   with the original index in .w -> k-d partition of every 8192-point window (searched set only)
//   -> k-d refinement of every 512-point window -> inverse permutation -> split-unit masks ->
//   tight leaf boxes written into their parents' wide-node records -> one kernel per upper level
//   (union of a node's 8 children into its parent's record).
#include <cfloat>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>

#include "internal.h"

namespace {

// slot j of the pitched layout -> (segment, position).  Almost every cloud set has fewer than 2^32 slots, and a
// 64-bit division costs a hundred instructions per thread in kernels that otherwise do a handful of loads.
__device__ __forceinline__ void split_slot(size_t j, uint32_t pitch, uint32_t& seg, uint32_t& i) {
    if (j <= 0xffffffffull) {
        const uint32_t j32 = (uint32_t)j;
        seg = j32 / pitch; i = j32 - seg * pitch;
    } else {
        seg = (uint32_t)(j / pitch); i = (uint32_t)(j - (size_t)seg * pitch);
    }
}

__global__ void bbox_init_kernel(uint32_t* bbox, uint32_t nseg, int* flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nseg * 8) bbox[i] = ((i & 7) < 4) ? 0xffffffffu : 0u;
    if (i == 0) flags[0] = 0;
}

// grid: (blocks per segment, nseg)
__global__ void __launch_bounds__(256) bbox_kernel(const char* __restrict__ raw, size_t stride_bytes,
                                                   const uint32_t* __restrict__ counts, uint32_t max_count,
                                                   uint32_t* __restrict__ bbox, int* __restrict__ flags) {
    const uint32_t seg = blockIdx.y;
    const uint32_t cnt = counts[seg];
    uint32_t lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    int bad = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const float* p = reinterpret_cast<const float*>(raw + ((size_t)seg * max_count + i) * stride_bytes);
        float v[3] = {p[0], p[1], p[2]};
#pragma unroll
        for (int d = 0; d < 3; d++) {
            if (!isfinite(v[d])) bad = 1;
            uint32_t o = float_to_ordered(v[d]);
            lo[d] = min(lo[d], o);
            hi[d] = max(hi[d], o);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        lo[d] = __reduce_min_sync(0xffffffffu, lo[d]);
        hi[d] = __reduce_max_sync(0xffffffffu, hi[d]);
    }
    bad = __any_sync(0xffffffffu, bad);
    // one set of global atomics per BLOCK: all blocks of a segment hit the same six words, and with one set per warp
    // the kernel spent most of its time queueing at them
    __shared__ uint32_t s_lo[8][3], s_hi[8][3];
    __shared__ int s_bad[8];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 3; d++) { s_lo[warp][d] = lo[d]; s_hi[warp][d] = hi[d]; }
        s_bad[warp] = bad;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        uint32_t l = 0xffffffffu, h = 0u;
#pragma unroll
        for (int w = 0; w < 8; w++) { l = min(l, s_lo[w][threadIdx.x]); h = max(h, s_hi[w][threadIdx.x]); }
        atomicMin(&bbox[seg * 8 + threadIdx.x], l);
        atomicMax(&bbox[seg * 8 + 4 + threadIdx.x], h);
    }
    if (threadIdx.x == 3) {
        int b = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) b |= s_bad[w];
        if (b) atomicAdd(flags, 1);
    }
}

__global__ void bbox_finish_kernel(const uint32_t* __restrict__ bbox, const uint32_t* __restrict__ counts,
                                   float4* __restrict__ bb_lo, float4* __restrict__ bb_hi, uint32_t nseg) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    if (counts[s] == 0) {
        bb_lo[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        bb_hi[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    bb_lo[s] = make_float4(ordered_to_float(bbox[s * 8]), ordered_to_float(bbox[s * 8 + 1]),
                           ordered_to_float(bbox[s * 8 + 2]), 0.f);
    bb_hi[s] = make_float4(ordered_to_float(bbox[s * 8 + 4]), ordered_to_float(bbox[s * 8 + 5]),
                           ordered_to_float(bbox[s * 8 + 6]), 0.f);
}

__device__ __forceinline__ uint32_t spread10(uint32_t v) {  // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// Hilbert index (3*bits wide) of a cell coordinate of `bits` bits per axis (Skilling's transpose
// algorithm).  Hilbert order keeps every run of consecutive points spatially compact, so the boxes of
// the implicit 8-ary hierarchy overlap far less than under Z-order (measured: ~1.5x fewer node visits
// per query).
__device__ __forceinline__ uint32_t hilbert_index(uint32_t x, uint32_t y, uint32_t z, int bits) {
    uint32_t X[3] = {x, y, z};
    // branch-free (the two cases of every step differ per thread: as branches they serialise the warp)
    for (uint32_t Q = 1u << (bits - 1); Q > 1u; Q >>= 1) {
        const uint32_t Pm = Q - 1u;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const uint32_t set = (X[i] & Q) ? 0xffffffffu : 0u;          // bit set: invert the low bits of X[0]
            const uint32_t t = (X[0] ^ X[i]) & Pm & ~set;                // bit clear: exchange the low bits of X[0] and X[i]
            X[0] ^= (Pm & set) | t;
            X[i] ^= t;
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = 1u << (bits - 1); Q > 1u; Q >>= 1) t ^= (X[2] & Q) ? Q - 1u : 0u;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (spread10(X[0]) << 2) | (spread10(X[1]) << 1) | spread10(X[2]);
}

// one thread per slot of the pitched layout
template <typename KeyT>   // 32-bit keys when segment bits + curve bits fit (every single pair): the sort moves a third less
__global__ void __launch_bounds__(256) curve_key_kernel(const char* __restrict__ raw, size_t stride_bytes,
                                                     const uint32_t* __restrict__ counts, uint32_t max_count,
                                                     uint32_t pitch, uint32_t nseg, const float4* __restrict__ bb_lo,
                                                     const float4* __restrict__ bb_hi, int bits,
                                                     KeyT* __restrict__ keys, uint32_t* __restrict__ vals) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (size_t)nseg * pitch) return;
    uint32_t seg, i;
    split_slot(j, pitch, seg, i);
    // padding takes the largest index: the sort is stable and padding slots follow the real ones in the
    // input, so they still end up last in their segment
    const uint32_t cells = 1u << bits;
    uint32_t code = (1u << (3 * bits)) - 1u;
    if (i < counts[seg]) {
        const float* p = reinterpret_cast<const float*>(raw + ((size_t)seg * max_count + i) * stride_bytes);
        float4 lo = bb_lo[seg], hi = bb_hi[seg];
        float ext = fmaxf(fmaxf(hi.x - lo.x, hi.y - lo.y), hi.z - lo.z);
        float scale = ext > 0.f ? ((float)cells - 0.5f) / ext : 0.f;
        uint32_t cx = min(cells - 1u, (uint32_t)((p[0] - lo.x) * scale));
        uint32_t cy = min(cells - 1u, (uint32_t)((p[1] - lo.y) * scale));
        uint32_t cz = min(cells - 1u, (uint32_t)((p[2] - lo.z) * scale));
        code = hilbert_index(cx, cy, cz, bits);
    }
    keys[j] = (KeyT)(((unsigned long long)seg << (3 * bits)) | code);
    vals[j] = i;
}

__global__ void __launch_bounds__(256) gather_kernel(const char* __restrict__ raw, size_t stride_bytes,
                                                     const uint32_t* __restrict__ counts, uint32_t max_count,
                                                     uint32_t pitch, uint32_t nseg, const uint32_t* __restrict__ vals,
                                                     float4* __restrict__ pts) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (size_t)nseg * pitch) return;
    uint32_t seg, i_unused;
    split_slot(j, pitch, seg, i_unused);
    uint32_t v = vals[j];
    if (v < counts[seg]) {
        const float* p = reinterpret_cast<const float*>(raw + ((size_t)seg * max_count + v) * stride_bytes);
        pts[j] = make_float4(p[0], p[1], p[2], __uint_as_float(v));
    } else {
        pts[j] = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, __uint_as_float((uint32_t)ICP_INVALID_ID));
    }
}

// every record starts as 8 empty boxes (lo=+inf, hi=-inf): padding nodes stay that way
__global__ void __launch_bounds__(256) init_nodes_kernel(WideNode* __restrict__ nodes, size_t n_floats) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_floats) return;
    const float inf = __int_as_float(0x7f800000);
    const uint32_t r = t <= 0xffffffffull ? (uint32_t)t % 48u : (uint32_t)(t % 48);
    reinterpret_cast<float*>(nodes)[t] = (r < 24) ? inf : -inf;
}

// ---- local k-d refinement of the curve order ----------------------------------------------
// Inside every window of 512 consecutive (Hilbert-sorted) points the points are re-ordered by six
// levels of median splits along the longest axis of each part (512 -> 256 -> ... -> 8).  Three
// binary splits are one level of the 8-ary hierarchy, so the leaf groups (8 points), the level-1
// nodes (64) and the level-2 nodes (512) of a window become k-d cells: disjoint and tighter than
// runs of the curve.  Measured on the 1M target: 20 % fewer leaf groups and 10 % fewer nodes reach a
// warp of queries.  Results do not depend on the order (ties go by original index, carried in .w).
template <uint32_t kKdWindow>   // 256 .. 2048 points per window; kKdWindow / 2 threads, thread t owns elements t and t + kKdHalf
__global__ void __launch_bounds__(kKdWindow / 2) kd_refine_kernel(float4* __restrict__ pts, const uint32_t* __restrict__ counts,
                                                                 uint32_t pitch, uint32_t wins_per_seg) {
    constexpr uint32_t kKdHalf = kKdWindow / 2, kKdIdxMask = kKdWindow - 1;
    __shared__ float4 sp[kKdWindow];
    __shared__ __align__(16) uint32_t skey[kKdWindow];             // ordered coordinate, low 9 bits replaced by the element index
    __shared__ __align__(16) uint32_t sbox[kKdWindow / ICP_LEAF][6];                 // per part: ordered-uint lo xyz, hi xyz
    // split levels with parts of 64 points or more only PARTITION at the median (radix select + stable
    // partition, as in kd_coarse_kernel below); the sort is kept for the last three levels (32, 16 -> 8)
    // (explicit alignment: without it ptxas merged neighbouring 4-byte shared accesses into misaligned 8-byte ones)
    constexpr uint32_t kMaxParts = kKdWindow / 64, kCells = kKdWindow / 32, kCellsPerLane = kCells >= 32 ? kCells / 32 : 1;
    static_assert(kKdWindow <= 1024, "histograms of larger windows do not fit the static shared memory");
    __shared__ __align__(16) uint32_t shist[kMaxParts][256];
    __shared__ __align__(16) uint32_t s_prefix[kMaxParts], s_rank[kMaxParts], s_klo[kMaxParts], s_khi[kMaxParts], s_axis[kMaxParts];
    __shared__ __align__(16) uint32_t s_cnt[kCells][2];
    const uint32_t seg = blockIdx.x / wins_per_seg, win = blockIdx.x - seg * wins_per_seg;
    const uint32_t base = win * kKdWindow;
    if (base + kKdWindow > counts[seg]) return;      // partial window (or padding): keep the curve order
    for (uint32_t v = threadIdx.x; v < kMaxParts * 256; v += kKdHalf) shist[v >> 8][v & 255] = 0u;
    float4* gp = pts + (size_t)seg * pitch + base;
    const uint32_t t = threadIdx.x;                  // kKdHalf threads, elements t and t + kKdHalf
    sp[t] = gp[t];
    sp[t + kKdHalf] = gp[t + kKdHalf];
    __syncthreads();
    for (uint32_t S = kKdWindow; S > ICP_LEAF; S >>= 1) {       // part size at this split level
        const uint32_t parts = kKdWindow / S;
        for (uint32_t v = t; v < parts * 6; v += kKdHalf) sbox[v / 6][v % 6] = (v % 6 < 3) ? 0xffffffffu : 0u;
        __syncthreads();
        // bounding box of every part: the 32 elements of a warp lie in one part while S >= 32
        // (one REDUX per component, one shared atomic per warp); below that, 16 elements per part
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t e = t + kKdHalf * h, part = e / S;
            const float4 p = sp[e];
            uint32_t ox = float_to_ordered(p.x), oy = float_to_ordered(p.y), oz = float_to_ordered(p.z);
            if (S >= 32) {
                const uint32_t lx = __reduce_min_sync(0xffffffffu, ox), hx = __reduce_max_sync(0xffffffffu, ox);
                const uint32_t ly = __reduce_min_sync(0xffffffffu, oy), hy = __reduce_max_sync(0xffffffffu, oy);
                const uint32_t lz = __reduce_min_sync(0xffffffffu, oz), hz = __reduce_max_sync(0xffffffffu, oz);
                if ((t & 31) == 0) {
                    atomicMin(&sbox[part][0], lx); atomicMax(&sbox[part][3], hx);
                    atomicMin(&sbox[part][1], ly); atomicMax(&sbox[part][4], hy);
                    atomicMin(&sbox[part][2], lz); atomicMax(&sbox[part][5], hz);
                }
            } else {
                atomicMin(&sbox[part][0], ox); atomicMax(&sbox[part][3], ox);
                atomicMin(&sbox[part][1], oy); atomicMax(&sbox[part][4], oy);
                atomicMin(&sbox[part][2], oz); atomicMax(&sbox[part][5], oz);
            }
        }
        __syncthreads();
        if (S >= 64) {
            const uint32_t lane = t & 31, warp = t >> 5;
            if (t < parts) {
                const float ex = ordered_to_float(sbox[t][3]) - ordered_to_float(sbox[t][0]);
                const float ey = ordered_to_float(sbox[t][4]) - ordered_to_float(sbox[t][1]);
                const float ez = ordered_to_float(sbox[t][5]) - ordered_to_float(sbox[t][2]);
                const uint32_t ax = (ex >= ey && ex >= ez) ? 0u : (ey >= ez ? 1u : 2u);
                s_axis[t] = ax; s_klo[t] = sbox[t][ax]; s_khi[t] = sbox[t][3 + ax];
                s_prefix[t] = 0u; s_rank[t] = S / 2 - 1;
            }
            __syncthreads();
            uint32_t key[2], prt[2];
            int nbits[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t e = t + kKdHalf * h;
                prt[h] = e / S;
                const float4 p = sp[e];
                const uint32_t ax = s_axis[prt[h]];
                key[h] = float_to_ordered(ax == 0u ? p.x : (ax == 1u ? p.y : p.z)) - s_klo[prt[h]];
                nbits[h] = 32 - __clz((int)(s_khi[prt[h]] - s_klo[prt[h]]));
            }
            for (int r = 0; r < 4; r++) {   // radix select of the median key, on the part's key range (see kd_coarse_kernel)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int remaining = nbits[h] - 8 * r;
                    if (remaining > 0) {
                        const int w = remaining < 8 ? remaining : 8, shift = remaining - w;
                        const bool in = remaining >= 32 || (key[h] >> remaining) == (s_prefix[prt[h]] >> remaining);
                        if (in) atomicAdd(&shist[prt[h]][(key[h] >> shift) & ((1u << w) - 1u)], 1u);
                    }
                }
                __syncthreads();
                if (warp < parts) {
                    const int remaining = (32 - __clz((int)(s_khi[warp] - s_klo[warp]))) - 8 * r;
                    if (remaining > 0) {
                        const int w = remaining < 8 ? remaining : 8, shift = remaining - w;
                        uint32_t c[8], sum = 0;
#pragma unroll
                        for (int b = 0; b < 8; b++) { c[b] = shist[warp][lane * 8 + b]; sum += c[b]; shist[warp][lane * 8 + b] = 0u; }
                        uint32_t incl = sum;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += v; }
                        const uint32_t excl = incl - sum, rank = s_rank[warp];
                        if (rank >= excl && rank < incl) {
                            uint32_t acc = excl;
                            int d = 0;
                            bool found = false;
#pragma unroll
                            for (int b = 0; b < 8; b++) {
                                if (!found) {
                                    if (rank >= acc + c[b]) { acc += c[b]; d = b + 1; }
                                    else found = true;
                                }
                            }
                            s_prefix[warp] |= (lane * 8u + (uint32_t)d) << shift;
                            s_rank[warp] = rank - acc;
                        }
                    }
                }
                __syncthreads();
            }
            // stable partition (ties at the median go low in the current order)
            uint32_t bl[2], be[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t m = s_prefix[prt[h]];
                bl[h] = __ballot_sync(0xffffffffu, key[h] < m);
                be[h] = __ballot_sync(0xffffffffu, key[h] == m);
                if (lane == 0) { const uint32_t cell = (t + kKdHalf * h) >> 5; s_cnt[cell][0] = __popc(bl[h]); s_cnt[cell][1] = __popc(be[h]); }
            }
            __syncthreads();
            if (warp == 0) {
                uint32_t cl[kCellsPerLane], ce[kCellsPerLane], suml = 0, sume = 0;
#pragma unroll
                for (uint32_t c = 0; c < kCellsPerLane; c++) {
                    const uint32_t cell = lane * kCellsPerLane + c;
                    cl[c] = cell < kCells ? s_cnt[cell][0] : 0u; ce[c] = cell < kCells ? s_cnt[cell][1] : 0u;
                    suml += cl[c]; sume += ce[c];
                }
                uint32_t il = suml, ie = sume;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t vl = __shfl_up_sync(0xffffffffu, il, o), ve = __shfl_up_sync(0xffffffffu, ie, o);
                    if ((int)lane >= o) { il += vl; ie += ve; }
                }
                uint32_t al = il - suml, ae = ie - sume;
#pragma unroll
                for (uint32_t c = 0; c < kCellsPerLane; c++) {
                    const uint32_t cell = lane * kCellsPerLane + c;
                    if (cell < kCells) { s_cnt[cell][0] = al; s_cnt[cell][1] = ae; }
                    al += cl[c]; ae += ce[c];
                }
            }
            __syncthreads();
            uint32_t dest[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t e = t + kKdHalf * h, cell = e >> 5, cell0 = (prt[h] * S) >> 5;
                const uint32_t m = s_prefix[prt[h]], need_eq = s_rank[prt[h]] + 1u;
                const uint32_t less_before = s_cnt[cell][0] - s_cnt[cell0][0] + __popc(bl[h] & ((1u << lane) - 1u));
                const uint32_t eq_before = s_cnt[cell][1] - s_cnt[cell0][1] + __popc(be[h] & ((1u << lane) - 1u));
                const bool less = key[h] < m, eq = key[h] == m;
                const uint32_t low_before = less_before + min(eq_before, need_eq);
                const bool goes_low = less || (eq && eq_before < need_eq);
                dest[h] = prt[h] * S + (goes_low ? low_before : S / 2 + ((e - prt[h] * S) - low_before));
#ifdef ICP_DEBUG_ASSERT
                if (dest[h] >= kKdWindow || (goes_low && low_before >= S / 2))
                    printf("kd_refine S=%u e=%u part=%u m=%u key=%u need_eq=%u less_before=%u eq_before=%u cell=%u cell0=%u cnt=%u,%u cnt0=%u,%u nbits=%d\n", S, e, prt[h], m, key[h], need_eq, less_before, eq_before, cell, cell0, s_cnt[cell][0], s_cnt[cell][1], s_cnt[cell0][0], s_cnt[cell0][1], nbits[h]);
                dest[h] &= kKdIdxMask;
#endif
            }
            const float4 a = sp[t], b = sp[t + kKdHalf];
            __syncthreads();
            sp[dest[0]] = a;
            sp[dest[1]] = b;
            __syncthreads();
            continue;
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t e = t + kKdHalf * h, part = e / S;
            const float ex = ordered_to_float(sbox[part][3]) - ordered_to_float(sbox[part][0]);
            const float ey = ordered_to_float(sbox[part][4]) - ordered_to_float(sbox[part][1]);
            const float ez = ordered_to_float(sbox[part][5]) - ordered_to_float(sbox[part][2]);
            const float4 p = sp[e];
            const float c = (ex >= ey && ex >= ez) ? p.x : (ey >= ez ? p.y : p.z);
            // 23 bits of the coordinate are plenty to find a median; the index makes the keys unique
            skey[e] = (float_to_ordered(c) & ~kKdIdxMask) | e;
        }
        __syncthreads();
        // bitonic sort inside every part (parts are aligned powers of two; keys are unique).  Thread t
        // keeps the keys of elements t and t + kKdHalf in registers; with parts of 32 and 16 points (the
        // only ones that are still sorted) every exchange is a warp shuffle.  The general form below
        // also handles larger parts: the thread's own pair at distance kKdHalf, shared memory from 32 up.
        uint32_t v0 = skey[t], v1 = skey[t + kKdHalf];
        for (uint32_t k = 2; k <= S; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                if (j == kKdHalf) {                   // k == kKdWindow == S: ascending, element t is the low one
                    const uint32_t lo = min(v0, v1), hi = max(v0, v1);
                    v0 = lo; v1 = hi;
                    continue;
                }
                uint32_t p0, p1;
                if (j >= 32) {
                    __syncthreads();
                    skey[t] = v0; skey[t + kKdHalf] = v1;
                    __syncthreads();
                    p0 = skey[t ^ j]; p1 = skey[(t ^ j) + kKdHalf];
                } else {
                    p0 = __shfl_xor_sync(0xffffffffu, v0, (int)j);
                    p1 = __shfl_xor_sync(0xffffffffu, v1, (int)j);
                }
                // element e keeps the smaller key when it is the low partner of an ascending pair or the
                // high partner of a descending one
                const bool is_lo = (t & j) == 0;
                const bool up0 = ((t & k) == 0) || (k == S), up1 = (((t + kKdHalf) & k) == 0) || (k == S);
                v0 = (is_lo == up0) ? min(v0, p0) : max(v0, p0);
                v1 = (is_lo == up1) ? min(v1, p1) : max(v1, p1);
            }
        }
        __syncthreads();
        skey[t] = v0; skey[t + kKdHalf] = v1;
        __syncthreads();
        // permute the payload
        const float4 a = sp[skey[t] & kKdIdxMask], b = sp[skey[t + kKdHalf] & kKdIdxMask];
        __syncthreads();
        sp[t] = a;
        sp[t + kKdHalf] = b;
        __syncthreads();
    }
    gp[t] = sp[t];
    gp[t + kKdHalf] = sp[t + kKdHalf];
}

// ---- coarse k-d splits of large windows ---------------------------------------------------------
// The refinement above sorts, which costs log^2 steps and is limited to what two elements per
// thread hold.  Above it only the PARTITION matters: every aligned window of W points is split at the
// median of its longest axis into two halves, the halves again, ... down to parts of kKdFine points
// (then kd_refine_kernel takes over).  A split level is an exact radix select of the median key in
// shared memory (4 rounds of 8 bits, one histogram per part) and a stable partition (block-wide prefix
// sums), so its cost does not grow with the part size.  Ties at the median go low in curve order.
// 1024 threads own W / 1024 elements each (element e = k * 1024 + t); parts are at least 1024 points,
// so the 1024 elements of one k all lie in the same part.
constexpr uint32_t kKdFine = 512;
template <uint32_t W>
__global__ void __launch_bounds__(1024) kd_coarse_kernel(float4* __restrict__ pts, const uint32_t* __restrict__ counts,
                                                         uint32_t pitch, uint32_t wins_per_seg) {
    constexpr uint32_t E = W / 1024, MAXP = W / (2 * kKdFine);
    extern __shared__ __align__(16) unsigned char kd_smem[];
    float4* sp = reinterpret_cast<float4*>(kd_smem);     // the window's points; keys live in registers
    __shared__ __align__(16) uint32_t sbox[MAXP][6];
    __shared__ __align__(16) uint32_t shist[MAXP][256];
    __shared__ __align__(16) uint32_t s_prefix[MAXP], s_rank[MAXP];       // radix select state per part
    __shared__ __align__(16) uint32_t s_klo[MAXP], s_khi[MAXP];           // smallest and largest key of the part
    __shared__ __align__(16) uint32_t s_cnt[E][32][2];                    // partition scan: (less, equal) counts per slab and warp, then their exclusive prefix
    const uint32_t seg = blockIdx.x / wins_per_seg, win = blockIdx.x - seg * wins_per_seg;
    const uint32_t base = win * W;
    if (base + W > counts[seg]) return;                      // partial window: left to the curve order + fine refinement
    float4* gp = pts + (size_t)seg * pitch + base;
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
#pragma unroll
    for (uint32_t k = 0; k < E; k++) sp[k * 1024 + t] = gp[k * 1024 + t];
    __syncthreads();
    for (uint32_t S = W, logS = 31 - __clz((int)W); S > kKdFine; S >>= 1, logS--) {
        const uint32_t parts = W / S;
        for (uint32_t v = t; v < parts * 6; v += 1024) sbox[v / 6][v % 6] = (v % 6 < 3) ? 0xffffffffu : 0u;
        if (t < parts) { s_prefix[t] = 0u; s_rank[t] = S / 2 - 1; }
        if (S == W) for (uint32_t v = t; v < MAXP * 256; v += 1024) shist[v >> 8][v & 255] = 0u;
        __syncthreads();
        // bounding box of every part
#pragma unroll
        for (uint32_t k = 0; k < E; k++) {
            const float4 p = sp[k * 1024 + t];
            const uint32_t part = (k << 10) >> logS;
            const uint32_t ox = float_to_ordered(p.x), oy = float_to_ordered(p.y), oz = float_to_ordered(p.z);
            const uint32_t lx = __reduce_min_sync(0xffffffffu, ox), hx = __reduce_max_sync(0xffffffffu, ox);
            const uint32_t ly = __reduce_min_sync(0xffffffffu, oy), hy = __reduce_max_sync(0xffffffffu, oy);
            const uint32_t lz = __reduce_min_sync(0xffffffffu, oz), hz = __reduce_max_sync(0xffffffffu, oz);
            if (lane == 0) {
                atomicMin(&sbox[part][0], lx); atomicMax(&sbox[part][3], hx);
                atomicMin(&sbox[part][1], ly); atomicMax(&sbox[part][4], hy);
                atomicMin(&sbox[part][2], lz); atomicMax(&sbox[part][5], hz);
            }
        }
        __syncthreads();
        // keys: the coordinate along the longest axis of the part
        if (t < parts) {
            const float ex = ordered_to_float(sbox[t][3]) - ordered_to_float(sbox[t][0]);
            const float ey = ordered_to_float(sbox[t][4]) - ordered_to_float(sbox[t][1]);
            const float ez = ordered_to_float(sbox[t][5]) - ordered_to_float(sbox[t][2]);
            const int ax = (ex >= ey && ex >= ez) ? 0 : (ey >= ez ? 1 : 2);
            s_klo[t] = sbox[t][ax]; s_khi[t] = sbox[t][3 + ax];    // float_to_ordered is monotone: these are the extreme keys
        }
        uint32_t key[E];
#pragma unroll
        for (uint32_t k = 0; k < E; k++) {
            const uint32_t part = (k << 10) >> logS;
            const float ex = ordered_to_float(sbox[part][3]) - ordered_to_float(sbox[part][0]);
            const float ey = ordered_to_float(sbox[part][4]) - ordered_to_float(sbox[part][1]);
            const float ez = ordered_to_float(sbox[part][5]) - ordered_to_float(sbox[part][2]);
            const float4 p = sp[k * 1024 + t];
            key[k] = float_to_ordered((ex >= ey && ex >= ez) ? p.x : (ey >= ez ? p.y : p.z));
        }
        __syncthreads();   // s_klo / s_khi
        // exact median key of every part: radix select, most significant digit first, on (key - smallest key
        // of the part).  Starting at the top bit of the part's key RANGE spreads the first histogram over all
        // of its 256 bins; the raw high bytes of neighbouring coordinates are nearly all equal and would
        // serialise thousands of shared-memory atomics on a handful of bins.
        int nbits[E];   // width of the part's key range
#pragma unroll
        for (uint32_t k = 0; k < E; k++) {
            const uint32_t part = (k << 10) >> logS;
            key[k] -= s_klo[part];
            nbits[k] = 32 - __clz((int)(s_khi[part] - s_klo[part]));
        }
        for (int r = 0; r < 4; r++) {   // (the histograms are zero here: cleared below by their readers)
#pragma unroll
            for (uint32_t k = 0; k < E; k++) {
                const uint32_t part = (k << 10) >> logS;
                const int remaining = nbits[k] - 8 * r;   // key bits not yet fixed
                if (remaining > 0) {
                    const int w = remaining < 8 ? remaining : 8, shift = remaining - w;
                    const bool in = remaining >= 32 || (key[k] >> remaining) == (s_prefix[part] >> remaining);
                    if (in) atomicAdd(&shist[part][(key[k] >> shift) & ((1u << w) - 1u)], 1u);
                }
            }
            __syncthreads();
            if (warp < parts) {   // one warp per part: find the digit in which the cumulative count passes the rank
                const int remaining = (32 - __clz((int)(s_khi[warp] - s_klo[warp]))) - 8 * r;
                if (remaining > 0) {
                    const int w = remaining < 8 ? remaining : 8, shift = remaining - w;
                    uint32_t c[8], sum = 0;
#pragma unroll
                    for (int b = 0; b < 8; b++) { c[b] = shist[warp][lane * 8 + b]; sum += c[b]; shist[warp][lane * 8 + b] = 0u; }
                    uint32_t incl = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += v; }
                    const uint32_t excl = incl - sum, rank = s_rank[warp];
                    if (rank >= excl && rank < incl) {          // exactly one lane
                        uint32_t acc = excl;
                        int d = 0;
                        bool found = false;   // d stops at the first digit whose cumulative count exceeds the rank
#pragma unroll
                        for (int b = 0; b < 8; b++) {
                            if (!found) {
                                if (rank >= acc + c[b]) { acc += c[b]; d = b + 1; }
                                else found = true;
                            }
                        }
                        s_prefix[warp] |= (lane * 8u + (uint32_t)d) << shift;
                        s_rank[warp] = rank - acc;
                    }
                }
            }
            __syncthreads();
        }
        // stable partition: elements below the median key go low, the first (rank + 1) of the equal ones too.
        // Counts of (less, equal) per warp and slab k -> one exclusive scan over the E x 32 (slab, warp) cells in
        // element order, restarted at every part boundary, done by warp 0.
        uint32_t bl[E], be[E];
#pragma unroll
        for (uint32_t k = 0; k < E; k++) {
            const uint32_t m = s_prefix[(k << 10) >> logS];
            bl[k] = __ballot_sync(0xffffffffu, key[k] < m);
            be[k] = __ballot_sync(0xffffffffu, key[k] == m);
            if (lane == 0) { s_cnt[k][warp][0] = __popc(bl[k]); s_cnt[k][warp][1] = __popc(be[k]); }
        }
        __syncthreads();
        if (warp == 0) {
            // element order is slab-major: cell (k, w) precedes (k, w+1) and (k+1, 0); lane l takes the E
            // consecutive cells l*E .. l*E+E-1
            uint32_t cl[E], ce[E], suml = 0, sume = 0;
#pragma unroll
            for (uint32_t c = 0; c < E; c++) {
                const uint32_t cell = lane * E + c, k = cell >> 5, w = cell & 31u;
                cl[c] = s_cnt[k][w][0]; ce[c] = s_cnt[k][w][1];
                suml += cl[c]; sume += ce[c];
            }
            uint32_t il = suml, ie = sume;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t vl = __shfl_up_sync(0xffffffffu, il, o), ve = __shfl_up_sync(0xffffffffu, ie, o);
                if ((int)lane >= o) { il += vl; ie += ve; }
            }
            uint32_t al = il - suml, ae = ie - sume;     // exclusive over the whole window, element order
#pragma unroll
            for (uint32_t c = 0; c < E; c++) {
                const uint32_t cell = lane * E + c, k = cell >> 5, w = cell & 31u;
                s_cnt[k][w][0] = al; s_cnt[k][w][1] = ae;
                al += cl[c]; ae += ce[c];
            }
        }
        __syncthreads();
        uint32_t dest[E];
#pragma unroll
        for (uint32_t k = 0; k < E; k++) {
            const uint32_t part = (k << 10) >> logS, m = s_prefix[part], need_eq = s_rank[part] + 1u;
            // window-wide exclusive counts minus those at the start of the part = counts inside the part
            const uint32_t k0 = (part * S) >> 10;
            const uint32_t less_before = s_cnt[k][warp][0] - s_cnt[k0][0][0] + __popc(bl[k] & ((1u << lane) - 1u));
            const uint32_t eq_before = s_cnt[k][warp][1] - s_cnt[k0][0][1] + __popc(be[k] & ((1u << lane) - 1u));
            const bool less = key[k] < m, eq = key[k] == m;
            const uint32_t e_local = (k * 1024 + t) - part * S;
            const uint32_t low_before = less_before + min(eq_before, need_eq);
            const bool goes_low = less || (eq && eq_before < need_eq);
            dest[k] = part * S + (goes_low ? low_before : S / 2 + (e_local - low_before));
        }
        float4 mine[E];
#pragma unroll
        for (uint32_t k = 0; k < E; k++) mine[k] = sp[k * 1024 + t];
        __syncthreads();
#pragma unroll
        for (uint32_t k = 0; k < E; k++) sp[dest[k]] = mine[k];
        __syncthreads();
    }
#pragma unroll
    for (uint32_t k = 0; k < E; k++) gp[k * 1024 + t] = sp[k * 1024 + t];
}

// ---- split units ------------------------------------------------------------------------------
// A unit is 32 consecutive sorted points: the 32 queries one warp searches together.  The curve leaves
// and re-enters a surface, so about one unit in a hundred holds two clusters that lie far apart; its
// warp would pay for the union of two unrelated search regions with half of its lanes idle in each.
// Such a unit is marked here (once per cloud) with the lane mask of its second cluster and the kNN search
// of the normal estimation walks the hierarchy once per cluster (normals of 1M points: -6 %, and no more
// straggler blocks at the end of the kernel).  Split = cut at the middle of the longest axis leaves two
// groups that are each at most 0.4x as wide as the unit (2-4 % of the units of a scanned surface).
#ifndef ICP_SPLIT_RATIO2
#define ICP_SPLIT_RATIO2 6.f    // (unit width / group width)^2 from which a unit counts as two clusters (swept: 4..9 equal, 2 and 16 worse)
#endif
__device__ __forceinline__ void store_child_box(WideNode* __restrict__ rec, uint32_t slot, float3 lo, float3 hi) {
    rec->lox[slot] = lo.x; rec->loy[slot] = lo.y; rec->loz[slot] = lo.z;
    rec->hix[slot] = hi.x; rec->hiy[slot] = hi.y; rec->hiz[slot] = hi.z;
}

// level >= 1 (below the root): union of the node's own 8 child boxes goes into its parent's record
__global__ void __launch_bounds__(256) upper_box_kernel(WideNode* __restrict__ nodes, TreeShape sh, uint32_t nseg, int level,
                                                        uint32_t level_off, uint32_t level_cnt, uint32_t parent_off) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nseg * level_cnt) return;
    uint32_t seg, i;
    split_slot(t, level_cnt, seg, i);
    WideNode* base = nodes + (size_t)seg * sh.nodes_per_seg;
    const WideNode* me = base + level_off + i;
    const float inf = __int_as_float(0x7f800000);
    float3 lo = make_float3(inf, inf, inf), hi = make_float3(-inf, -inf, -inf);
#pragma unroll
    for (int k = 0; k < ICP_FAN; k++) {
        lo.x = fminf(lo.x, me->lox[k]); lo.y = fminf(lo.y, me->loy[k]); lo.z = fminf(lo.z, me->loz[k]);
        hi.x = fmaxf(hi.x, me->hix[k]); hi.y = fmaxf(hi.y, me->hiy[k]); hi.z = fmaxf(hi.z, me->hiz[k]);
    }
    store_child_box(base + parent_off + (i >> ICP_LOG_FAN), i & (ICP_FAN - 1), lo, hi);
}

// One warp per 32-point unit, one pass over the final order: the inverse permutation
// (pos_of_orig), the split mask of the unit (rule above) and the tight boxes of its four leaf groups,
// written into slot g&7 of record (1, g>>3).
__global__ void __launch_bounds__(256) finish_units_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ counts,
                                                           TreeShape sh, uint32_t nseg, uint32_t* __restrict__ pos_of_orig,
                                                           uint32_t* __restrict__ unit_split, WideNode* __restrict__ nodes) {
    const size_t wid = ((size_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (wid >= (size_t)nseg * sh.units_per_seg) return;
    uint32_t seg, unit;
    split_slot(wid, sh.units_per_seg, seg, unit);
    const uint32_t i = unit * 32u + lane;
    const bool in_pitch = i < sh.pitch;
    float4 p = make_float4(0.f, 0.f, 0.f, __uint_as_float((uint32_t)ICP_INVALID_ID));
    if (in_pitch) p = pts[(size_t)seg * sh.pitch + i];
    const bool act = in_pitch && i < counts[seg] && __float_as_uint(p.w) != (uint32_t)ICP_INVALID_ID;
    if (act) pos_of_orig[(size_t)seg * sh.pitch + __float_as_uint(p.w)] = i;
    // ---- split mask ----
    const uint32_t none_lo = 0xffffffffu, none_hi = 0u;
    const uint32_t ox = float_to_ordered(p.x), oy = float_to_ordered(p.y), oz = float_to_ordered(p.z);
    const float lx = ordered_to_float(__reduce_min_sync(0xffffffffu, act ? ox : none_lo)), hx = ordered_to_float(__reduce_max_sync(0xffffffffu, act ? ox : none_hi));
    const float ly = ordered_to_float(__reduce_min_sync(0xffffffffu, act ? oy : none_lo)), hy = ordered_to_float(__reduce_max_sync(0xffffffffu, act ? oy : none_hi));
    const float lz = ordered_to_float(__reduce_min_sync(0xffffffffu, act ? oz : none_lo)), hz = ordered_to_float(__reduce_max_sync(0xffffffffu, act ? oz : none_hi));
    uint32_t mask = 0u;
    if (__popc(__ballot_sync(0xffffffffu, act)) >= 2) {
        const float ex = hx - lx, ey = hy - ly, ez = hz - lz;
        const float c = (ex >= ey && ex >= ez) ? p.x : (ey >= ez ? p.y : p.z);
        const float mid = (ex >= ey && ex >= ez) ? 0.5f * (lx + hx) : (ey >= ez ? 0.5f * (ly + hy) : 0.5f * (lz + hz));
        const bool g1 = act && c > mid, g0 = act && !g1;
        float d2[2];
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const bool in = g ? g1 : g0;
            const float ax = ordered_to_float(__reduce_max_sync(0xffffffffu, in ? ox : none_hi)) - ordered_to_float(__reduce_min_sync(0xffffffffu, in ? ox : none_lo));
            const float ay = ordered_to_float(__reduce_max_sync(0xffffffffu, in ? oy : none_hi)) - ordered_to_float(__reduce_min_sync(0xffffffffu, in ? oy : none_lo));
            const float az = ordered_to_float(__reduce_max_sync(0xffffffffu, in ? oz : none_hi)) - ordered_to_float(__reduce_min_sync(0xffffffffu, in ? oz : none_lo));
            d2[g] = ax * ax + ay * ay + az * az;
        }
        const uint32_t m1 = __ballot_sync(0xffffffffu, g1), m0 = __ballot_sync(0xffffffffu, g0);
        const float whole = ex * ex + ey * ey + ez * ez;
        if (m0 && m1 && fmaxf(d2[0], d2[1]) * ICP_SPLIT_RATIO2 < whole) mask = m1;
    }
    if (lane == 0) unit_split[(size_t)seg * sh.units_per_seg + unit] = mask;
    // ---- leaf boxes: lanes 8j .. 8j+7 hold leaf group i / 8 ----
    if (!nodes) return;
    const float inf = __int_as_float(0x7f800000);
    float3 lo = act ? make_float3(p.x, p.y, p.z) : make_float3(inf, inf, inf);
    float3 hi = act ? make_float3(p.x, p.y, p.z) : make_float3(-inf, -inf, -inf);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o)); lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if ((lane & 7u) == 0u && in_pitch) {
        const uint32_t g = i >> ICP_LOG_LEAF;
        store_child_box(nodes + (size_t)seg * sh.nodes_per_seg + sh.level_off[1] + (g >> ICP_LOG_FAN), g & (ICP_FAN - 1), lo, hi);
    }
}

// levels first_level .. root-1 of one segment in one block (they hold a few thousand nodes at most): every
// node's own 8 child boxes are united into its parent's record, level by level
__global__ void __launch_bounds__(512) upper_levels_kernel(WideNode* __restrict__ nodes, TreeShape sh, uint32_t first_level) {
    WideNode* base = nodes + (size_t)blockIdx.x * sh.nodes_per_seg;
    const float inf = __int_as_float(0x7f800000);
    for (uint32_t l = first_level; l + 1 < sh.n_levels; l++) {
        for (uint32_t i = threadIdx.x; i < sh.level_cnt[l]; i += blockDim.x) {
            const WideNode* me = base + sh.level_off[l] + i;
            float3 lo = make_float3(inf, inf, inf), hi = make_float3(-inf, -inf, -inf);
#pragma unroll
            for (int k = 0; k < ICP_FAN; k++) {
                lo.x = fminf(lo.x, me->lox[k]); lo.y = fminf(lo.y, me->loy[k]); lo.z = fminf(lo.z, me->loz[k]);
                hi.x = fmaxf(hi.x, me->hix[k]); hi.y = fmaxf(hi.y, me->hiy[k]); hi.z = fmaxf(hi.z, me->hiz[k]);
            }
            store_child_box(base + sh.level_off[l + 1] + (i >> ICP_LOG_FAN), i & (ICP_FAN - 1), lo, hi);
        }
        __syncthreads();   // the next level reads what this one wrote (same block)
    }
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

}  // namespace

template <typename T>
int icp_reserve(icpsymm_ctx* ctx, T** p, size_t* cap, size_t need) {
    if (need <= *cap && *p) return ICPSYMM_OK;
    if (*p) { ICP_CUDA_CHECK(cudaFree(*p)); *p = nullptr; *cap = 0; }
    size_t n = need ? need : 1;
    ICP_CUDA_CHECK(cudaMalloc((void**)p, n * sizeof(T)));
    *cap = n;
    if (ctx) ctx->generation++;   // captured graphs hold the old pointer
    return ICPSYMM_OK;
}
template int icp_reserve<float4>(icpsymm_ctx*, float4**, size_t*, size_t);
template int icp_reserve<WideNode>(icpsymm_ctx*, WideNode**, size_t*, size_t);
template int icp_reserve<uint32_t>(icpsymm_ctx*, uint32_t**, size_t*, size_t);
template int icp_reserve<unsigned long long>(icpsymm_ctx*, unsigned long long**, size_t*, size_t);
template int icp_reserve<double>(icpsymm_ctx*, double**, size_t*, size_t);
template int icp_reserve<float>(icpsymm_ctx*, float**, size_t*, size_t);
template int icp_reserve<char>(icpsymm_ctx*, char**, size_t*, size_t);
template int icp_reserve<uint8_t>(icpsymm_ctx*, uint8_t**, size_t*, size_t);
template int icp_reserve<SegState>(icpsymm_ctx*, SegState**, size_t*, size_t);

cudaStream_t icp_cloud_stream(icpsymm_ctx* ctx, CloudSet& cs) {
    if (!ctx->opt_async_setup) return ctx->stream;
    const int k = cs.slot;
    if (!ctx->aux[k]) {
        if (cudaStreamCreateWithFlags(&ctx->aux[k], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return ctx->stream; }
        cudaEventCreateWithFlags(&ctx->ready[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->built[k], cudaEventDisableTiming);
    }
    if (!ctx->fork_ev) cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
    // order after everything already enqueued on the context stream (it may still read this cloud)
    cudaEventRecord(ctx->fork_ev, ctx->stream);
    cudaStreamWaitEvent(ctx->aux[k], ctx->fork_ev, 0);
    return ctx->aux[k];
}

void icp_cloud_done(icpsymm_ctx* ctx, CloudSet& cs) {
    if (!ctx->opt_async_setup || !ctx->aux[cs.slot]) return;
    cudaEventRecord(ctx->ready[cs.slot], ctx->aux[cs.slot]);
    cs.on_aux = true;
}

int icp_join_clouds(icpsymm_ctx* ctx, bool check) {
    CloudSet* both[2] = {&ctx->src, &ctx->tgt};
    int rc = ICPSYMM_OK;
    for (CloudSet* cs : both) {
        if (cs->on_aux) {
            cudaStreamWaitEvent(ctx->stream, ctx->ready[cs->slot], 0);
            cs->on_aux = false;
        }
        if (check && cs->pending_check) {
            cs->pending_check = false;
            // wait for the BUILD only: the normals that may follow it on the same stream keep running while the host
            // enqueues the loop (waiting for `ready` made the host idle through the normals of every pipelined pair)
            cudaEventSynchronize(ctx->built[cs->slot]);
            if (ctx->scratch[cs->slot].h_flag && *ctx->scratch[cs->slot].h_flag) {
                cs->valid = false;
                rc = icp_fail_msg(ctx, ICPSYMM_ERR_NONFINITE, "cloud holds NaN/Inf coordinates");
            }
        }
    }
    return rc;
}

TreeShape icp_make_shape(uint32_t max_count) {
    TreeShape sh;
    memset(&sh, 0, sizeof sh);
    sh.pitch = round_up(max_count ? max_count : 1, ICP_LEAF);
    uint32_t cnt = sh.pitch / ICP_LEAF;
    sh.level_cnt[0] = cnt == 1 ? 1 : round_up(cnt, ICP_FAN);
    cnt = sh.level_cnt[0];
    uint32_t off = 0, l = 0;
    while (cnt > 1 && l + 1 < ICP_MAX_LEVELS) {
        l++;
        uint32_t n = cnt / ICP_FAN;
        sh.level_cnt[l] = n == 1 ? 1 : round_up(n, ICP_FAN);
        sh.level_off[l] = off;
        off += sh.level_cnt[l];
        cnt = sh.level_cnt[l];
    }
    sh.n_levels = (cnt > 1) ? 0 : l + 1;   // 0 = too many points for ICP_MAX_LEVELS
    sh.nodes_per_seg = off ? off : 1;
    sh.units_per_seg = (sh.pitch + 31u) / 32u;
    return sh;
}

void icp_free_cloudset(CloudSet& cs) {
    cudaFree(cs.pts); cudaFree(cs.nrm); cudaFree(cs.nodes); cudaFree(cs.unit_split); cudaFree(cs.pos_of_orig); cudaFree(cs.counts);
    cudaFree(cs.bbox); cudaFree(cs.bb_lo); cudaFree(cs.bb_hi);
    free(cs.h_counts);
    cs = CloudSet();
}

int icp_build_cloudset(icpsymm_ctx* ctx, CloudSet& cs, const float* xyz, bool on_device, const uint32_t* counts,
                       size_t nseg, size_t max_count, size_t stride_bytes) {
    if (!xyz || nseg == 0 || max_count == 0) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "empty cloud");
    if (stride_bytes < 12 || (stride_bytes & 3)) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "stride_bytes must be >= 12 and a multiple of 4");
    if (max_count > 0x7ffffff0ull || nseg > 0x100000ull) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "cloud too large");
    cudaStream_t st = icp_cloud_stream(ctx, cs);
    BuildScratch& B = ctx->scratch[cs.slot];
    cs.valid = false; cs.has_normals = false; cs.pending_check = false;
    TreeShape sh = icp_make_shape((uint32_t)max_count);
    if (sh.n_levels == 0) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "cloud too large for the 8-level hierarchy");
    size_t slots = nseg * (size_t)sh.pitch, nodes = nseg * (size_t)sh.nodes_per_seg;

    // host copy of the counts
    free(cs.h_counts);
    cs.h_counts = (uint32_t*)malloc(sizeof(uint32_t) * nseg);
    size_t total = 0;
    for (size_t s = 0; s < nseg; s++) {
        uint32_t c = counts ? counts[s] : (uint32_t)max_count;
        if (c > max_count) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "counts[b] > max_count");
        cs.h_counts[s] = c;
        total += c;
    }
    int rc;
    size_t cap;
    cap = cs.cap_pts;  if ((rc = icp_reserve(ctx, &cs.pts, &cap, slots))) return rc;
    cap = cs.cap_pts;  if ((rc = icp_reserve(ctx, &cs.nrm, &cap, slots))) return rc;
    cap = cs.cap_pts;  if ((rc = icp_reserve(ctx, &cs.pos_of_orig, &cap, slots))) return rc;
    cs.cap_pts = cap;
    if ((rc = icp_reserve(ctx, &cs.nodes, &cs.cap_nodes, nodes))) return rc;
    if ((rc = icp_reserve(ctx, &cs.unit_split, &cs.cap_units, nseg * (size_t)sh.units_per_seg))) return rc;
    cap = cs.cap_seg;  if ((rc = icp_reserve(ctx, &cs.counts, &cap, nseg))) return rc;
    cap = cs.cap_seg * 8; if ((rc = icp_reserve(ctx, &cs.bbox, &cap, nseg * 8))) return rc;
    cap = cs.cap_seg;  if ((rc = icp_reserve(ctx, &cs.bb_lo, &cap, nseg))) return rc;
    cap = cs.cap_seg;  if ((rc = icp_reserve(ctx, &cs.bb_hi, &cap, nseg))) return rc;
    cs.cap_seg = cap;
    cap = B.sort_cap; if ((rc = icp_reserve(ctx, &B.keys_a, &cap, slots))) return rc;
    cap = B.sort_cap; if ((rc = icp_reserve(ctx, &B.keys_b, &cap, slots))) return rc;
    cap = B.sort_cap; if ((rc = icp_reserve(ctx, &B.vals_a, &cap, slots))) return rc;
    cap = B.sort_cap; if ((rc = icp_reserve(ctx, &B.vals_b, &cap, slots))) return rc;
    B.sort_cap = cap;
    if (!B.d_flags) ICP_CUDA_CHECK(cudaMalloc((void**)&B.d_flags, 4 * sizeof(int)));
    if (!B.h_flag) ICP_CUDA_CHECK(cudaMallocHost((void**)&B.h_flag, sizeof(int)));

    const char* raw = reinterpret_cast<const char*>(xyz);
    if (!on_device) {
        size_t bytes = nseg * max_count * stride_bytes;
        char* stg = (char*)B.staging;
        if ((rc = icp_reserve(ctx, &stg, &B.staging_bytes, bytes))) return rc;
        B.staging = stg;
        ICP_CUDA_CHECK(cudaMemcpyAsync(stg, xyz, bytes, cudaMemcpyHostToDevice, st));
        raw = stg;
    }
    ICP_CUDA_CHECK(cudaMemcpyAsync(cs.counts, cs.h_counts, sizeof(uint32_t) * nseg, cudaMemcpyHostToDevice, st));

    const uint32_t ns = (uint32_t)nseg, mc = (uint32_t)max_count;
    bbox_init_kernel<<<(ns * 8 + 255) / 256, 256, 0, st>>>(cs.bbox, ns, B.d_flags);
    {
        uint32_t bx = (uint32_t)((max_count + 256 * 8 - 1) / (256 * 8));
        uint32_t cap_bx = (uint32_t)((ctx->sm_count * 8 + ns - 1) / ns);
        if (bx > cap_bx) bx = cap_bx;
        if (bx < 1) bx = 1;
        dim3 grid(bx, ns);
        bbox_kernel<<<grid, 256, 0, st>>>(raw, stride_bytes, cs.counts, mc, cs.bbox, B.d_flags);
    }
    bbox_finish_kernel<<<(ns + 127) / 128, 128, 0, st>>>(cs.bbox, cs.counts, cs.bb_lo, cs.bb_hi, ns);
    const uint32_t gb = (uint32_t)((slots + 255) / 256);
    // 10 bits per axis.  A coarser grid would do for the windows of the k-d refinement and saves a radix
    // pass (8 bits: index build -36 us per 1M cloud), but the upper levels of the hierarchy get worse:
    // loop +1.6 %, normals +1.4 % -- measured, not adopted.
#ifdef ICP_CURVE_BITS
    const int curve_bits = ICP_CURVE_BITS;
#else
    const int curve_bits = 10;
#endif
    int seg_bits = 0;
    while ((1ull << seg_bits) < nseg) seg_bits++;
    const int end_bit = 3 * curve_bits + seg_bits;
    size_t temp_bytes = 0;
    if (end_bit <= 32 && slots < 0x7fffffffull) {
        // 32-bit keys and 32-bit offsets (the key buffers are reused as they are)
        uint32_t* ka = reinterpret_cast<uint32_t*>(B.keys_a);
        uint32_t* kb = reinterpret_cast<uint32_t*>(B.keys_b);
        curve_key_kernel<uint32_t><<<gb, 256, 0, st>>>(raw, stride_bytes, cs.counts, mc, sh.pitch, ns, cs.bb_lo, cs.bb_hi,
                                                       curve_bits, ka, B.vals_a);
        ICP_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, ka, kb, B.vals_a, B.vals_b, (int)slots, 0, end_bit, st));
        char* tmp = (char*)B.cub_temp;
        if ((rc = icp_reserve(ctx, &tmp, &B.cub_temp_bytes, temp_bytes))) return rc;
        B.cub_temp = tmp;
        ICP_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(B.cub_temp, temp_bytes, ka, kb, B.vals_a, B.vals_b, (int)slots, 0, end_bit, st));
    } else {
        curve_key_kernel<unsigned long long><<<gb, 256, 0, st>>>(raw, stride_bytes, cs.counts, mc, sh.pitch, ns, cs.bb_lo,
                                                                 cs.bb_hi, curve_bits, B.keys_a, B.vals_a);
        ICP_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, B.keys_a, B.keys_b, B.vals_a,
                                                       B.vals_b, slots, 0, end_bit, st));
        char* tmp = (char*)B.cub_temp;
        if ((rc = icp_reserve(ctx, &tmp, &B.cub_temp_bytes, temp_bytes))) return rc;
        B.cub_temp = tmp;
        ICP_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(B.cub_temp, temp_bytes, B.keys_a, B.keys_b, B.vals_a,
                                                       B.vals_b, slots, 0, end_bit, st));
    }
    gather_kernel<<<gb, 256, 0, st>>>(raw, stride_bytes, cs.counts, mc, sh.pitch, ns, B.vals_b, cs.pts);
    {
        // k-d refinement of the curve order: coarse median partitions of 8192-point windows (the set that is
        // searched only: they shape the upper levels of its hierarchy; measured on the 1M pair: loop -10 %
        // for 60 us of build time, while the source gains nothing from them), then the sorting
        // refinement of every 512-point window (both sets; only full windows are touched)
#ifndef ICP_KD_COARSE_TGT
#define ICP_KD_COARSE_TGT 1
#endif
#ifndef ICP_KD_COARSE_SRC
#define ICP_KD_COARSE_SRC 0
#endif
        const bool coarse = cs.slot == 1 ? (ICP_KD_COARSE_TGT != 0) : (ICP_KD_COARSE_SRC != 0);
        if (coarse && sh.pitch >= 8192) {
            const size_t sm = (size_t)8192 * sizeof(float4);
            ICP_CUDA_CHECK(cudaFuncSetAttribute(kd_coarse_kernel<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            kd_coarse_kernel<8192><<<ns * (sh.pitch / 8192), 1024, sm, st>>>(cs.pts, cs.counts, sh.pitch, sh.pitch / 8192);
        } else if (coarse && sh.pitch >= 4096) {
            const size_t sm = (size_t)4096 * sizeof(float4);
            ICP_CUDA_CHECK(cudaFuncSetAttribute(kd_coarse_kernel<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            kd_coarse_kernel<4096><<<ns * (sh.pitch / 4096), 1024, sm, st>>>(cs.pts, cs.counts, sh.pitch, sh.pitch / 4096);
        }
        const uint32_t wins = sh.pitch / 512;
        if (wins) kd_refine_kernel<512><<<ns * wins, 256, 0, st>>>(cs.pts, cs.counts, sh.pitch, wins);
    }
    {
        size_t nfl = nodes * 48;
        init_nodes_kernel<<<(uint32_t)((nfl + 255) / 256), 256, 0, st>>>(cs.nodes, nfl);
        const size_t thr = (size_t)ns * sh.units_per_seg * 32;
        finish_units_kernel<<<(uint32_t)((thr + 255) / 256), 256, 0, st>>>(cs.pts, cs.counts, sh, ns, cs.pos_of_orig, cs.unit_split,
                                                                           sh.n_levels > 1 ? cs.nodes : nullptr);
        // large levels: one thread per node over the whole grid; the small ones at the top (at most 512 nodes
        // each): together in one block per segment
        uint32_t l = 1;
        for (; l + 1 < sh.n_levels && sh.level_cnt[l] > 512; l++) {
            const size_t nl = (size_t)ns * sh.level_cnt[l];
            upper_box_kernel<<<(uint32_t)((nl + 255) / 256), 256, 0, st>>>(cs.nodes, sh, ns, (int)l, sh.level_off[l],
                                                                           sh.level_cnt[l], sh.level_off[l + 1]);
        }
        if (l + 1 < sh.n_levels) upper_levels_kernel<<<ns, 512, 0, st>>>(cs.nodes, sh, l);
    }
    ICP_CUDA_CHECK(cudaMemsetAsync(cs.nrm, 0, slots * sizeof(float4), st));
    ICP_CUDA_CHECK(cudaGetLastError());
    ICP_CUDA_CHECK(cudaMemcpyAsync(B.h_flag, B.d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (ctx->opt_async_setup && st != ctx->stream) {
        // asynchronous setup: no host wait here; the flag is read at the next point that hands data
        // to the host (icp_join_clouds), so the upload of the other cloud overlaps this build
        ICP_CUDA_CHECK(cudaEventRecord(ctx->built[cs.slot], st));
        icp_cloud_done(ctx, cs);
        cs.pending_check = true;
    } else {
        ICP_CUDA_CHECK(cudaStreamSynchronize(st));
        if (*B.h_flag) return icp_fail_msg(ctx, ICPSYMM_ERR_NONFINITE, "cloud holds NaN/Inf coordinates");
    }
    cs.nseg = ns; cs.max_count = mc; cs.total = total; cs.shape = sh; cs.valid = true;
    return ICPSYMM_OK;
}
