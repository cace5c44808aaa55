// normals.cu -- per-point normal estimation (MyICP::estimateNormals, ICP/myicp.cpp:152-172).
//
// The reference calls pcl::NormalEstimation with setKSearch(10) (myicp.cpp:160) on each cloud:
// kd-tree kNN, 3x3 covariance of the neighbourhood, eigenvector of the smallest eigenvalue,
// flip towards the viewpoint (0,0,0), curvature = l0 / (l0+l1+l2).  Here one thread owns one
// Hilbert-sorted point: exact kNN-k over the cloud's own box hierarchy (self included, ties by
// lowest original index), covariance about the neighbourhood mean taken relative to the query
// point (so the fp32 subtractions are exact), cyclic-Jacobi eigen-decomposition in registers.
#include <cfloat>

#include "internal.h"
#include "search.cuh"

#ifdef ICP_BLOCK_TIMES
__device__ unsigned long long icp_blk_times[3 * 65536];
struct BlockTimer {
    unsigned long long t0;
    __device__ BlockTimer() { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); }
    __device__ ~BlockTimer() {
        __syncthreads();
        if (threadIdx.x == 0 && blockIdx.x < 65536) {
            unsigned long long t1; unsigned int sm;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            icp_blk_times[3 * blockIdx.x] = t0; icp_blk_times[3 * blockIdx.x + 1] = t1; icp_blk_times[3 * blockIdx.x + 2] = sm;
        }
    }
};
extern "C" __attribute__((visibility("default"))) int icpsymm_debug_sorted_points(icpsymm_ctx* ctx, int which, float* out, size_t n) {
    cudaDeviceSynchronize();
    const CloudSet& cs = which ? ctx->tgt : ctx->src;
    return (int)cudaMemcpy(out, cs.pts, n * sizeof(float4), cudaMemcpyDeviceToHost);
}
extern "C" __attribute__((visibility("default"))) int icpsymm_debug_unit_split(icpsymm_ctx* ctx, int which, uint32_t* out, size_t n) {
    cudaDeviceSynchronize();
    const CloudSet& cs = which ? ctx->tgt : ctx->src;
    return (int)cudaMemcpy(out, cs.unit_split, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
}
extern "C" __attribute__((visibility("default"))) int icpsymm_debug_block_times(unsigned long long* out, int n) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, icp_blk_times, sizeof(unsigned long long) * 3 * n);
}
#endif

namespace {

struct NormalParams {
    CloudView cloud;
    float4* nrm_out;   // sorted layout
    float vp[3];
    int k;
};

__device__ __forceinline__ void jacobi_rotate(float& app, float& aqq, float& apq, float& arp, float& arq,
                                              float* vp, float* vq) {
    // annihilate apq of the symmetric 3x3; (arp, arq) are the two remaining off-diagonal entries
    if (apq == 0.f) return;
    float theta = (aqq - app) / (2.f * apq);
    float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
    float c = rsqrtf(t * t + 1.f), s = t * c;
    float napp = app - t * apq, naqq = aqq + t * apq;
    float narp = c * arp - s * arq, narq = s * arp + c * arq;
    app = napp; aqq = naqq; apq = 0.f; arp = narp; arq = narq;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float a = vp[k], b = vq[k];
        vp[k] = c * a - s * b;
        vq[k] = s * a + c * b;
    }
}


#ifndef ICP_NRM_BLOCKS
#define ICP_NRM_BLOCKS 7
#endif
template <int KMAX, bool EXACT>
__global__ void __launch_bounds__(128, (KMAX <= 16) ? ICP_NRM_BLOCKS : 4) normals_kernel(const __grid_constant__ NormalParams P) {
#ifdef ICP_BLOCK_TIMES
    BlockTimer bt_;
#endif
    // one block = 128 consecutive sorted points of ONE segment, so that a warp never straddles two
    // clouds and its 32 lanes can walk the hierarchy together
    const uint32_t pitch = P.cloud.shape.pitch;
    const uint32_t chunks = (pitch + 127u) / 128u;
    const uint32_t seg = blockIdx.x / chunks, i = (blockIdx.x - seg * chunks) * 128u + threadIdx.x;
    const size_t j = (size_t)seg * pitch + i;
    const uint32_t cnt = P.cloud.counts[seg];
    const bool active = i < cnt;
    const float4* __restrict__ pts = P.cloud.pts + (size_t)seg * pitch;
    const WideNode* __restrict__ boxes = P.cloud.nodes + (size_t)seg * P.cloud.shape.nodes_per_seg;
    const float4 q = active ? ldg4(pts + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = P.k;

    typedef KList<KMAX, EXACT> List;
    __shared__ unsigned long long s_park[List::CAP][128];   // [slot][thread]: conflict-free columns
    List L;
    L.init(&s_park[0][threadIdx.x], 128);
    const uint32_t n_groups = pitch / ICP_LEAF;
    // first bound: the aligned 64-point (k <= 16) or 128-point cell that holds the query -- a k-d cell of the refined order
    const uint32_t g = active ? i / ICP_LEAF : 0u, w = (k <= 16) ? 8u : 16u;
    const uint32_t g_lo = g / w * w, g_hi = (g_lo + w - 1 < n_groups) ? g_lo + w - 1 : n_groups - 1;
    // the 32 lanes of a warp share the cell: every lane takes all of its leaf groups
    {
        const uint32_t w_lo = __reduce_min_sync(0xffffffffu, active ? g_lo : 0xffffffffu);
        const uint32_t w_hi = __reduce_max_sync(0xffffffffu, active ? g_hi : 0u);
        if (w_lo != 0xffffffffu) {
#pragma unroll 1
            for (uint32_t gg = w_lo; gg <= w_hi; gg++) {
                knn_offer_leaf(L, pts, gg, q.x, q.y, q.z, active && gg >= g_lo && gg <= g_hi);
                L.maybe_flush(k, pts);
            }
        }
    }
    // warp-uniform: blocks start at multiples of 128 points, so a warp is exactly one 32-point unit
    const uint32_t split = __ldg(P.cloud.unit_split + (size_t)seg * P.cloud.shape.units_per_seg + (i >> 5));
    knn_traverse_warp<KMAX, EXACT>(P.cloud.shape, boxes, pts, q.x, q.y, q.z, k, g_lo, g_hi, active, split, L);
    if (!active) return;

    // covariance of the neighbourhood (the parked k nearest), coordinates relative to the query point
    float ux[KMAX], uy[KMAX], uz[KMAX];
    int m = 0;
    float mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll
    for (int t = 0; t < KMAX; t++) {
        ux[t] = uy[t] = uz[t] = 0.f;
        if (t < L.cnt) {
            ICP_ASSERT((uint32_t)L.buf[t * L.stride] < cnt && L.cnt <= k);
            const float4 c = ldg4(pts + (uint32_t)L.buf[t * L.stride]);
            ux[t] = c.x - q.x; uy[t] = c.y - q.y; uz[t] = c.z - q.z;
            mx += ux[t]; my += uy[t]; mz += uz[t];
            m++;
        }
    }
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m >= 3) {
        const float inv = 1.f / (float)m;
        mx *= inv; my *= inv; mz *= inv;
        float cxx = 0.f, cxy = 0.f, cxz = 0.f, cyy = 0.f, cyz = 0.f, czz = 0.f;
#pragma unroll
        for (int t = 0; t < KMAX; t++) {
            if (t < m) {
                float ex = ux[t] - mx, ey = uy[t] - my, ez = uz[t] - mz;
                cxx += ex * ex; cxy += ex * ey; cxz += ex * ez; cyy += ey * ey; cyz += ey * ez; czz += ez * ez;
            }
        }
        cxx *= inv; cxy *= inv; cxz *= inv; cyy *= inv; cyz *= inv; czz *= inv;
        // scale to O(1) so tiny neighbourhoods do not underflow the rotation tests
        float sc = fmaxf(fmaxf(fabsf(cxx), fabsf(cyy)), fmaxf(fabsf(czz), fmaxf(fabsf(cxy), fmaxf(fabsf(cxz), fabsf(cyz)))));
        if (sc > 0.f) {
            float is = 1.f / sc;
            float a00 = cxx * is, a01 = cxy * is, a02 = cxz * is, a11 = cyy * is, a12 = cyz * is, a22 = czz * is;
            float v0[3] = {1.f, 0.f, 0.f}, v1[3] = {0.f, 1.f, 0.f}, v2[3] = {0.f, 0.f, 1.f};
#pragma unroll 1
            for (int sweep = 0; sweep < 12; sweep++) {
                float off = a01 * a01 + a02 * a02 + a12 * a12;
                if (off < 1e-18f) break;
                jacobi_rotate(a00, a11, a01, a02, a12, v0, v1);  // (p,q)=(0,1): r=2 -> a20,a21
                jacobi_rotate(a00, a22, a02, a01, a12, v0, v2);  // (p,q)=(0,2): r=1 -> a10,a12
                jacobi_rotate(a11, a22, a12, a01, a02, v1, v2);  // (p,q)=(1,2): r=0 -> a01,a02
            }
            float l0 = a00, l1 = a11, l2 = a22;
            float nx = v0[0], ny = v0[1], nz = v0[2];
            if (l1 < l0) { l0 = a11; nx = v1[0]; ny = v1[1]; nz = v1[2]; }
            if (l2 < l0) { l0 = a22; nx = v2[0]; ny = v2[1]; nz = v2[2]; }
            float nn = rsqrtf(nx * nx + ny * ny + nz * nz);
            nx *= nn; ny *= nn; nz *= nn;
            // flipNormalTowardsViewpoint: n . (vp - p) >= 0
            float dotv = (P.vp[0] - q.x) * nx + (P.vp[1] - q.y) * ny + (P.vp[2] - q.z) * nz;
            if (dotv < 0.f) { nx = -nx; ny = -ny; nz = -nz; }
            float tr = a00 + a11 + a22;
            out = make_float4(nx, ny, nz, tr > 0.f ? fabsf(l0 / tr) : 0.f);
        } else {
            // all neighbours coincide: no direction
            out = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    P.nrm_out[j] = out;
}

__global__ void __launch_bounds__(256) export_normals_kernel(CloudView c, uint32_t max_count, float4* __restrict__ out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pitch = c.shape.pitch;
    if (j >= (size_t)c.nseg * pitch) return;
    uint32_t seg = (uint32_t)(j / pitch), i = (uint32_t)(j - (size_t)seg * pitch);
    if (i >= c.counts[seg]) return;
    out[(size_t)seg * max_count + __float_as_uint(c.pts[j].w)] = c.nrm[j];
}

__global__ void __launch_bounds__(256) import_normals_kernel(CloudView c, uint32_t max_count, const float4* __restrict__ in,
                                                             float4* __restrict__ nrm) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pitch = c.shape.pitch;
    if (j >= (size_t)c.nseg * pitch) return;
    uint32_t seg = (uint32_t)(j / pitch), i = (uint32_t)(j - (size_t)seg * pitch);
    nrm[j] = (i < c.counts[seg]) ? in[(size_t)seg * max_count + __float_as_uint(c.pts[j].w)] : make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace

int icp_estimate_normals(icpsymm_ctx* ctx, CloudSet& cs, int k, const float vp[3]) {
    if (!cs.valid) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "cloud not set");
    if (k < 3 || k > 32) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "k must be in [3, 32]");
    NormalParams P;
    P.cloud = cs.view();
    P.nrm_out = cs.nrm;
    P.vp[0] = vp ? vp[0] : 0.f; P.vp[1] = vp ? vp[1] : 0.f; P.vp[2] = vp ? vp[2] : 0.f;
    P.k = k;
    size_t slots = (size_t)cs.nseg * cs.shape.pitch;
    uint32_t grid = cs.nseg * ((cs.shape.pitch + 127u) / 128u);
    cudaStream_t st = icp_cloud_stream(ctx, cs);   // the cloud's aux stream under asynchronous setup
    ICP_CUDA_CHECK(cudaMemsetAsync(cs.nrm, 0, slots * sizeof(float4), st));
    if (k == 10) normals_kernel<10, true><<<grid, 128, 0, st>>>(P);       // the reference's k (myicp.cpp:160)
    else if (k < 10) normals_kernel<10, false><<<grid, 128, 0, st>>>(P);
    else if (k <= 16) normals_kernel<16, false><<<grid, 128, 0, st>>>(P);
    else normals_kernel<32, false><<<grid, 128, 0, st>>>(P);
    ICP_CUDA_CHECK(cudaGetLastError());
    icp_cloud_done(ctx, cs);
    cs.has_normals = true;
    return ICPSYMM_OK;
}

static int ensure_io_n(icpsymm_ctx* ctx, size_t bytes) {
    char* p = (char*)ctx->io_buf;
    int rc = icp_reserve(ctx, &p, &ctx->io_bytes, bytes);
    ctx->io_buf = p;
    return rc;
}

int icp_export_normals(icpsymm_ctx* ctx, const CloudSet& cs, float* out_host) {
    { int jr = icp_join_clouds(ctx, true); if (jr < 0) return jr; }
    if (!cs.valid || !cs.has_normals) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "normals missing");
    size_t n = (size_t)cs.nseg * cs.max_count;
    int rc = ensure_io_n(ctx, n * sizeof(float4));
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    ICP_CUDA_CHECK(cudaMemsetAsync(ctx->io_buf, 0, n * sizeof(float4), st));
    size_t slots = (size_t)cs.nseg * cs.shape.pitch;
    export_normals_kernel<<<(uint32_t)((slots + 255) / 256), 256, 0, st>>>(cs.view(), cs.max_count, (float4*)ctx->io_buf);
    ICP_CUDA_CHECK(cudaGetLastError());
    ICP_CUDA_CHECK(cudaMemcpyAsync(out_host, ctx->io_buf, n * sizeof(float4), cudaMemcpyDeviceToHost, st));
    ICP_CUDA_CHECK(cudaStreamSynchronize(st));
    return ICPSYMM_OK;
}

int icp_import_normals(icpsymm_ctx* ctx, CloudSet& cs, const float* in_host) {
    { int jr = icp_join_clouds(ctx, true); if (jr < 0) return jr; }
    if (!cs.valid) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "cloud not set");
    if (!in_host) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "null normals");
    size_t n = (size_t)cs.nseg * cs.max_count;
    int rc = ensure_io_n(ctx, n * sizeof(float4));
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    ICP_CUDA_CHECK(cudaMemcpyAsync(ctx->io_buf, in_host, n * sizeof(float4), cudaMemcpyHostToDevice, st));
    size_t slots = (size_t)cs.nseg * cs.shape.pitch;
    import_normals_kernel<<<(uint32_t)((slots + 255) / 256), 256, 0, st>>>(cs.view(), cs.max_count,
                                                                           (const float4*)ctx->io_buf, cs.nrm);
    ICP_CUDA_CHECK(cudaGetLastError());
    ICP_CUDA_CHECK(cudaStreamSynchronize(st));
    cs.has_normals = true;
    return ICPSYMM_OK;
}
