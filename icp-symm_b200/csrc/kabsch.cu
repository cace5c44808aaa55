// kabsch.cu -- closed-form rigid fit of paired points (registrateNPoint, ICP/regist.h:8-72), fp64.
//
// regist.h: centroids by sequential sums (:16-35), centred 3xN matrices (:36-50), S = src~ dst~^T
// (:52), cv::SVDecomp (:55), det(U Vt) (:57-58), R = Vt^T diag(1,1,det) U^T (:60-63),
// T = dstbar - R srcbar (:64-69); net effect dst = R src + T.  Here: two deterministic
// block-partial reductions (centroids, then the centred 3x3 cross-covariance) and a one-thread
// one-sided Jacobi SVD of the 3x3.
#include <cstring>

#include "internal.h"

namespace {

constexpr int kThreads = 256;
constexpr int kRow = 16;

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// mode 0: rows = [sum src (3), sum dst (3)];  mode 1: rows = S (9) about the means in `mean`
__global__ void __launch_bounds__(kThreads) kabsch_reduce_kernel(const double* __restrict__ src, const double* __restrict__ dst,
                                                                 size_t n, const double* __restrict__ mean, int mode,
                                                                 double* __restrict__ partial) {
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) acc[k] = 0.0;
    double ms[3] = {0, 0, 0}, md[3] = {0, 0, 0};
    if (mode == 1) {
        for (int k = 0; k < 3; k++) { ms[k] = mean[k]; md[k] = mean[3 + k]; }
    }
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        double a[3] = {src[3 * i], src[3 * i + 1], src[3 * i + 2]};
        double b[3] = {dst[3 * i], dst[3 * i + 1], dst[3 * i + 2]};
        if (mode == 0) {
            acc[0] += a[0]; acc[1] += a[1]; acc[2] += a[2];
            acc[3] += b[0]; acc[4] += b[1]; acc[5] += b[2];
        } else {
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) acc[3 * r + c] += (a[r] - ms[r]) * (b[c] - md[c]);
        }
    }
    __shared__ double sw[kThreads / 32][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        double v = warp_sum(acc[k]);
        if (lane == 0) sw[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; w++) s += sw[w][threadIdx.x];
        partial[(size_t)blockIdx.x * kRow + threadIdx.x] = s;
    }
}

__global__ void kabsch_means_kernel(const double* __restrict__ partial, int nblocks, size_t n, double* __restrict__ mean) {
    int k = threadIdx.x;
    if (k >= 6) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; b++) s += partial[(size_t)b * kRow + k];
    mean[k] = s / (double)n;
}

__device__ void svd3_dev(const double* A, double* U, double* s, double* V) {
    double a[9];
    for (int i = 0; i < 9; i++) { a[i] = A[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 80; sweep++) {
        int rotated = 0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double al = 0, be = 0, ga = 0;
                for (int i = 0; i < 3; i++) {
                    al += a[3 * i + p] * a[3 * i + p];
                    be += a[3 * i + q] * a[3 * i + q];
                    ga += a[3 * i + p] * a[3 * i + q];
                }
                if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
                rotated = 1;
                double zeta = (be - al) / (2.0 * ga);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int i = 0; i < 3; i++) {
                    double ap = a[3 * i + p], aq = a[3 * i + q];
                    a[3 * i + p] = c * ap - sn * aq;
                    a[3 * i + q] = sn * ap + c * aq;
                    double vp = V[3 * i + p], vq = V[3 * i + q];
                    V[3 * i + p] = c * vp - sn * vq;
                    V[3 * i + q] = sn * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double sv[3];
    for (int j = 0; j < 3; j++) sv[j] = sqrt(a[j] * a[j] + a[3 + j] * a[3 + j] + a[6 + j] * a[6 + j]);
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2 - i; j++)
            if (sv[ord[j]] < sv[ord[j + 1]]) { int t = ord[j]; ord[j] = ord[j + 1]; ord[j + 1] = t; }
    double Vs[9], Us[9];
    for (int j = 0; j < 3; j++) {
        s[j] = sv[ord[j]];
        for (int i = 0; i < 3; i++) {
            Vs[3 * i + j] = V[3 * i + ord[j]];
            Us[3 * i + j] = s[j] > 0 ? a[3 * i + ord[j]] / s[j] : 0.0;
        }
    }
    double tol = 1e-13 * (s[0] > 0 ? s[0] : 1.0);
    if (s[0] <= tol) {
        for (int i = 0; i < 9; i++) Us[i] = (i % 4 == 0) ? 1.0 : 0.0;
    } else {
        if (s[1] <= tol) {
            double u0[3] = {Us[0], Us[3], Us[6]};
            int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
            double e[3] = {0, 0, 0};
            e[m] = 1.0;
            double w[3] = {e[0] - u0[m] * u0[0], e[1] - u0[m] * u0[1], e[2] - u0[m] * u0[2]};
            double nw = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            Us[1] = w[0] / nw; Us[4] = w[1] / nw; Us[7] = w[2] / nw;
        }
        if (s[2] <= tol) {
            Us[2] = Us[3] * Us[7] - Us[6] * Us[4];
            Us[5] = Us[6] * Us[1] - Us[0] * Us[7];
            Us[8] = Us[0] * Us[4] - Us[3] * Us[1];
        }
    }
    for (int i = 0; i < 9; i++) { U[i] = Us[i]; V[i] = Vs[i]; }
}

__global__ void kabsch_solve_kernel(const double* __restrict__ partial, int nblocks, const double* __restrict__ mean,
                                    double* __restrict__ out /*R 9, T 3*/) {
    if (threadIdx.x != 0) return;
    double S[9];
    for (int k = 0; k < 9; k++) {
        double s = 0.0;
        for (int b = 0; b < nblocks; b++) s += partial[(size_t)b * kRow + k];
        S[k] = s;
    }
    double U[9], W[3], V[9];
    svd3_dev(S, U, W, V);
    double UVt[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) UVt[3 * r + c] = U[3 * r] * V[3 * c] + U[3 * r + 1] * V[3 * c + 1] + U[3 * r + 2] * V[3 * c + 2];
    double det = UVt[0] * (UVt[4] * UVt[8] - UVt[5] * UVt[7]) - UVt[1] * (UVt[3] * UVt[8] - UVt[5] * UVt[6]) +
                 UVt[2] * (UVt[3] * UVt[7] - UVt[4] * UVt[6]);
    double M[3] = {1.0, 1.0, det};
    double R[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            R[3 * r + c] = V[3 * r] * M[0] * U[3 * c] + V[3 * r + 1] * M[1] * U[3 * c + 1] + V[3 * r + 2] * M[2] * U[3 * c + 2];
    for (int i = 0; i < 9; i++) out[i] = R[i];
    for (int r = 0; r < 3; r++)
        out[9 + r] = mean[3 + r] - (R[3 * r] * mean[0] + R[3 * r + 1] * mean[1] + R[3 * r + 2] * mean[2]);
}

// pairs of the point-to-point loop, straight from the context's device state: (T p_j, target point matched to p_j)
// for every sorted source position j, as the fp64 arrays the reduction kernels read
struct T12f { float m[12]; };
__global__ void __launch_bounds__(256) p2p_pairs_kernel(CloudView src, CloudView tgt, const uint32_t* __restrict__ corr,
                                                        T12f T, double* __restrict__ a, double* __restrict__ b) {
    const uint32_t j = blockIdx.x * 256u + threadIdx.x;
    if (j >= src.counts[0]) return;
    const float4 s = src.pts[j];
    const float3 p = apply_rn(T.m, s.x, s.y, s.z);
    const uint32_t pos = corr[j];
    const float4 q = tgt.pts[pos < tgt.counts[0] ? pos : 0];
    a[3 * (size_t)j] = p.x; a[3 * (size_t)j + 1] = p.y; a[3 * (size_t)j + 2] = p.z;
    b[3 * (size_t)j] = q.x; b[3 * (size_t)j + 1] = q.y; b[3 * (size_t)j + 2] = q.z;
}

}  // namespace

// Point-to-point ICP, the loop MyICP::RegisterP2P stands for (the reference's is a stub that applies an identity guess,
// myicp.cpp:43-59): exact correspondences (K1), closed-form fit of the matched pairs (registrateNPoint, regist.h:8-72),
// transform = incre * transform, until diff <= threshold or max_iters (the loop shape of myicp.cpp:121-142).
// Everything stays on the device; per iteration the host reads back one diff and one 3x4 increment.
int icp_register_p2p(icpsymm_ctx* ctx, int max_iters, float thr, float* T_out, icpsymm_stats* stats) {
    CloudSet& S = ctx->src;
    CloudSet& Tg = ctx->tgt;
    { int jr = icp_join_clouds(ctx, true); if (jr < 0) return jr; }
    if (!S.valid || !Tg.valid) return icp_fail_msg(ctx, ICPSYMM_ERR_NOT_READY, "set_source / set_target first");
    if (S.nseg != 1 || Tg.nseg != 1) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "RegisterP2P takes single pairs");
    if (max_iters < 0) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "max_iters < 0");
    const size_t n = S.h_counts[0];
    if (n < 3 || Tg.h_counts[0] < 1) return icp_fail_msg(ctx, ICPSYMM_ERR_SIZE, "registrateNPoint needs >= 3 pairs (regist.h:9)");
    int nblocks = (int)((n + kThreads - 1) / kThreads);
    if (nblocks > 1024) nblocks = 1024;
    const size_t pts_bytes = sizeof(double) * 3 * n;
    const size_t need = 2 * pts_bytes + sizeof(double) * ((size_t)nblocks * kRow + 32);
    cudaStream_t st = ctx->stream;
    double Tt[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    icpsymm_stats last;
    memset(&last, 0, sizeof last);
    int it = 0, rc = ICPSYMM_OK;
    float total_ms = 0.f;
    for (;;) {
        float Tf[16];
        for (int i = 0; i < 16; i++) Tf[i] = (float)Tt[i];
        rc = icp_run_single(ctx, Tf, nullptr, &last);            // search + diff on the device (K1 + K2)
        if (rc < 0) return rc;
        total_ms += last.loop_ms;
        if (!(last.diff > (double)thr) || it >= max_iters) break;   // myicp.cpp:123
        char* p = (char*)ctx->io_buf;                            // (icp_run_single does not use the scratch)
        if ((rc = icp_reserve(ctx, &p, &ctx->io_bytes, need))) return rc;
        ctx->io_buf = p;
        double* d_src = (double*)p;
        double* d_dst = d_src + 3 * n;
        double* d_part = d_dst + 3 * n;
        double* d_mean = d_part + (size_t)nblocks * kRow;
        double* d_out = d_mean + 8;
        T12f t12;
        for (int i = 0; i < 12; i++) t12.m[i] = Tf[i];
        p2p_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S.view(), Tg.view(), ctx->corr_pos, t12, d_src, d_dst);
        kabsch_reduce_kernel<<<nblocks, kThreads, 0, st>>>(d_src, d_dst, n, nullptr, 0, d_part);
        kabsch_means_kernel<<<1, 32, 0, st>>>(d_part, nblocks, n, d_mean);
        kabsch_reduce_kernel<<<nblocks, kThreads, 0, st>>>(d_src, d_dst, n, d_mean, 1, d_part);
        kabsch_solve_kernel<<<1, 32, 0, st>>>(d_part, nblocks, d_mean, d_out);
        ICP_CUDA_CHECK(cudaGetLastError());
        double o[12];
        ICP_CUDA_CHECK(cudaMemcpyAsync(o, d_out, sizeof o, cudaMemcpyDeviceToHost, st));
        ICP_CUDA_CHECK(cudaStreamSynchronize(st));
        const double inc[16] = {o[0], o[1], o[2], o[9], o[3], o[4], o[5], o[10], o[6], o[7], o[8], o[11], 0, 0, 0, 1};
        double out[16];
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) {
                double sacc = 0;
                for (int k = 0; k < 4; k++) sacc += inc[4 * r + k] * Tt[4 * k + c];
                out[4 * r + c] = sacc;
            }
        memcpy(Tt, out, sizeof out);                             // transform = incre * transform (myicp.cpp:138)
        it++;
    }
    if (T_out) for (int i = 0; i < 16; i++) T_out[i] = (float)Tt[i];
    if (stats) {
        *stats = last;
        stats->iters = it; stats->searches = it + 1; stats->converged = !(last.diff > (double)thr) ? 1 : 0;
        stats->loop_ms = total_ms;
    }
    return ICPSYMM_OK;
}

int icp_kabsch(icpsymm_ctx* ctx, const double* src, const double* dst, size_t n, double R[9], double T[3]) {
    if (!src || !dst || !R || !T) return icp_fail_msg(ctx, ICPSYMM_ERR_INVALID, "null argument");
    if (n < 3) return icp_fail_msg(ctx, ICPSYMM_ERR_SIZE, "registrateNPoint needs >= 3 pairs (regist.h:9)");
    int nblocks = (int)((n + kThreads - 1) / kThreads);
    if (nblocks > 1024) nblocks = 1024;
    size_t pts_bytes = sizeof(double) * 3 * n;
    size_t need = 2 * pts_bytes + sizeof(double) * ((size_t)nblocks * kRow + 32);
    char* p = (char*)ctx->io_buf;
    int rc = icp_reserve(ctx, &p, &ctx->io_bytes, need);
    ctx->io_buf = p;
    if (rc) return rc;
    double* d_src = (double*)p;
    double* d_dst = d_src + 3 * n;
    double* d_part = d_dst + 3 * n;
    double* d_mean = d_part + (size_t)nblocks * kRow;
    double* d_out = d_mean + 8;
    cudaStream_t st = ctx->stream;
    ICP_CUDA_CHECK(cudaMemcpyAsync(d_src, src, pts_bytes, cudaMemcpyHostToDevice, st));
    ICP_CUDA_CHECK(cudaMemcpyAsync(d_dst, dst, pts_bytes, cudaMemcpyHostToDevice, st));
    kabsch_reduce_kernel<<<nblocks, kThreads, 0, st>>>(d_src, d_dst, n, nullptr, 0, d_part);
    kabsch_means_kernel<<<1, 32, 0, st>>>(d_part, nblocks, n, d_mean);
    kabsch_reduce_kernel<<<nblocks, kThreads, 0, st>>>(d_src, d_dst, n, d_mean, 1, d_part);
    kabsch_solve_kernel<<<1, 32, 0, st>>>(d_part, nblocks, d_mean, d_out);
    ICP_CUDA_CHECK(cudaGetLastError());
    double out[12];
    ICP_CUDA_CHECK(cudaMemcpyAsync(out, d_out, sizeof out, cudaMemcpyDeviceToHost, st));
    ICP_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int i = 0; i < 9; i++) R[i] = out[i];
    for (int i = 0; i < 3; i++) T[i] = out[9 + i];
    return ICPSYMM_OK;
}
