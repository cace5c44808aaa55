/*
 * icp_oracle.h -- CPU ORACLE for the symmetric-ICP registration path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (icp-symm_b200/libicpsymm_b200.so) never links, loads or calls anything in oracle/.
 *
 * It is a plain-C restatement ("port") of the reference's algorithm:
 *   Oracle-L  literal, statement-by-statement restatement of ICP/func.cpp, ICP/myicp.cpp,
 *             ICP/regist.h (identity correspondences, alternating 3+3 SVD solve, quirks kept).
 *   Oracle-C  canonical algorithm the north-star names (exact NN, centred joint 6x6 solve,
 *             half-rotation composition as in the comment at func.cpp:94).
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - pinned: the only known-answer data the reference holds -- the cat fixture
 *     cat_out = Rz(pi/4) * cat + (2.5,0,0) (main.cpp:44-52, register-test.cpp:34-39) --
 *     reproduced by orc_registrate_npoint and by orc_register_symm; cross-checked against
 *     numpy / scipy / OpenCV-FLANN golden vectors in tests/golden (generator committed).
 *   - PARITY UNPINNED at bit level for the arithmetic that lives in un-vendored third-party
 *     code the reference calls (PCL 1.9.1 NormalEstimation / KdTreeFLANN tie order, Eigen
 *     JacobiSVD packet order, OpenCV SVDecomp): the reference cannot be compiled here
 *     (no PCL/Eigen/OpenCV C++ headers), so those are restated from their published
 *     algorithms and compared with tolerances, never bit-for-bit.
 *
 * Conventions: xyz arrays are N x 3 float, row-major, contiguous.  4x4 transforms are
 * row-major (T[4*r+c]).  Normals are N x 4 float (nx, ny, nz, curvature).
 * Build with -ffp-contract=off: the fp32 formulas below are part of the contract.
 */
#ifndef ICP_ORACLE_H
#define ICP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NSUMS 37 /* layout of the normal-equation vector, shared with the device ABI */
/* [0..5]   sum m m^T upper (xx xy xz yy yz zz)     m = (p~ + q~) x n
 * [6..14]  sum m n^T row-major 3x3                 n = n_p + n_q
 * [15..20] sum n n^T upper
 * [21..23] sum m r                                 r = (p~ - q~) . n
 * [24..26] sum n r
 * [27..29] sum (p - origin)   [30..32] sum (q - origin)
 * [33] sum ||p-q||   [34] sum ||p-q||^2   [35] sum r^2   [36] pair count            */

/* ---------- I/O ---------- */
/* ASCII PCD reader (cat.pcd:1-11, cat_out.pcd:1-11 headers): keeps fields x,y,z, drops the
 * rest, like pcl::fromPCLPointCloud2 at myicp.cpp:24,26.  Caller frees *xyz with orc_free. */
int orc_pcd_read_xyz(const char* path, float** xyz, size_t* n);
void orc_free(void* p);

/* ---------- pinned fp32 primitives (Oracle-C definitions) ---------- */
/* d2 = ((dx*dx)+(dy*dy))+(dz*dz), no FMA */
float orc_dist2(const float* a, const float* b);
/* x' = (((T00*x)+(T01*y))+(T02*z))+T03, no FMA; T is 4x4 row-major, last row ignored */
void orc_transform_points(const float T[16], const float* xyz, size_t n, float* out);
void orc_rotate_vectors(const float T[16], const float* v, size_t n, size_t stride_floats, float* out);

/* ---------- exact nearest neighbour (replaces the missing search, myicp.cpp:128-131) ---------- */
/* lowest target index wins ties.  idx[i] = -1 when nt == 0. */
void orc_nn_brute(const float* q, size_t nq, const float* t, size_t nt, int32_t* idx, float* d2);

typedef struct orc_kdtree orc_kdtree;
/* exact kd-tree, leaf size 15 (the pcl::KdTreeFLANN / FLANN KDTreeSingleIndex configuration) */
orc_kdtree* orc_kdtree_build(const float* xyz, size_t n);
void orc_kdtree_free(orc_kdtree* t);
void orc_kdtree_nn(const orc_kdtree* t, const float* q, size_t nq, int32_t* idx, float* d2, int threads);
/* k nearest, ascending (d2, index); rows padded with -1 / inf when n < k */
void orc_kdtree_knn(const orc_kdtree* t, const float* q, size_t nq, int k, int32_t* idx, float* d2, int threads);

/* ---------- normals (MyICP::estimateNormals, myicp.cpp:152-172; PCL NormalEstimation k=10) ---------- */
/* exact kNN-k set (self included), centred covariance, eigenvector of the smallest eigenvalue,
 * flipped so that n.(viewpoint - p) >= 0, curvature = l0/(l0+l1+l2).  gap[i] (optional) receives
 * (l1-l0)/trace, the stability measure used by the parity tests. */
void orc_estimate_normals(const float* xyz, size_t n, int k, const float viewpoint[3],
                          float* normals4, float* gap, int threads);

/* ---------- Oracle-L: literal restatement ---------- */
float orc_eval_diff(const float* src, const float* tgt, size_t n);                    /* func.cpp:19-32 */
void orc_matrix_notation(const float* src, const float* src_n, const float* tgt, const float* tgt_n,
                         size_t n, float* M, float* N, float* c);                     /* func.cpp:43-60 */
void orc_solve_lls(const float* A, const float* b, size_t n, float x[3]);             /* func.cpp:64-73 */
void orc_estimate_transform_symm_literal(const float* src, const float* src_n, const float* tgt,
                                         const float* tgt_n, size_t n, float T[16]);  /* func.cpp:76-102 */
void orc_apply_transform_literal(const float* src, size_t n, const float T[16], float* out); /* func.cpp:104-121 */
/* myicp.cpp:100-150 as written (identity pairing, normals translated, all quirks).
 * normals are N x 3 here (pasteInMatrix drops curvature, func.cpp:5-15). diffs gets 1+iters values. */
int orc_register_symm_literal(const float* src, const float* src_n3, const float* tgt, const float* tgt_n3,
                              size_t n, int max_iters, float diff_threshold, float T_out[16], float* diffs);
/* regist.h:8-72: dst = R*src + T, fp64. returns -1 on size < 3 (the reference only prints). */
int orc_registrate_npoint(const double* src, const double* dst, size_t n, double R[9], double T[3]);

/* ---------- Oracle-C: canonical symmetric ICP ---------- */
/* p, n_p: current (already transformed) source points / rotated normals; q, n_q: their matched
 * target points / normals (already gathered).  normals stride 4.  Sums are about `origin`. */
void orc_symm_accumulate(const float* p, const float* np4, const float* q, const float* nq4, size_t n,
                         const double origin[3], double sums[ORC_NSUMS]);
/* Re-centre raw sums to the pair centroids, solve the 6x6 system (LDL^T, pivot guard
 * d_j <= 1e-6*A_jj -> component dropped, flag set), compose the half-rotation update
 * Tr(qbar) R Tr(t cos) R Tr(-pbar).  Tinc is 4x4 row-major fp64.  returns degenerate flag. */
int orc_solve_compose(const double sums[ORC_NSUMS], const double origin[3], double Tinc[16], double x[6],
                      double* min_pivot_ratio /* optional: min_j d_j/A_jj, 1 = perfectly conditioned */);

typedef struct {
    int max_iters;          /* myicp.cpp:6 default 10 */
    float diff_threshold;   /* myicp.cpp:6 default 1.0f; test is diff > thr (myicp.cpp:123) */
    double rel_diff_tol;    /* 0 = off (reference behaviour); else stop when |d(k-1)-d(k)| <= tol*d(k-1) */
    int identity_pairs;     /* 1 = pair row i with row i (what the reference actually does) */
    int threads;
    int use_brute;          /* 1 = brute-force NN instead of the kd-tree */
    float max_pair_dist;    /* 0 = off; else pairs with ||p-q|| > this are rejected (next-row f3) */
    float min_normal_dot;   /* <= -1 = off; else pairs with n_p . n_q < this are rejected (next-row f3) */
} orc_params;

typedef struct {
    int iters;              /* number of solves applied */
    int converged;          /* 1 diff<=thr, 2 rel tol, 0 hit max_iters */
    int degenerate;
    double diff, rmse;
    size_t pairs;
    double min_pivot_ratio; /* min over iterations of min_j d_j/A_jj (weak-constraint indicator) */
} orc_stats;

/* trace (optional): per search k = 0..iters: Tf (16 floats, the fp32 transform actually applied
 * to the points), diff.  trace_T must hold 16*(max_iters+1) floats, trace_diff max_iters+1 doubles.
 * idx_out (optional, ns) receives the final correspondences. */
int orc_register_symm(const float* src, const float* src_n4, size_t ns,
                      const float* tgt, const float* tgt_n4, size_t nt,
                      const orc_params* prm, float T_out[16], orc_stats* st,
                      float* trace_T, double* trace_diff, int32_t* idx_out);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
