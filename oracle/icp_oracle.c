/*
 * icp_oracle.c -- CPU ORACLE (test infrastructure; see icp_oracle.h for scope and pinning).
 * Compile with:  gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC
 * Every function cites the reference file:line it restates (paths relative to the
 * reference repository root, ICP/...).
 */
#include "icp_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * ASCII PCD reader.  Header grammar as in ICP/cat.pcd:1-11 and ICP/cat_out.pcd:1-11
 * (FIELDS / SIZE / TYPE / COUNT / WIDTH / HEIGHT / VIEWPOINT / POINTS / DATA ascii).
 * Like myicp.cpp:22-26 (PCDReader::read + fromPCLPointCloud2<PointXYZ>) only x,y,z survive.
 * ------------------------------------------------------------------------------------------ */
int orc_pcd_read_xyz(const char* path, float** xyz_out, size_t* n_out) {
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    char line[4096];
    int nfields = 0, ix = -1, iy = -1, iz = -1;
    size_t npts = 0;
    int have_data = 0;
    while (fgets(line, sizeof line, f)) {
        if (line[0] == '#') continue;
        if (!strncmp(line, "FIELDS", 6)) {
            char* save = NULL;
            char* tok = strtok_r(line + 6, " \t\r\n", &save);
            while (tok) {
                if (!strcmp(tok, "x")) ix = nfields;
                else if (!strcmp(tok, "y")) iy = nfields;
                else if (!strcmp(tok, "z")) iz = nfields;
                nfields++;
                tok = strtok_r(NULL, " \t\r\n", &save);
            }
        } else if (!strncmp(line, "POINTS", 6)) {
            npts = (size_t)strtoull(line + 6, NULL, 10);
        } else if (!strncmp(line, "DATA", 4)) {
            have_data = (strstr(line, "ascii") != NULL);
            break;
        }
    }
    if (!have_data || ix < 0 || iy < 0 || iz < 0 || nfields > 64) { fclose(f); return -2; }
    float* xyz = (float*)malloc(sizeof(float) * 3 * (npts ? npts : 1));
    size_t got = 0;
    while (got < npts && fgets(line, sizeof line, f)) {
        char* save = NULL;
        char* tok = strtok_r(line, " \t\r\n", &save);
        int k = 0, seen = 0;
        float v[3] = {0, 0, 0};
        while (tok && k < nfields) {
            if (k == ix) { v[0] = strtof(tok, NULL); seen++; }
            else if (k == iy) { v[1] = strtof(tok, NULL); seen++; }
            else if (k == iz) { v[2] = strtof(tok, NULL); seen++; }
            k++;
            tok = strtok_r(NULL, " \t\r\n", &save);
        }
        if (seen != 3) continue;
        xyz[3 * got + 0] = v[0]; xyz[3 * got + 1] = v[1]; xyz[3 * got + 2] = v[2];
        got++;
    }
    fclose(f);
    if (got != npts) { free(xyz); return -3; }
    *xyz_out = xyz; *n_out = npts;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Pinned fp32 primitives (Oracle-C definitions 1 and 3 of SURVEY.md 8c).
 * ------------------------------------------------------------------------------------------ */
float orc_dist2(const float* a, const float* b) {
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    float s = xx + yy;
    return s + zz;
}

void orc_transform_points(const float T[16], const float* xyz, size_t n, float* out) {
    for (size_t i = 0; i < n; i++) {
        float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        for (int r = 0; r < 3; r++) {
            float a = T[4 * r] * x, b = T[4 * r + 1] * y, c = T[4 * r + 2] * z;
            float s = a + b;
            s = s + c;
            out[3 * i + r] = s + T[4 * r + 3];
        }
    }
}

void orc_rotate_vectors(const float T[16], const float* v, size_t n, size_t stride, float* out) {
    for (size_t i = 0; i < n; i++) {
        float x = v[stride * i], y = v[stride * i + 1], z = v[stride * i + 2];
        for (int r = 0; r < 3; r++) {
            float a = T[4 * r] * x, b = T[4 * r + 1] * y, c = T[4 * r + 2] * z;
            float s = a + b;
            out[stride * i + r] = s + c;
        }
        for (size_t k = 3; k < stride; k++) out[stride * i + k] = v[stride * i + k];
    }
}

/* ------------------------------------------------------------------------------------------
 * Exact nearest neighbour.  The reference has none (myicp.cpp:128-131 "todo indices = ...");
 * the north-star defines it: argmin of the fp32 d2 above, lowest target index on ties.
 * ------------------------------------------------------------------------------------------ */
void orc_nn_brute(const float* q, size_t nq, const float* t, size_t nt, int32_t* idx, float* d2) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)nq; i++) {
        float best = INFINITY;
        int32_t bi = -1;
        for (size_t j = 0; j < nt; j++) {
            float d = orc_dist2(q + 3 * i, t + 3 * j);
            if (d < best) { best = d; bi = (int32_t)j; } /* ascending j: strict < keeps lowest index */
        }
        idx[i] = bi;
        if (d2) d2[i] = best;
    }
}

/* ---- kd-tree: exact, leaf size 15, tight per-node boxes, nearer child first ---- */
#define KD_LEAF 15
typedef struct {
    float lo[3], hi[3];
    int32_t left, right;  /* children, -1 for leaf */
    int32_t start, count; /* range in perm */
} kd_node;

struct orc_kdtree {
    const float* xyz;
    size_t n;
    int32_t* perm;
    kd_node* nodes;
    int32_t n_nodes, cap_nodes;
};

static int kd_less(const float* xyz, int32_t a, int32_t b, int dim) {
    float va = xyz[3 * a + dim], vb = xyz[3 * b + dim];
    return (va < vb) || (va == vb && a < b);
}

/* quickselect so that perm[k] holds the k-th element under kd_less */
static void kd_select(const float* xyz, int32_t* perm, int lo, int hi, int k, int dim) {
    while (hi - lo > 1) {
        int mid = lo + (hi - lo) / 2;
        /* median of three pivot */
        int a = lo, b = mid, c = hi - 1;
        if (kd_less(xyz, perm[b], perm[a], dim)) { int32_t t = perm[a]; perm[a] = perm[b]; perm[b] = t; }
        if (kd_less(xyz, perm[c], perm[a], dim)) { int32_t t = perm[a]; perm[a] = perm[c]; perm[c] = t; }
        if (kd_less(xyz, perm[c], perm[b], dim)) { int32_t t = perm[b]; perm[b] = perm[c]; perm[c] = t; }
        int32_t pivot = perm[b];
        { int32_t t = perm[b]; perm[b] = perm[hi - 1]; perm[hi - 1] = t; }
        int store = lo;
        for (int i = lo; i < hi - 1; i++) {
            if (kd_less(xyz, perm[i], pivot, dim)) {
                int32_t t = perm[i]; perm[i] = perm[store]; perm[store] = t;
                store++;
            }
        }
        { int32_t t = perm[store]; perm[store] = perm[hi - 1]; perm[hi - 1] = t; }
        if (k == store) return;
        if (k < store) hi = store; else lo = store + 1;
    }
}

static int32_t kd_build_rec(orc_kdtree* t, int start, int count) {
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes * 2 + 16;
        t->nodes = (kd_node*)realloc(t->nodes, sizeof(kd_node) * (size_t)t->cap_nodes);
    }
    int32_t id = t->n_nodes++;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = start; i < start + count; i++) {
        const float* p = t->xyz + 3 * (size_t)t->perm[i];
        for (int d = 0; d < 3; d++) {
            if (p[d] < lo[d]) lo[d] = p[d];
            if (p[d] > hi[d]) hi[d] = p[d];
        }
    }
    int32_t left = -1, right = -1;
    if (count > KD_LEAF) {
        int dim = 0;
        float ext = hi[0] - lo[0];
        if (hi[1] - lo[1] > ext) { ext = hi[1] - lo[1]; dim = 1; }
        if (hi[2] - lo[2] > ext) { ext = hi[2] - lo[2]; dim = 2; }
        int mid = start + count / 2;
        kd_select(t->xyz, t->perm, start, start + count, mid, dim);
        left = kd_build_rec(t, start, mid - start);
        right = kd_build_rec(t, mid, start + count - mid);
    }
    kd_node* nd = &t->nodes[id];
    memcpy(nd->lo, lo, sizeof lo);
    memcpy(nd->hi, hi, sizeof hi);
    nd->left = left; nd->right = right; nd->start = start; nd->count = count;
    return id;
}

orc_kdtree* orc_kdtree_build(const float* xyz, size_t n) {
    orc_kdtree* t = (orc_kdtree*)calloc(1, sizeof *t);
    t->xyz = xyz; t->n = n;
    t->perm = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) t->perm[i] = (int32_t)i;
    if (n) kd_build_rec(t, 0, (int)n);
    return t;
}

void orc_kdtree_free(orc_kdtree* t) {
    if (!t) return;
    free(t->perm); free(t->nodes); free(t);
}

/* lower bound of the fp32 d2 from q to any point inside the box; same operation order as
 * orc_dist2 so that rounding is monotone and the bound never exceeds a member's d2 */
static float kd_box_d2(const kd_node* nd, const float* q) {
    float e[3];
    for (int d = 0; d < 3; d++) {
        float a = nd->lo[d] - q[d], b = q[d] - nd->hi[d];
        float m = a > b ? a : b;
        e[d] = m > 0.f ? m : 0.f;
    }
    float xx = e[0] * e[0], yy = e[1] * e[1], zz = e[2] * e[2];
    float s = xx + yy;
    return s + zz;
}

static void kd_nn_rec(const orc_kdtree* t, int32_t id, const float* q, float* best, int32_t* bi) {
    const kd_node* nd = &t->nodes[id];
    if (nd->left < 0) {
        for (int i = nd->start; i < nd->start + nd->count; i++) {
            int32_t j = t->perm[i];
            float d = orc_dist2(q, t->xyz + 3 * (size_t)j);
            if (d < *best || (d == *best && j < *bi)) { *best = d; *bi = j; }
        }
        return;
    }
    float dl = kd_box_d2(&t->nodes[nd->left], q), dr = kd_box_d2(&t->nodes[nd->right], q);
    int32_t first = nd->left, second = nd->right;
    float d1 = dl, d2 = dr;
    if (dr < dl) { first = nd->right; second = nd->left; d1 = dr; d2 = dl; }
    if (d1 <= *best) kd_nn_rec(t, first, q, best, bi);   /* <= : an equal-d2 lower index may hide there */
    if (d2 <= *best) kd_nn_rec(t, second, q, best, bi);
}

void orc_kdtree_nn(const orc_kdtree* t, const float* q, size_t nq, int32_t* idx, float* d2, int threads) {
    if (threads <= 0) threads = orc_num_threads();
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads)
    for (long long i = 0; i < (long long)nq; i++) {
        float best = INFINITY;
        int32_t bi = INT32_MAX;
        if (t->n) kd_nn_rec(t, 0, q + 3 * i, &best, &bi);
        idx[i] = t->n ? bi : -1;
        if (d2) d2[i] = best;
    }
}

typedef struct { float* d; int32_t* id; int k; int cnt; } knn_list;

static int pair_less(float da, int32_t ia, float db, int32_t ib) { return da < db || (da == db && ia < ib); }

static void knn_insert(knn_list* L, float d, int32_t j) {
    if (L->cnt == L->k && !pair_less(d, j, L->d[L->k - 1], L->id[L->k - 1])) return;
    int pos = (L->cnt < L->k) ? L->cnt++ : L->k - 1;
    while (pos > 0 && pair_less(d, j, L->d[pos - 1], L->id[pos - 1])) {
        L->d[pos] = L->d[pos - 1]; L->id[pos] = L->id[pos - 1];
        pos--;
    }
    L->d[pos] = d; L->id[pos] = j;
}

static void kd_knn_rec(const orc_kdtree* t, int32_t id, const float* q, knn_list* L) {
    const kd_node* nd = &t->nodes[id];
    if (nd->left < 0) {
        for (int i = nd->start; i < nd->start + nd->count; i++) {
            int32_t j = t->perm[i];
            knn_insert(L, orc_dist2(q, t->xyz + 3 * (size_t)j), j);
        }
        return;
    }
    float dl = kd_box_d2(&t->nodes[nd->left], q), dr = kd_box_d2(&t->nodes[nd->right], q);
    int32_t first = nd->left, second = nd->right;
    float d1 = dl, d2 = dr;
    if (dr < dl) { first = nd->right; second = nd->left; d1 = dr; d2 = dl; }
    if (L->cnt < L->k || d1 <= L->d[L->k - 1]) kd_knn_rec(t, first, q, L);
    if (L->cnt < L->k || d2 <= L->d[L->k - 1]) kd_knn_rec(t, second, q, L);
}

void orc_kdtree_knn(const orc_kdtree* t, const float* q, size_t nq, int k, int32_t* idx, float* d2, int threads) {
    if (threads <= 0) threads = orc_num_threads();
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
    for (long long i = 0; i < (long long)nq; i++) {
        float dbuf[64];
        int32_t ibuf[64];
        knn_list L = {dbuf, ibuf, k > 64 ? 64 : k, 0};
        if (t->n) kd_knn_rec(t, 0, q + 3 * i, &L);
        for (int j = 0; j < k; j++) {
            idx[(size_t)i * k + j] = j < L.cnt ? ibuf[j] : -1;
            if (d2) d2[(size_t)i * k + j] = j < L.cnt ? dbuf[j] : INFINITY;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * 3x3 symmetric eigen-decomposition, cyclic Jacobi, fp64.  a is row-major symmetric;
 * on exit w[] ascending eigenvalues, v[3*r+c] = component r of eigenvector c.
 * Stands in for pcl::eigen33 (called inside pcl::NormalEstimation, myicp.cpp:165,167).
 * ------------------------------------------------------------------------------------------ */
static void jacobi_eig3(const double a_in[9], double w[3], double v[9]) {
    double a[9];
    memcpy(a, a_in, sizeof a);
    for (int i = 0; i < 9; i++) v[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
        double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
        if (off <= 1e-300 || off <= 1e-32 * diag) break;
        static const int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2};
        for (int r = 0; r < 3; r++) {
            int p = P[r], q = Q[r];
            double apq = a[3 * p + q];
            if (apq == 0.0) continue;
            double theta = (a[3 * q + q] - a[3 * p + p]) / (2.0 * apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; k++) { /* A <- A J */
                double akp = a[3 * k + p], akq = a[3 * k + q];
                a[3 * k + p] = c * akp - s * akq;
                a[3 * k + q] = s * akp + c * akq;
            }
            for (int k = 0; k < 3; k++) { /* A <- J^T A */
                double apk = a[3 * p + k], aqk = a[3 * q + k];
                a[3 * p + k] = c * apk - s * aqk;
                a[3 * q + k] = s * apk + c * aqk;
            }
            for (int k = 0; k < 3; k++) {
                double vkp = v[3 * k + p], vkq = v[3 * k + q];
                v[3 * k + p] = c * vkp - s * vkq;
                v[3 * k + q] = s * vkp + c * vkq;
            }
        }
    }
    w[0] = a[0]; w[1] = a[4]; w[2] = a[8];
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2 - i; j++)
            if (w[j] > w[j + 1]) {
                double t = w[j]; w[j] = w[j + 1]; w[j + 1] = t;
                for (int k = 0; k < 3; k++) { double u = v[3 * k + j]; v[3 * k + j] = v[3 * k + j + 1]; v[3 * k + j + 1] = u; }
            }
}

/* MyICP::estimateNormals (myicp.cpp:152-172): NormalEstimation with setKSearch(10) (:160),
 * default viewpoint (0,0,0).  PCL internals restated from the published algorithm
 * (computePointNormal -> covariance -> eigen33 -> flipNormalTowardsViewpoint); PARITY UNPINNED. */
void orc_estimate_normals(const float* xyz, size_t n, int k, const float vp[3], float* normals4, float* gap,
                          int threads) {
    if (threads <= 0) threads = orc_num_threads();
    if (k > 64) k = 64;
    orc_kdtree* t = orc_kdtree_build(xyz, n);
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
    for (long long i = 0; i < (long long)n; i++) {
        float dbuf[64];
        int32_t ibuf[64];
        knn_list L = {dbuf, ibuf, k, 0};
        kd_knn_rec(t, 0, xyz + 3 * i, &L);
        int m = L.cnt;
        double c[3] = {0, 0, 0};
        for (int j = 0; j < m; j++)
            for (int d = 0; d < 3; d++) c[d] += xyz[3 * (size_t)ibuf[j] + d];
        for (int d = 0; d < 3; d++) c[d] /= (double)m;
        double cov[9] = {0};
        for (int j = 0; j < m; j++) {
            double e[3];
            for (int d = 0; d < 3; d++) e[d] = (double)xyz[3 * (size_t)ibuf[j] + d] - c[d];
            for (int r = 0; r < 3; r++)
                for (int s = 0; s < 3; s++) cov[3 * r + s] += e[r] * e[s];
        }
        for (int r = 0; r < 9; r++) cov[r] /= (double)m;
        double w[3], v[9];
        jacobi_eig3(cov, w, v);
        double nx = v[0], ny = v[3], nz = v[6];
        double tr = w[0] + w[1] + w[2];
        double vx = (double)vp[0] - xyz[3 * i], vy = (double)vp[1] - xyz[3 * i + 1], vz = (double)vp[2] - xyz[3 * i + 2];
        if (vx * nx + vy * ny + vz * nz < 0) { nx = -nx; ny = -ny; nz = -nz; }
        normals4[4 * i + 0] = (float)nx; normals4[4 * i + 1] = (float)ny; normals4[4 * i + 2] = (float)nz;
        normals4[4 * i + 3] = tr > 0 ? (float)fabs(w[0] / tr) : 0.f;
        if (gap) gap[i] = tr > 0 ? (float)((w[1] - w[0]) / tr) : 0.f;
    }
    orc_kdtree_free(t);
}

/* ------------------------------------------------------------------------------------------
 * Oracle-L: literal restatement of ICP/func.cpp and ICP/myicp.cpp.
 * ------------------------------------------------------------------------------------------ */

/* func.cpp:19-32: diff += (src.row(i) - tgt.row(i)).norm(), fp32, sequential in i. */
float orc_eval_diff(const float* src, const float* tgt, size_t n) {
    float diff = 0.f;
    for (size_t i = 0; i < n; i++) {
        float d = orc_dist2(src + 3 * i, tgt + 3 * i);
        diff += sqrtf(d);
    }
    return diff;
}

/* func.cpp:43-60: n = n_src + n_tgt (:51), M.row = (p+q) x n (:54), N.row = n (:56),
 * c = (p-q).n (:58).  p, q raw (NOT centred) exactly as coded. */
void orc_matrix_notation(const float* src, const float* src_n, const float* tgt, const float* tgt_n, size_t n,
                         float* M, float* N, float* c) {
    for (size_t i = 0; i < n; i++) {
        const float *p = src + 3 * i, *q = tgt + 3 * i;
        float nn[3] = {src_n[3 * i] + tgt_n[3 * i], src_n[3 * i + 1] + tgt_n[3 * i + 1], src_n[3 * i + 2] + tgt_n[3 * i + 2]};
        float s[3] = {p[0] + q[0], p[1] + q[1], p[2] + q[2]};
        float d[3] = {p[0] - q[0], p[1] - q[1], p[2] - q[2]};
        M[3 * i + 0] = s[1] * nn[2] - s[2] * nn[1];
        M[3 * i + 1] = s[2] * nn[0] - s[0] * nn[2];
        M[3 * i + 2] = s[0] * nn[1] - s[1] * nn[0];
        N[3 * i + 0] = nn[0]; N[3 * i + 1] = nn[1]; N[3 * i + 2] = nn[2];
        c[i] = (d[0] * nn[0] + d[1] * nn[1]) + d[2] * nn[2];
    }
}

/* func.cpp:64-73: x = V S^-1 U^T b from JacobiSVD(A, ThinU|FullV), no rank truncation.
 * Restated as a one-sided (Hestenes) Jacobi SVD in fp64 of the N x 3 matrix:
 * A V = U S  =>  x = sum_j v_j (a'_j . b) / s_j^2 with a'_j the rotated columns. */
void orc_solve_lls(const float* A, const float* b, size_t n, float x[3]) {
    double* a = (double*)malloc(sizeof(double) * 3 * (n ? n : 1));
    for (size_t i = 0; i < 3 * n; i++) a[i] = A[i];
    double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int sweep = 0; sweep < 60; sweep++) {
        int rotated = 0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double al = 0, be = 0, ga = 0;
                for (size_t i = 0; i < n; i++) {
                    al += a[3 * i + p] * a[3 * i + p];
                    be += a[3 * i + q] * a[3 * i + q];
                    ga += a[3 * i + p] * a[3 * i + q];
                }
                if (fabs(ga) <= 1e-15 * sqrt(al * be) || ga == 0.0) continue;
                rotated = 1;
                double zeta = (be - al) / (2.0 * ga);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (size_t i = 0; i < n; i++) {
                    double ap = a[3 * i + p], aq = a[3 * i + q];
                    a[3 * i + p] = c * ap - s * aq;
                    a[3 * i + q] = s * ap + c * aq;
                }
                for (int k = 0; k < 3; k++) {
                    double vp = V[3 * k + p], vq = V[3 * k + q];
                    V[3 * k + p] = c * vp - s * vq;
                    V[3 * k + q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double xs[3] = {0, 0, 0};
    for (int j = 0; j < 3; j++) {
        double s2 = 0, ub = 0;
        for (size_t i = 0; i < n; i++) { s2 += a[3 * i + j] * a[3 * i + j]; ub += a[3 * i + j] * (double)b[i]; }
        double coef = ub / s2; /* unguarded like func.cpp:70 (Sigma^-1) */
        for (int k = 0; k < 3; k++) xs[k] += V[3 * k + j] * coef;
    }
    for (int k = 0; k < 3; k++) x[k] = (float)xs[k];
    free(a);
}

static void mat4_mul(const double A[16], const double B[16], double C[16]) {
    double R[16];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += A[4 * r + k] * B[4 * k + c];
            R[4 * r + c] = s;
        }
    memcpy(C, R, sizeof R);
}
static void mat4_identity(double A[16]) {
    for (int i = 0; i < 16; i++) A[i] = (i % 5 == 0) ? 1.0 : 0.0;
}
static void mat4_translation(const double t[3], double A[16]) {
    mat4_identity(A);
    A[3] = t[0]; A[7] = t[1]; A[11] = t[2];
}
/* Rodrigues rotation about unit axis u by angle th (Eigen::AngleAxisf, func.cpp:96,98) */
static void mat4_rotation(const double u[3], double th, double A[16]) {
    double c = cos(th), s = sin(th), C = 1 - c;
    mat4_identity(A);
    A[0] = c + u[0] * u[0] * C;         A[1] = u[0] * u[1] * C - u[2] * s;  A[2] = u[0] * u[2] * C + u[1] * s;
    A[4] = u[1] * u[0] * C + u[2] * s;  A[5] = c + u[1] * u[1] * C;         A[6] = u[1] * u[2] * C - u[0] * s;
    A[8] = u[2] * u[0] * C - u[1] * s;  A[9] = u[2] * u[1] * C + u[0] * s;  A[10] = c + u[2] * u[2] * C;
}

/* func.cpp:76-102 as coded: centroids (:85), t_ = tgt_mean - src_mean (:86),
 * a_ = LLS(M, -(N t_ + c)) (:87), t_ = LLS(N, -(M a_ + c)) (:88), theta = atan|a_| (:93),
 * then Eigen's translate()/rotate() POST-multiply (:95-99), i.e. the matrix product
 * Tr(-src_mean) R Tr(t_ cos) R Tr(tgt_mean) -- the reverse of the comment at :94. */
void orc_estimate_transform_symm_literal(const float* src, const float* src_n, const float* tgt, const float* tgt_n,
                                         size_t n, float T[16]) {
    float* M = (float*)malloc(sizeof(float) * 3 * n);
    float* N = (float*)malloc(sizeof(float) * 3 * n);
    float* c = (float*)malloc(sizeof(float) * n);
    float* rhs = (float*)malloc(sizeof(float) * n);
    orc_matrix_notation(src, src_n, tgt, tgt_n, n, M, N, c);
    double sm[3] = {0, 0, 0}, tm[3] = {0, 0, 0};
    for (size_t i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) { sm[d] += src[3 * i + d]; tm[d] += tgt[3 * i + d]; }
    float src_mean[3], tgt_mean[3], a_[3], t_[3];
    for (int d = 0; d < 3; d++) {
        src_mean[d] = (float)(sm[d] / (double)n); tgt_mean[d] = (float)(tm[d] / (double)n);
        t_[d] = tgt_mean[d] - src_mean[d];
    }
    for (size_t i = 0; i < n; i++)
        rhs[i] = -((N[3 * i] * t_[0] + N[3 * i + 1] * t_[1] + N[3 * i + 2] * t_[2]) + c[i]);
    orc_solve_lls(M, rhs, n, a_);
    for (size_t i = 0; i < n; i++)
        rhs[i] = -((M[3 * i] * a_[0] + M[3 * i + 1] * a_[1] + M[3 * i + 2] * a_[2]) + c[i]);
    orc_solve_lls(N, rhs, n, t_);
    double an = sqrt((double)a_[0] * a_[0] + (double)a_[1] * a_[1] + (double)a_[2] * a_[2]);
    double theta = atan(an);
    double u[3] = {a_[0] / an, a_[1] / an, a_[2] / an}; /* unguarded like func.cpp:96 */
    double A[16], B[16], Tm[16];
    double neg_src[3] = {-src_mean[0], -src_mean[1], -src_mean[2]};
    double tc[3] = {t_[0] * cos(theta), t_[1] * cos(theta), t_[2] * cos(theta)};
    double tgm[3] = {tgt_mean[0], tgt_mean[1], tgt_mean[2]};
    mat4_translation(neg_src, Tm);
    mat4_rotation(u, theta, A); mat4_mul(Tm, A, Tm);
    mat4_translation(tc, B);    mat4_mul(Tm, B, Tm);
    mat4_mul(Tm, A, Tm);
    mat4_translation(tgm, B);   mat4_mul(Tm, B, Tm);
    for (int i = 0; i < 16; i++) T[i] = (float)Tm[i];
    free(M); free(N); free(c); free(rhs);
}

/* func.cpp:104-121: pad with w=1 (:111-115), (T * med^T)^T (:118), left 3 columns (:120). */
void orc_apply_transform_literal(const float* src, size_t n, const float T[16], float* out) {
    for (size_t i = 0; i < n; i++) {
        float x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
        float o[3];
        for (int r = 0; r < 3; r++) o[r] = T[4 * r] * x + T[4 * r + 1] * y + T[4 * r + 2] * z + T[4 * r + 3] * 1.0f;
        out[3 * i] = o[0]; out[3 * i + 1] = o[1]; out[3 * i + 2] = o[2];
    }
}

/* myicp.cpp:100-150 as written: identity pairing (:128-131), applyTransform on the xyz AND
 * on the normals (:136-137, so normals are translated), transform = incre * transform (:138),
 * while (diff > thr && iters++ < max_iters) (:123). */
int orc_register_symm_literal(const float* src, const float* src_n3, const float* tgt, const float* tgt_n3, size_t n,
                              int max_iters, float thr, float T_out[16], float* diffs) {
    float* s = (float*)malloc(sizeof(float) * 3 * n);
    float* sn = (float*)malloc(sizeof(float) * 3 * n);
    memcpy(s, src, sizeof(float) * 3 * n);
    memcpy(sn, src_n3, sizeof(float) * 3 * n);
    double T[16];
    mat4_identity(T);
    int iters = 0, k = 0;
    float diff = orc_eval_diff(s, tgt, n);
    if (diffs) diffs[k++] = diff;
    while (diff > thr && iters++ < max_iters) {
        float inc[16];
        orc_estimate_transform_symm_literal(s, sn, tgt, tgt_n3, n, inc);
        orc_apply_transform_literal(s, n, inc, s);
        orc_apply_transform_literal(sn, n, inc, sn);
        double incd[16];
        for (int i = 0; i < 16; i++) incd[i] = inc[i];
        mat4_mul(incd, T, T);
        for (int i = 0; i < 16; i++) T[i] = (double)(float)T[i]; /* Affine3f storage */
        diff = orc_eval_diff(s, tgt, n);
        if (diffs) diffs[k++] = diff;
    }
    for (int i = 0; i < 16; i++) T_out[i] = (float)T[i];
    free(s); free(sn);
    /* `iters++ < max_iters` leaves iters = max_iters+1 after the failing test; report trips */
    return k - 1;
}

/* one-sided Jacobi SVD of a 3x3 (fp64): A = U diag(s) V^T, s descending, U completed to a
 * full orthonormal basis when rank-deficient.  Stands in for cv::SVDecomp at regist.h:55. */
static void svd3(const double A[9], double U[9], double s[3], double V[9]) {
    double a[9];
    memcpy(a, A, sizeof a);
    for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
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
    for (int j = 0; j < 3; j++) s[j] = sqrt(a[j] * a[j] + a[3 + j] * a[3 + j] + a[6 + j] * a[6 + j]);
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2 - i; j++)
            if (s[ord[j]] < s[ord[j + 1]]) { int t = ord[j]; ord[j] = ord[j + 1]; ord[j + 1] = t; }
    double ss[3], Vs[9], Us[9];
    for (int j = 0; j < 3; j++) {
        ss[j] = s[ord[j]];
        for (int i = 0; i < 3; i++) {
            Vs[3 * i + j] = V[3 * i + ord[j]];
            Us[3 * i + j] = ss[j] > 0 ? a[3 * i + ord[j]] / ss[j] : 0.0;
        }
    }
    double tol = 1e-13 * (ss[0] > 0 ? ss[0] : 1.0);
    if (ss[0] <= tol) {
        for (int i = 0; i < 9; i++) Us[i] = (i % 4 == 0) ? 1.0 : 0.0;
    } else {
        if (ss[1] <= tol) { /* pick any unit vector orthogonal to u0 */
            double u0[3] = {Us[0], Us[3], Us[6]};
            int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
            double e[3] = {0, 0, 0};
            e[m] = 1.0;
            double dot = u0[m];
            double w[3] = {e[0] - dot * u0[0], e[1] - dot * u0[1], e[2] - dot * u0[2]};
            double nw = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            Us[1] = w[0] / nw; Us[4] = w[1] / nw; Us[7] = w[2] / nw;
        }
        if (ss[2] <= tol) { /* u2 = u0 x u1 */
            Us[2] = Us[3] * Us[7] - Us[6] * Us[4];
            Us[5] = Us[6] * Us[1] - Us[0] * Us[7];
            Us[8] = Us[0] * Us[4] - Us[3] * Us[1];
        }
    }
    memcpy(U, Us, sizeof Us); memcpy(V, Vs, sizeof Vs); memcpy(s, ss, sizeof ss);
}

static double det3(const double m[9]) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

/* regist.h:8-72.  Centroids by sequential fp64 sums (:16-35), centred 3xN matrices (:36-50),
 * S = src~ dst~^T (:52), SVD S = U W Vt (:55), det(U Vt) (:57-58), R = Vt^T diag(1,1,det) U^T (:60-63),
 * T = dstbar - R srcbar (:64-69).  Net effect dst = R src + T (the comment at :7 says the opposite). */
int orc_registrate_npoint(const double* src, const double* dst, size_t n, double R[9], double T[3]) {
    if (n < 3) return -1;
    double sc[3] = {0, 0, 0}, dc[3] = {0, 0, 0};
    for (size_t i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) { sc[d] += src[3 * i + d]; dc[d] += dst[3 * i + d]; }
    for (int d = 0; d < 3; d++) { sc[d] /= (double)n; dc[d] /= (double)n; }
    double S[9] = {0};
    for (size_t i = 0; i < n; i++) {
        double a[3], b[3];
        for (int d = 0; d < 3; d++) { a[d] = src[3 * i + d] - sc[d]; b[d] = dst[3 * i + d] - dc[d]; }
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) S[3 * r + c] += a[r] * b[c];
    }
    double U[9], W[3], V[9];
    svd3(S, U, W, V);
    /* matTemp = U * Vt */
    double UVt[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += U[3 * r + k] * V[3 * c + k];
            UVt[3 * r + c] = s;
        }
    double det = det3(UVt);
    double Mdiag[3] = {1, 1, det};
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += V[3 * r + k] * Mdiag[k] * U[3 * c + k];
            R[3 * r + c] = s;
        }
    for (int r = 0; r < 3; r++) T[r] = dc[r] - (sc[0] * R[3 * r] + sc[1] * R[3 * r + 1] + sc[2] * R[3 * r + 2]);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Oracle-C: canonical symmetric ICP (SURVEY.md 8c definition 3; Rusinkiewicz 2019 as the
 * comments at func.cpp:53,57,94 intend).
 * ------------------------------------------------------------------------------------------ */
void orc_symm_accumulate(const float* p, const float* np4, const float* q, const float* nq4, size_t n,
                         const double origin[3], double S[ORC_NSUMS]) {
    for (int i = 0; i < ORC_NSUMS; i++) S[i] = 0;
    for (size_t i = 0; i < n; i++) {
        /* n = n_p + n_q in fp32 like func.cpp:51 */
        float nf[3] = {np4[4 * i] + nq4[4 * i], np4[4 * i + 1] + nq4[4 * i + 1], np4[4 * i + 2] + nq4[4 * i + 2]};
        double nn[3] = {nf[0], nf[1], nf[2]};
        double pp[3], qq[3], s[3], d[3];
        for (int k = 0; k < 3; k++) {
            pp[k] = (double)p[3 * i + k] - origin[k];
            qq[k] = (double)q[3 * i + k] - origin[k];
            s[k] = pp[k] + qq[k];
            d[k] = pp[k] - qq[k];
        }
        double m[3] = {s[1] * nn[2] - s[2] * nn[1], s[2] * nn[0] - s[0] * nn[2], s[0] * nn[1] - s[1] * nn[0]};
        double r = d[0] * nn[0] + d[1] * nn[1] + d[2] * nn[2];
        S[0] += m[0] * m[0]; S[1] += m[0] * m[1]; S[2] += m[0] * m[2];
        S[3] += m[1] * m[1]; S[4] += m[1] * m[2]; S[5] += m[2] * m[2];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) S[6 + 3 * a + b] += m[a] * nn[b];
        S[15] += nn[0] * nn[0]; S[16] += nn[0] * nn[1]; S[17] += nn[0] * nn[2];
        S[18] += nn[1] * nn[1]; S[19] += nn[1] * nn[2]; S[20] += nn[2] * nn[2];
        for (int a = 0; a < 3; a++) { S[21 + a] += m[a] * r; S[24 + a] += nn[a] * r; }
        for (int a = 0; a < 3; a++) { S[27 + a] += pp[a]; S[30 + a] += qq[a]; }
        float d2 = orc_dist2(p + 3 * i, q + 3 * i);
        S[33] += (double)sqrtf(d2);
        S[34] += (double)d2;
        S[35] += r * r;
        S[36] += 1.0;
    }
}

/* skew(v) w = v x w */
static void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

int orc_solve_compose(const double S[ORC_NSUMS], const double origin[3], double Tinc[16], double x[6],
                      double* min_pivot_ratio) {
    double cnt = S[36];
    if (min_pivot_ratio) *min_pivot_ratio = 0.0;
    mat4_identity(Tinc);
    for (int i = 0; i < 6; i++) x[i] = 0;
    if (cnt < 1.0) return 1;
    /* raw blocks about `origin` */
    double A[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};   /* sum m'm'^T */
    double B[9]; for (int i = 0; i < 9; i++) B[i] = S[6 + i];                /* sum m'n^T  */
    double C[9] = {S[15], S[16], S[17], S[16], S[18], S[19], S[17], S[19], S[20]}; /* sum nn^T */
    double bm[3] = {S[21], S[22], S[23]}, bn[3] = {S[24], S[25], S[26]};
    double dp[3], dq[3], de[3], ep[3];
    for (int k = 0; k < 3; k++) {
        dp[k] = S[27 + k] / cnt; dq[k] = S[30 + k] / cnt;  /* pbar-origin, qbar-origin */
        de[k] = dp[k] + dq[k];                              /* delta: p'+q' = p~+q~ + delta */
        ep[k] = dp[k] - dq[k];                              /* eps:   p'-q' = p~-q~ + eps   */
    }
    /* m = m' - delta x n = m' - K n, K = skew(delta);  r = r' - eps.n
     * A~ = A' - B'K^T - K B'^T + K C K^T ; B~ = B' - K C ; C unchanged
     * bm~ = bm' - B' eps - K bn' + K C eps ; bn~ = bn' - C eps */
    double K[9] = {0, -de[2], de[1], de[2], 0, -de[0], -de[1], de[0], 0};
    double BKt[9], KC[9], KCKt[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s1 = 0, s2 = 0;
            for (int k = 0; k < 3; k++) { s1 += B[3 * r + k] * K[3 * c + k]; s2 += K[3 * r + k] * C[3 * k + c]; }
            BKt[3 * r + c] = s1; KC[3 * r + c] = s2;
        }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += KC[3 * r + k] * K[3 * c + k];
            KCKt[3 * r + c] = s;
        }
    double At[9], Bt[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            At[3 * r + c] = A[3 * r + c] - BKt[3 * r + c] - BKt[3 * c + r] + KCKt[3 * r + c];
            Bt[3 * r + c] = B[3 * r + c] - KC[3 * r + c];
        }
    double Ce[3], Be[3], Kbn[3], KCe[3];
    for (int r = 0; r < 3; r++) {
        Ce[r] = C[3 * r] * ep[0] + C[3 * r + 1] * ep[1] + C[3 * r + 2] * ep[2];
        Be[r] = B[3 * r] * ep[0] + B[3 * r + 1] * ep[1] + B[3 * r + 2] * ep[2];
    }
    cross3(de, bn, Kbn);
    cross3(de, Ce, KCe);
    double g[6];
    for (int r = 0; r < 3; r++) {
        g[r] = bm[r] - Be[r] - Kbn[r] + KCe[r];
        g[3 + r] = bn[r] - Ce[r];
    }
    double H[36];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            H[6 * r + c] = At[3 * r + c];
            H[6 * r + 3 + c] = Bt[3 * r + c];
            H[6 * (3 + c) + r] = Bt[3 * r + c];
            H[6 * (3 + r) + 3 + c] = C[3 * r + c];
        }
    /* LDL^T with the relative pivot guard */
    double L[36] = {0}, D[6];
    int dead[6] = {0}, degenerate = 0;
    double minratio = 1.0;
    for (int j = 0; j < 6; j++) {
        double d = H[6 * j + j];
        for (int k = 0; k < j; k++) if (!dead[k]) d -= L[6 * j + k] * L[6 * j + k] * D[k];
        double ratio = H[6 * j + j] > 0 ? d / H[6 * j + j] : 0.0;
        if (ratio < minratio) minratio = ratio;
        if (!(d > 1e-6 * H[6 * j + j]) || !(H[6 * j + j] > 0)) { dead[j] = 1; degenerate = 1; D[j] = 0; continue; }
        D[j] = d;
        for (int i = j + 1; i < 6; i++) {
            double v = H[6 * i + j];
            for (int k = 0; k < j; k++) if (!dead[k]) v -= L[6 * i + k] * L[6 * j + k] * D[k];
            L[6 * i + j] = v / d;
        }
    }
    if (min_pivot_ratio) *min_pivot_ratio = minratio;
    double y[6];
    for (int i = 0; i < 6; i++) {
        double v = -g[i];
        for (int k = 0; k < i; k++) if (!dead[k]) v -= L[6 * i + k] * y[k];
        y[i] = dead[i] ? 0.0 : v;
    }
    for (int i = 0; i < 6; i++) y[i] = dead[i] ? 0.0 : y[i] / D[i];
    for (int i = 5; i >= 0; i--) {
        double v = y[i];
        if (dead[i]) { x[i] = 0; continue; }
        for (int k = i + 1; k < 6; k++) if (!dead[k]) v -= L[6 * k + i] * x[k];
        x[i] = v;
    }
    /* compose: Tr(qbar) R(th,a) Tr(t~ cos th) R(th,a) Tr(-pbar)   (comment at func.cpp:94, cos per :97) */
    double an = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    double th = atan(an);
    double Rm[16];
    if (an > 1e-300) { double u[3] = {x[0] / an, x[1] / an, x[2] / an}; mat4_rotation(u, th, Rm); }
    else mat4_identity(Rm);
    double pbar[3], qbar[3], tc[3];
    for (int k = 0; k < 3; k++) { pbar[k] = -(origin[k] + dp[k]); qbar[k] = origin[k] + dq[k]; tc[k] = x[3 + k] * cos(th); }
    double M1[16], M2[16];
    mat4_translation(qbar, M1);
    mat4_mul(M1, Rm, M1);
    mat4_translation(tc, M2);  mat4_mul(M1, M2, M1);
    mat4_mul(M1, Rm, M1);
    mat4_translation(pbar, M2); mat4_mul(M1, M2, M1);
    memcpy(Tinc, M1, sizeof M1);
    return degenerate;
}

int orc_register_symm(const float* src, const float* src_n4, size_t ns, const float* tgt, const float* tgt_n4,
                      size_t nt, const orc_params* prm, float T_out[16], orc_stats* st, float* trace_T,
                      double* trace_diff, int32_t* idx_out) {
    if (!ns || !nt) return -1;
    if (prm->identity_pairs && ns != nt) return -2; /* func.cpp:21 assert */
    int threads = prm->threads > 0 ? prm->threads : orc_num_threads();
    orc_kdtree* tree = (prm->identity_pairs || prm->use_brute) ? NULL : orc_kdtree_build(tgt, nt);
    float* p = (float*)malloc(sizeof(float) * 3 * ns);
    float* np = (float*)malloc(sizeof(float) * 4 * ns);
    float* q = (float*)malloc(sizeof(float) * 3 * ns);
    float* nq = (float*)malloc(sizeof(float) * 4 * ns);
    float* pk = (float*)malloc(sizeof(float) * 3 * ns);
    float* npk = (float*)malloc(sizeof(float) * 4 * ns);
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * ns);
    float* d2 = (float*)malloc(sizeof(float) * ns);
    double T[16];
    mat4_identity(T);
    /* fixed origin for the raw sums: target bounding-box centre (any fixed point works) */
    double origin[3];
    {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (size_t i = 0; i < nt; i++)
            for (int d = 0; d < 3; d++) {
                if (tgt[3 * i + d] < lo[d]) lo[d] = tgt[3 * i + d];
                if (tgt[3 * i + d] > hi[d]) hi[d] = tgt[3 * i + d];
            }
        for (int d = 0; d < 3; d++) origin[d] = 0.5 * ((double)lo[d] + (double)hi[d]);
    }
    int iters = 0, converged = 0, degenerate = 0;
    double diff = 0, prev_diff = 0, sumd2 = 0, min_ratio = 1.0;
    size_t pairs = 0;
    for (int k = 0;; k++) {
        float Tf[16];
        for (int i = 0; i < 16; i++) Tf[i] = (float)T[i];
        orc_transform_points(Tf, src, ns, p);
        orc_rotate_vectors(Tf, src_n4, ns, 4, np);
        if (prm->identity_pairs) {
            for (size_t i = 0; i < ns; i++) { idx[i] = (int32_t)i; d2[i] = orc_dist2(p + 3 * i, tgt + 3 * i); }
        } else if (prm->use_brute) {
            orc_nn_brute(p, ns, tgt, nt, idx, d2);
        } else {
            orc_kdtree_nn(tree, p, ns, idx, d2, threads);
        }
        /* diff over ALL pairs (evalDiff semantics, func.cpp:19-32 with the indices of :36-40) */
        diff = 0; sumd2 = 0;
        for (size_t i = 0; i < ns; i++) { diff += (double)sqrtf(d2[i]); sumd2 += (double)d2[i]; }
        if (trace_T) memcpy(trace_T + 16 * k, Tf, sizeof Tf);
        if (trace_diff) trace_diff[k] = diff;
        if (!(diff > (double)prm->diff_threshold)) { converged = 1; break; }           /* myicp.cpp:123 */
        if (k >= prm->max_iters) break;                                                /* iters++ < max_iters */
        if (prm->rel_diff_tol > 0 && k > 0 && fabs(prev_diff - diff) <= prm->rel_diff_tol * prev_diff) { converged = 2; break; }
        prev_diff = diff;
        /* gather matched pairs (optionally rejecting far pairs) */
        size_t m = 0;
        float maxd2 = prm->max_pair_dist > 0 ? prm->max_pair_dist * prm->max_pair_dist : INFINITY;
        for (size_t i = 0; i < ns; i++) {
            if (d2[i] > maxd2) continue;
            if (prm->min_normal_dot > -1.0f) {
                const float* a = np + 4 * i;
                const float* b = tgt_n4 + 4 * (size_t)idx[i];
                if (a[0] * b[0] + a[1] * b[1] + a[2] * b[2] < prm->min_normal_dot) continue;
            }
            memcpy(pk + 3 * m, p + 3 * i, 3 * sizeof(float));
            memcpy(npk + 4 * m, np + 4 * i, 4 * sizeof(float));
            memcpy(q + 3 * m, tgt + 3 * (size_t)idx[i], 3 * sizeof(float));
            memcpy(nq + 4 * m, tgt_n4 + 4 * (size_t)idx[i], 4 * sizeof(float));
            m++;
        }
        pairs = m;
        double S[ORC_NSUMS], Tinc[16], x[6];
        orc_symm_accumulate(pk, npk, q, nq, m, origin, S);
        double ratio;
        degenerate |= orc_solve_compose(S, origin, Tinc, x, &ratio);
        if (ratio < min_ratio) min_ratio = ratio;
        mat4_mul(Tinc, T, T);                                                          /* myicp.cpp:138 */
        iters++;
    }
    for (int i = 0; i < 16; i++) T_out[i] = (float)T[i];
    if (idx_out) memcpy(idx_out, idx, sizeof(int32_t) * ns);
    if (st) {
        st->iters = iters; st->converged = converged; st->degenerate = degenerate;
        st->diff = diff; st->rmse = sqrt(sumd2 / (double)ns); st->pairs = pairs;
        st->min_pivot_ratio = min_ratio;
    }
    free(p); free(np); free(q); free(nq); free(pk); free(npk); free(idx); free(d2);
    orc_kdtree_free(tree);
    return 0;
}
