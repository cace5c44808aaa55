#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

Run in the build container only (needs /root/reference and numpy/scipy/cv2):
    python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read only the committed .npy/.npz files.

What is pinned here, independently of oracle/ (none of this imports the oracle):
  * the reference's fixture data (ICP/cat.pcd, ICP/cat_out.pcd, ICP/txt2pcd_bunny1.pcd),
    parsed with plain Python and stored as float32 arrays;
  * the fixture's analytic known answer  cat_out = Rz(pi/4) cat + (2.5,0,0)
    (ICP/main.cpp:44-52, ICP/register-test.cpp:34-39);
  * independent numpy / scipy / OpenCV results for each oracle function:
      evalDiff (func.cpp:19-32) sequential fp32,
      calculateMatrixNotation rows (func.cpp:43-60),
      solveLLS (func.cpp:64-73) via numpy SVD pseudo-inverse,
      estimateTransformSymm + RegisterSymm literal loop (func.cpp:76-102, myicp.cpp:100-150),
      registrateNPoint (regist.h:8-72) via cv2.SVDecomp exactly as the reference calls it,
      exact fp32 brute-force 1-NN (+ OpenCV-bundled FLANN KDTREE_SINGLE, leaf 15: the
      pcl::KdTreeFLANN configuration),
      k=10 PCA normals via scipy cKDTree + numpy eigh (PCL NormalEstimation restatement),
      the canonical joint 6-unknown solve by numpy.linalg.lstsq on the stacked rows + the composition of the
      comment at func.cpp:94 (cat_solve_golden.npz) -- pins oracle and device solve_compose to a third derivation.
"""
import os

import numpy as np

REF = "/root/reference/ICP"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_pcd_xyz(path):
    with open(path) as f:
        lines = f.read().splitlines()
    fields, npts, start = None, None, None
    for i, ln in enumerate(lines):
        if ln.startswith("FIELDS"):
            fields = ln.split()[1:]
        elif ln.startswith("POINTS"):
            npts = int(ln.split()[1])
        elif ln.startswith("DATA"):
            assert "ascii" in ln
            start = i + 1
            break
    ix, iy, iz = fields.index("x"), fields.index("y"), fields.index("z")
    rows = [ln.split() for ln in lines[start:start + npts]]
    return np.array([[np.float32(r[ix]), np.float32(r[iy]), np.float32(r[iz])] for r in rows], np.float32)


def d2_pinned(q, t):
    """((dx*dx)+(dy*dy))+(dz*dz) in fp32, q (nq,3) vs t (nt,3) -> (nq,nt)"""
    dx = q[:, None, 0] - t[None, :, 0]
    dy = q[:, None, 1] - t[None, :, 1]
    dz = q[:, None, 2] - t[None, :, 2]
    return ((dx * dx) + (dy * dy)) + (dz * dz)


def eval_diff_seq(src, tgt):
    diff = np.float32(0)
    for i in range(len(src)):
        d = src[i] - tgt[i]
        diff = np.float32(diff + np.sqrt(np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])))
    return diff


def pca_normals(xyz, k=10, vp=(0, 0, 0)):
    from scipy.spatial import cKDTree
    x64 = xyz.astype(np.float64)
    _, nb = cKDTree(x64).query(x64, k=k)
    P = x64[nb]                                   # (n,k,3)
    c = P.mean(axis=1, keepdims=True)
    E = P - c
    cov = np.einsum("nki,nkj->nij", E, E) / k
    w, v = np.linalg.eigh(cov)
    n = v[:, :, 0]
    flip = np.einsum("ni,ni->n", np.asarray(vp)[None, :] - x64, n) < 0
    n[flip] *= -1
    tr = w.sum(axis=1)
    curv = np.where(tr > 0, np.abs(w[:, 0]) / np.where(tr > 0, tr, 1), 0)
    gap = np.where(tr > 0, (w[:, 1] - w[:, 0]) / np.where(tr > 0, tr, 1), 0)
    return np.concatenate([n, curv[:, None]], axis=1).astype(np.float32), gap.astype(np.float32), nb.astype(np.int32)


def lls_pinv(A, b):
    U, s, Vt = np.linalg.svd(A.astype(np.float64), full_matrices=False)
    return (Vt.T @ np.diag(1.0 / s) @ U.T @ b.astype(np.float64)).astype(np.float32)


def rot(u, th):
    u = np.asarray(u, np.float64)
    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    R = np.eye(4)
    R[:3, :3] = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    return R


def tr(t):
    T = np.eye(4)
    T[:3, 3] = t
    return T


def estimate_literal(src, sn, tgt, tn):
    """func.cpp:76-102 exactly as coded (Eigen translate/rotate post-multiply)."""
    n = sn + tn
    M = np.cross(src + tgt, n).astype(np.float32)
    c = np.einsum("ni,ni->n", src - tgt, n).astype(np.float32)
    sm = src.mean(axis=0, dtype=np.float64).astype(np.float32)
    tm = tgt.mean(axis=0, dtype=np.float64).astype(np.float32)
    t_ = tm - sm
    a_ = lls_pinv(M, -(n @ t_ + c))
    t_ = lls_pinv(n, -(M @ a_ + c))
    an = np.linalg.norm(a_.astype(np.float64))
    th = np.arctan(an)
    u = a_.astype(np.float64) / an
    T = tr(-sm.astype(np.float64)) @ rot(u, th) @ tr(t_.astype(np.float64) * np.cos(th)) @ rot(u, th) @ tr(tm.astype(np.float64))
    return T.astype(np.float32), M, n.astype(np.float32), c, a_, t_


def canonical_solve(p, n_p, q, n_q):
    """Independent fp64 restatement of the canonical iteration (what func.cpp:43-60,85-99 intend):
    rows [ (p~+q~) x n | n ] and residuals (p~-q~).n about the pair centroids, ONE joint least-squares solve by
    numpy.linalg.lstsq (not normal equations, not LDL^T), theta = atan|a|, and the composition in the order of the
    comment at func.cpp:94:  T = Tr(qbar) R(theta, a^) Tr(t cos theta) R(theta, a^) Tr(-pbar)."""
    p, q = p.astype(np.float64), q.astype(np.float64)
    n = n_p.astype(np.float64) + n_q.astype(np.float64)
    pbar, qbar = p.mean(axis=0), q.mean(axis=0)
    pc, qc = p - pbar, q - qbar
    A = np.c_[np.cross(pc + qc, n), n]
    b = -np.einsum("ij,ij->i", pc - qc, n)
    x = np.linalg.lstsq(A, b, rcond=None)[0]
    a, t = x[:3], x[3:]
    th = np.arctan(np.linalg.norm(a))
    u = a / np.linalg.norm(a)
    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)

    def tr(v):
        M = np.eye(4); M[:3, 3] = v
        return M

    R4 = np.eye(4); R4[:3, :3] = R
    return x, tr(qbar) @ R4 @ tr(t * np.cos(th)) @ R4 @ tr(-pbar)


def main():
    cat = read_pcd_xyz(f"{REF}/cat.pcd")
    cat_out = read_pcd_xyz(f"{REF}/cat_out.pcd")
    bunny = read_pcd_xyz(f"{REF}/txt2pcd_bunny1.pcd")
    za = np.loadtxt(f"{REF}/za.txt", dtype=np.float32)
    assert cat.shape == (3400, 3) and cat_out.shape == (3400, 3) and bunny.shape == (93, 3)
    assert np.abs(bunny - za).max() < 1e-6
    np.save(f"{OUT}/cat_xyz.npy", cat)
    np.save(f"{OUT}/cat_out_xyz.npy", cat_out)
    np.save(f"{OUT}/bunny_xyz.npy", bunny)

    g = {}
    # analytic known answer of the fixture
    th = np.pi / 4
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    g["truth_R"], g["truth_t"] = Rz, np.array([2.5, 0.0, 0.0])
    g["fixture_max_dev"] = np.abs(cat.astype(np.float64) @ Rz.T + g["truth_t"] - cat_out).max()

    g["eval_diff_seq_fp32"] = eval_diff_seq(cat, cat_out)

    # exact brute-force 1-NN cat -> cat_out at T = identity (pinned formula, lowest index on ties)
    D = d2_pinned(cat, cat_out)
    g["nn_idx"] = D.argmin(axis=1).astype(np.int32)   # argmin returns the first (lowest) index on ties
    g["nn_d2"] = D.min(axis=1).astype(np.float32)
    g["nn_ties"] = int(((D == D.min(axis=1, keepdims=True)).sum(axis=1) > 1).sum())
    # bunny self-NN excluding nothing (d2 = 0 at self) and bunny->perturbed made in the tests

    # OpenCV-bundled FLANN, KDTREE_SINGLE (algorithm 4), leaf_max_size 15
    import cv2
    fl = cv2.flann_Index(cat_out, dict(algorithm=4, leaf_max_size=15))
    fi, fd = fl.knnSearch(cat, 1, params=dict(checks=-1, eps=0.0, sorted=True))
    g["flann_idx"] = fi[:, 0].astype(np.int32)
    g["flann_d2"] = fd[:, 0].astype(np.float32)

    # regist.h:8-72 through OpenCV exactly as the reference calls it
    s64, d64 = cat.astype(np.float64), cat_out.astype(np.float64)
    sc, dc = s64.sum(axis=0) / len(s64), d64.sum(axis=0) / len(d64)
    S = (s64 - sc).T @ (d64 - dc)
    W, U, Vt = cv2.SVDecomp(S)
    det = np.linalg.det(U @ Vt)
    Rk = Vt.T @ np.diag([1, 1, det]) @ U.T
    g["kabsch_R"], g["kabsch_T"] = Rk, dc - Rk @ sc

    # normals
    for name, xyz in (("cat", cat), ("cat_out", cat_out)):
        nrm, gap, nb = pca_normals(xyz)
        g[f"normals_{name}"], g[f"normals_gap_{name}"], g[f"knn_{name}"] = nrm, gap, nb

    # literal first iteration pieces + literal loop (identity pairing, all quirks)
    sn, tn = g["normals_cat"][:, :3].copy(), g["normals_cat_out"][:, :3].copy()
    T1, M, N, c, a_, t_ = estimate_literal(cat, sn, cat_out, tn)
    g["lit_M"], g["lit_N"], g["lit_c"], g["lit_a"], g["lit_t"], g["lit_T1"] = M, N, c, a_, t_, T1
    src, srcn = cat.copy(), sn.copy()
    T = np.eye(4, dtype=np.float32)
    diffs = [eval_diff_seq(src, cat_out)]
    it = 0
    while diffs[-1] > 1.0 and it < 10:
        it += 1
        inc = estimate_literal(src, srcn, cat_out, tn)[0]
        src = (np.c_[src, np.ones(len(src), np.float32)] @ inc.T)[:, :3].astype(np.float32)
        srcn = (np.c_[srcn, np.ones(len(srcn), np.float32)] @ inc.T)[:, :3].astype(np.float32)
        T = (inc @ T).astype(np.float32)
        diffs.append(eval_diff_seq(src, cat_out))
    g["lit_diffs"], g["lit_T_final"] = np.array(diffs, np.float32), T

    # canonical joint solve, independent of the oracle and of the device code (same derivation nowhere else):
    # identity pairs (what myicp.cpp:128-131 does) and exact-NN pairs, PCA normals from above
    sol = {}
    n_src, n_tgt = g["normals_cat"][:, :3], g["normals_cat_out"][:, :3]
    sol["x_identity"], sol["T_identity"] = canonical_solve(cat, n_src, cat_out, n_tgt)
    nn = g["nn_idx"]
    sol["x_nn"], sol["T_nn"] = canonical_solve(cat, n_src, cat_out[nn], n_tgt[nn])
    np.savez_compressed(f"{OUT}/cat_solve_golden.npz", **sol)
    for k, v in sol.items():
        print(k, np.asarray(v).shape, np.asarray(v).round(6).tolist() if np.asarray(v).size <= 6 else "")

    np.savez_compressed(f"{OUT}/cat_golden.npz", **g)
    for k, v in g.items():
        v = np.asarray(v)
        print(k, v.shape, v.dtype, v if v.size <= 4 else "")


if __name__ == "__main__":
    main()
