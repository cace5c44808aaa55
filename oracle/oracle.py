"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NSUMS = 37


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("icp_oracle.c", "icp_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return so


class Params(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("diff_threshold", C.c_float), ("rel_diff_tol", C.c_double),
                ("identity_pairs", C.c_int), ("threads", C.c_int), ("use_brute", C.c_int),
                ("max_pair_dist", C.c_float), ("min_normal_dot", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("iters", C.c_int), ("converged", C.c_int), ("degenerate", C.c_int),
                ("diff", C.c_double), ("rmse", C.c_double), ("pairs", C.c_size_t),
                ("min_pivot_ratio", C.c_double)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double)
        L.orc_pcd_read_xyz.argtypes = [C.c_char_p, C.POINTER(fp), C.POINTER(C.c_size_t)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_transform_points.argtypes = [fp, fp, C.c_size_t, fp]
        L.orc_rotate_vectors.argtypes = [fp, fp, C.c_size_t, C.c_size_t, fp]
        L.orc_nn_brute.argtypes = [fp, C.c_size_t, fp, C.c_size_t, ip, fp]
        L.orc_kdtree_build.restype = C.c_void_p
        L.orc_kdtree_build.argtypes = [fp, C.c_size_t]
        L.orc_kdtree_free.argtypes = [C.c_void_p]
        L.orc_kdtree_nn.argtypes = [C.c_void_p, fp, C.c_size_t, ip, fp, C.c_int]
        L.orc_kdtree_knn.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_int, ip, fp, C.c_int]
        L.orc_estimate_normals.argtypes = [fp, C.c_size_t, C.c_int, fp, fp, fp, C.c_int]
        L.orc_eval_diff.restype = C.c_float
        L.orc_eval_diff.argtypes = [fp, fp, C.c_size_t]
        L.orc_matrix_notation.argtypes = [fp, fp, fp, fp, C.c_size_t, fp, fp, fp]
        L.orc_solve_lls.argtypes = [fp, fp, C.c_size_t, fp]
        L.orc_estimate_transform_symm_literal.argtypes = [fp, fp, fp, fp, C.c_size_t, fp]
        L.orc_apply_transform_literal.argtypes = [fp, C.c_size_t, fp, fp]
        L.orc_register_symm_literal.restype = C.c_int
        L.orc_register_symm_literal.argtypes = [fp, fp, fp, fp, C.c_size_t, C.c_int, C.c_float, fp, fp]
        L.orc_registrate_npoint.restype = C.c_int
        L.orc_registrate_npoint.argtypes = [dp, dp, C.c_size_t, dp, dp]
        L.orc_symm_accumulate.argtypes = [fp, fp, fp, fp, C.c_size_t, dp, dp]
        L.orc_solve_compose.restype = C.c_int
        L.orc_solve_compose.argtypes = [dp, dp, dp, dp, dp]
        L.orc_register_symm.restype = C.c_int
        L.orc_register_symm.argtypes = [fp, fp, C.c_size_t, fp, fp, C.c_size_t, C.POINTER(Params), fp,
                                        C.POINTER(Stats), fp, dp, ip]
        L.orc_num_threads.restype = C.c_int
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def pcd_read_xyz(path: str) -> np.ndarray:
    p = C.POINTER(C.c_float)()
    n = C.c_size_t()
    rc = lib().orc_pcd_read_xyz(path.encode(), C.byref(p), C.byref(n))
    if rc != 0:
        raise IOError(f"orc_pcd_read_xyz({path}) -> {rc}")
    out = np.ctypeslib.as_array(p, shape=(n.value, 3)).copy()
    lib().orc_free(p)
    return out


def transform_points(T, xyz):
    T, Tp = _f(np.asarray(T).reshape(16))
    xyz, xp = _f(xyz)
    out = np.empty_like(xyz)
    lib().orc_transform_points(Tp, xp, len(xyz), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def rotate_vectors(T, v):
    T, Tp = _f(np.asarray(T).reshape(16))
    v, vp = _f(v)
    out = np.empty_like(v)
    lib().orc_rotate_vectors(Tp, vp, len(v), v.shape[1], out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def nn_brute(q, t):
    q, qp = _f(q)
    t, tp = _f(t)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.float32)
    lib().orc_nn_brute(qp, len(q), tp, len(t), idx.ctypes.data_as(C.POINTER(C.c_int32)),
                       d2.ctypes.data_as(C.POINTER(C.c_float)))
    return idx, d2


class KdTree:
    def __init__(self, xyz):
        self.xyz, p = _f(xyz)
        self.h = lib().orc_kdtree_build(p, len(self.xyz))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_kdtree_free(self.h)
            self.h = None

    def nn(self, q, threads=0):
        q, qp = _f(q)
        idx = np.empty(len(q), np.int32)
        d2 = np.empty(len(q), np.float32)
        lib().orc_kdtree_nn(self.h, qp, len(q), idx.ctypes.data_as(C.POINTER(C.c_int32)),
                            d2.ctypes.data_as(C.POINTER(C.c_float)), threads)
        return idx, d2

    def knn(self, q, k, threads=0):
        q, qp = _f(q)
        idx = np.empty((len(q), k), np.int32)
        d2 = np.empty((len(q), k), np.float32)
        lib().orc_kdtree_knn(self.h, qp, len(q), k, idx.ctypes.data_as(C.POINTER(C.c_int32)),
                             d2.ctypes.data_as(C.POINTER(C.c_float)), threads)
        return idx, d2


def estimate_normals(xyz, k=10, viewpoint=(0.0, 0.0, 0.0), threads=0):
    xyz, xp = _f(xyz)
    vp, vpp = _f(np.asarray(viewpoint, np.float32))
    nrm = np.empty((len(xyz), 4), np.float32)
    gap = np.empty(len(xyz), np.float32)
    lib().orc_estimate_normals(xp, len(xyz), k, vpp, nrm.ctypes.data_as(C.POINTER(C.c_float)),
                               gap.ctypes.data_as(C.POINTER(C.c_float)), threads)
    return nrm, gap


def eval_diff(src, tgt) -> float:
    src, sp = _f(src)
    tgt, tp = _f(tgt)
    assert src.shape == tgt.shape
    return float(lib().orc_eval_diff(sp, tp, len(src)))


def matrix_notation(src, src_n, tgt, tgt_n):
    src, a = _f(src); src_n, b = _f(src_n); tgt, c = _f(tgt); tgt_n, d = _f(tgt_n)
    n = len(src)
    M = np.empty((n, 3), np.float32); N = np.empty((n, 3), np.float32); cc = np.empty(n, np.float32)
    fp = C.POINTER(C.c_float)
    lib().orc_matrix_notation(a, b, c, d, n, M.ctypes.data_as(fp), N.ctypes.data_as(fp), cc.ctypes.data_as(fp))
    return M, N, cc


def solve_lls(A, b):
    A, ap = _f(A); b, bp = _f(b)
    x = np.empty(3, np.float32)
    lib().orc_solve_lls(ap, bp, len(A), x.ctypes.data_as(C.POINTER(C.c_float)))
    return x


def estimate_transform_symm_literal(src, src_n, tgt, tgt_n):
    src, a = _f(src); src_n, b = _f(src_n); tgt, c = _f(tgt); tgt_n, d = _f(tgt_n)
    T = np.empty(16, np.float32)
    lib().orc_estimate_transform_symm_literal(a, b, c, d, len(src), T.ctypes.data_as(C.POINTER(C.c_float)))
    return T.reshape(4, 4)


def apply_transform_literal(src, T):
    src, a = _f(src); T, tp = _f(np.asarray(T).reshape(16))
    out = np.empty_like(src)
    lib().orc_apply_transform_literal(a, len(src), tp, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def register_symm_literal(src, src_n3, tgt, tgt_n3, max_iters=10, thr=1.0):
    src, a = _f(src); src_n3, b = _f(src_n3); tgt, c = _f(tgt); tgt_n3, d = _f(tgt_n3)
    T = np.empty(16, np.float32)
    diffs = np.zeros(max_iters + 2, np.float32)
    fp = C.POINTER(C.c_float)
    it = lib().orc_register_symm_literal(a, b, c, d, len(src), max_iters, thr, T.ctypes.data_as(fp),
                                         diffs.ctypes.data_as(fp))
    return T.reshape(4, 4), diffs[: it + 1], it


def registrate_npoint(src, dst):
    src, a = _d(src); dst, b = _d(dst)
    R = np.empty(9, np.float64); T = np.empty(3, np.float64)
    dp = C.POINTER(C.c_double)
    rc = lib().orc_registrate_npoint(a, b, len(src), R.ctypes.data_as(dp), T.ctypes.data_as(dp))
    return rc, R.reshape(3, 3), T


def symm_accumulate(p, np4, q, nq4, origin=(0, 0, 0)):
    p, a = _f(p); np4, b = _f(np4); q, c = _f(q); nq4, d = _f(nq4)
    o, op = _d(np.asarray(origin, np.float64))
    S = np.empty(NSUMS, np.float64)
    lib().orc_symm_accumulate(a, b, c, d, len(p), op, S.ctypes.data_as(C.POINTER(C.c_double)))
    return S


def solve_compose(S, origin=(0, 0, 0)):
    S, sp = _d(S)
    o, op = _d(np.asarray(origin, np.float64))
    T = np.empty(16, np.float64); x = np.empty(6, np.float64)
    dp = C.POINTER(C.c_double)
    ratio = C.c_double()
    deg = lib().orc_solve_compose(sp, op, T.ctypes.data_as(dp), x.ctypes.data_as(dp), C.byref(ratio))
    solve_compose.last_min_pivot_ratio = ratio.value
    return T.reshape(4, 4), x, int(deg)


def register_symm(src, src_n4, tgt, tgt_n4, max_iters=10, diff_threshold=1.0, rel_diff_tol=0.0,
                  identity_pairs=False, threads=0, use_brute=False, max_pair_dist=0.0, min_normal_dot=-2.0):
    """Oracle-C loop. Returns dict(T, stats, trace_T, trace_diff, idx)."""
    src, a = _f(src); src_n4, b = _f(src_n4); tgt, c = _f(tgt); tgt_n4, d = _f(tgt_n4)
    prm = Params(max_iters, diff_threshold, rel_diff_tol, int(identity_pairs), threads, int(use_brute),
                 max_pair_dist, min_normal_dot)
    st = Stats()
    T = np.empty(16, np.float32)
    trT = np.zeros((max_iters + 1, 16), np.float32)
    trd = np.zeros(max_iters + 1, np.float64)
    idx = np.empty(len(src), np.int32)
    fp = C.POINTER(C.c_float)
    rc = lib().orc_register_symm(a, b, len(src), c, d, len(tgt), C.byref(prm), T.ctypes.data_as(fp),
                                 C.byref(st), trT.ctypes.data_as(fp), trd.ctypes.data_as(C.POINTER(C.c_double)),
                                 idx.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise RuntimeError(f"orc_register_symm -> {rc}")
    ns = st.iters + 1
    return dict(T=T.reshape(4, 4), iters=st.iters, converged=st.converged, degenerate=st.degenerate,
                diff=st.diff, rmse=st.rmse, pairs=st.pairs, min_pivot_ratio=st.min_pivot_ratio, trace_T=trT[:ns].reshape(ns, 4, 4),
                trace_diff=trd[:ns], idx=idx)
