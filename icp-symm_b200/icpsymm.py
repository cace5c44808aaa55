"""ctypes binding of libicpsymm_b200.so (the C ABI in include/icpsymm.h).

Used by tests/, bench.py and __graft_entry__.py.  It adds nothing to the product: every call is
one C-ABI call on host numpy buffers.  There is no CPU fallback -- if the shared library is
missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libicpsymm_b200.so")   # developer tools point this at an alternative build before load()
NSUMS = 37

OK, WARN_DEGENERATE = 0, 1
ERR_INVALID, ERR_CUDA, ERR_NOT_READY, ERR_SIZE, ERR_NONFINITE, ERR_COMM = -1, -2, -3, -4, -5, -6
OPT_REL_DIFF_TOL, OPT_IDENTITY_PAIRS, OPT_MAX_PAIR_DIST, OPT_USE_GRAPH = 1, 2, 3, 4
OPT_KERNEL_TIMING, OPT_WARM_START, OPT_TRACE, OPT_GRID_BLOCKS, OPT_ASYNC_SETUP = 5, 6, 7, 8, 9
OPT_MIN_NORMAL_DOT, OPT_PEER_TIMEOUT_MS, OPT_TILE_ORDER, OPT_SKIP_WALK, OPT_PROBE = 10, 11, 12, 13, 14


class Stats(C.Structure):
    _fields_ = [("iters", C.c_int32), ("converged", C.c_int32), ("degenerate", C.c_int32), ("searches", C.c_int32),
                ("diff", C.c_double), ("rmse", C.c_double), ("min_pivot_ratio", C.c_double), ("pairs", C.c_uint64),
                ("loop_ms", C.c_float), ("search_ms", C.c_float), ("solve_ms", C.c_float),
                ("search_launches", C.c_int32), ("solve_launches", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class IcpError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"icpsymm status {status}: {msg}")
        self.status = status


_lib = None


def load():
    """dlopen the product library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp, fp, ip, dp, u32p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    sz = C.c_size_t
    sig = {
        "icpsymm_create": (C.c_int, [C.POINTER(vp), C.c_int]),
        "icpsymm_create_on_stream": (C.c_int, [C.POINTER(vp), C.c_int, vp]),
        "icpsymm_destroy": (None, [vp]),
        "icpsymm_status_string": (C.c_char_p, [C.c_int]),
        "icpsymm_last_error": (C.c_char_p, [vp]),
        "icpsymm_get_stream": (vp, [vp]),
        "icpsymm_synchronize": (C.c_int, [vp]),
        "icpsymm_set_option": (C.c_int, [vp, C.c_int, C.c_double]),
        "icpsymm_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(sz)]),
        "icpsymm_set_source": (C.c_int, [vp, vp, sz, sz]),
        "icpsymm_set_target": (C.c_int, [vp, vp, sz, sz]),
        "icpsymm_set_source_dev": (C.c_int, [vp, vp, sz, sz]),
        "icpsymm_set_target_dev": (C.c_int, [vp, vp, sz, sz]),
        "icpsymm_set_source_batch": (C.c_int, [vp, vp, u32p, sz, sz, sz]),
        "icpsymm_set_target_batch": (C.c_int, [vp, vp, u32p, sz, sz, sz]),
        "icpsymm_estimate_normals": (C.c_int, [vp, C.c_int, fp]),
        "icpsymm_set_normals": (C.c_int, [vp, C.c_int, fp]),
        "icpsymm_get_normals": (C.c_int, [vp, C.c_int, fp]),
        "icpsymm_register_symm": (C.c_int, [vp, C.c_int, C.c_float, fp, C.POINTER(Stats)]),
        "icpsymm_register_symm_async": (C.c_int, [vp, C.c_int, C.c_float]),
        "icpsymm_wait": (C.c_int, [vp, fp, C.POINTER(Stats)]),
        "icpsymm_find_correspondences": (C.c_int, [vp, fp, ip, fp]),
        "icpsymm_get_correspondences": (C.c_int, [vp, ip, fp]),
        "icpsymm_eval_diff": (C.c_int, [vp, fp, dp]),
        "icpsymm_normal_equations": (C.c_int, [vp, fp, dp, dp]),
        "icpsymm_solve_compose": (C.c_int, [vp, dp, dp, dp, dp, dp]),
        "icpsymm_get_transformed_source": (C.c_int, [vp, fp, fp, sz]),
        "icpsymm_get_trace": (C.c_int, [vp, sz, C.c_int, fp, dp]),
        "icpsymm_register_npoint": (C.c_int, [vp, dp, dp, sz, dp, dp]),
        "icpsymm_get_search_times": (C.c_int, [vp, fp, fp, C.c_int]),
        "icpsymm_register_p2p": (C.c_int, [vp, C.c_int, C.c_float, fp, C.POINTER(Stats)]),
        "icpsymm_comm_unique_id": (C.c_int, [vp]),
        "icpsymm_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "icpsymm_comm_destroy": (C.c_int, [vp]),
        "icpsymm_set_query_shard": (C.c_int, [vp, C.c_int, C.c_int]),
        "icpsymm_peer_export": (C.c_int, [vp, vp]),
        "icpsymm_peer_attach": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "icpsymm_peer_detach": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def tile_rows(n_points):
    """fp64 rows K1 leaves per pair and search: one per 256-query tile (csrc/common.cuh)"""
    return (n_points + 255) // 256


def exported_symbols():
    return [n for n in dir(load()) if n.startswith("icpsymm_")]


def device_info(device=0):
    sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
    rc = load().icpsymm_device_info(device, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem))
    if rc != 0:
        raise IcpError(rc, "no CUDA device (the product has no CPU fallback)")
    return dict(sm_count=sm.value, cc=(ma.value, mi.value), total_mem=mem.value)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _rows(a):
    """(n, >=3) float32 rows; keeps a strided view (e.g. 16-byte pcl::PointXYZ rows) when it is addressable"""
    a = np.asarray(a)
    if a.ndim != 2 or a.shape[1] < 3:
        raise ValueError("expected an (n, 3) array")
    if a.dtype == np.float32 and a.shape[0] > 0 and a.strides[1] == 4 and a.strides[0] >= 12 and a.strides[0] % 4 == 0:
        return a
    return np.ascontiguousarray(a[:, :3], dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _T16(T):
    T = np.eye(4, dtype=np.float32) if T is None else _f32(np.asarray(T).reshape(4, 4))
    return np.ascontiguousarray(T.reshape(16))


class Context:
    """One registration context = one MyICP object (ICP/myicp.h:7-36) on one GPU."""

    def __init__(self, device=0, stream=None):
        self.L = load()
        self.h = C.c_void_p()
        if stream is None:
            rc = self.L.icpsymm_create(C.byref(self.h), device)
        else:
            rc = self.L.icpsymm_create_on_stream(C.byref(self.h), device, C.c_void_p(stream))
        if rc != 0:
            self.h = C.c_void_p()
            raise IcpError(rc, self.L.icpsymm_status_string(rc).decode())
        self.n_src = self.n_tgt = 0
        self.n_pairs = 1

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.icpsymm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, allow_warn=True):
        if rc < 0 or (rc > 0 and not allow_warn):
            raise IcpError(rc, self.L.icpsymm_last_error(self.h).decode() or self.L.icpsymm_status_string(rc).decode())
        return rc

    @property
    def stream(self):
        return self.L.icpsymm_get_stream(self.h)

    def synchronize(self):
        self._check(self.L.icpsymm_synchronize(self.h))

    def set_option(self, key, value):
        self._check(self.L.icpsymm_set_option(self.h, key, float(value)))

    # ---- clouds ----
    def set_source(self, xyz):
        xyz = _rows(xyz)
        self._src_keep = xyz
        self._check(self.L.icpsymm_set_source(self.h, xyz.ctypes.data, xyz.shape[0], xyz.strides[0]))
        self.n_src, self.n_pairs = xyz.shape[0], 1

    def set_target(self, xyz):
        xyz = _rows(xyz)
        self._tgt_keep = xyz
        self._check(self.L.icpsymm_set_target(self.h, xyz.ctypes.data, xyz.shape[0], xyz.strides[0]))
        self.n_tgt, self.n_pairs = xyz.shape[0], 1

    def set_source_ptr(self, host_ptr, n, stride):
        """host pointer variant (pinned buffers of the bench)"""
        self._check(self.L.icpsymm_set_source(self.h, C.c_void_p(host_ptr), n, stride))
        self.n_src, self.n_pairs = n, 1

    def set_target_ptr(self, host_ptr, n, stride):
        self._check(self.L.icpsymm_set_target(self.h, C.c_void_p(host_ptr), n, stride))
        self.n_tgt, self.n_pairs = n, 1

    def set_source_dev(self, dev_ptr, n, stride):
        self._check(self.L.icpsymm_set_source_dev(self.h, C.c_void_p(dev_ptr), n, stride))
        self.n_src, self.n_pairs = n, 1

    def set_target_dev(self, dev_ptr, n, stride):
        self._check(self.L.icpsymm_set_target_dev(self.h, C.c_void_p(dev_ptr), n, stride))
        self.n_tgt, self.n_pairs = n, 1

    def set_source_batch_ptr(self, host_ptr, n_pairs, n, stride):
        """host pointer variant (pinned buffers of the bench): n_pairs clouds of n points each"""
        self._check(self.L.icpsymm_set_source_batch(self.h, C.c_void_p(host_ptr), None, n_pairs, n, stride))
        self.n_pairs, self.n_src = n_pairs, n

    def set_target_batch_ptr(self, host_ptr, n_pairs, n, stride):
        self._check(self.L.icpsymm_set_target_batch(self.h, C.c_void_p(host_ptr), None, n_pairs, n, stride))
        self.n_pairs, self.n_tgt = n_pairs, n

    def _set_batch(self, fn, xyz, counts):
        xyz = _f32(xyz)
        assert xyz.ndim == 3 and xyz.shape[2] >= 3
        b, m = xyz.shape[0], xyz.shape[1]
        cp = None
        if counts is not None:
            counts = np.ascontiguousarray(counts, dtype=np.uint32)
            cp = counts.ctypes.data_as(C.POINTER(C.c_uint32))
        self._check(fn(self.h, xyz.ctypes.data, cp, b, m, xyz.strides[1]))
        return b, m

    def set_source_batch(self, xyz, counts=None):
        self.n_pairs, self.n_src = self._set_batch(self.L.icpsymm_set_source_batch, xyz, counts)

    def set_target_batch(self, xyz, counts=None):
        self.n_pairs, self.n_tgt = self._set_batch(self.L.icpsymm_set_target_batch, xyz, counts)

    # ---- normals ----
    def estimate_normals(self, k=10, viewpoint=None):
        vp = None if viewpoint is None else _fp(_f32(np.asarray(viewpoint)))
        self._check(self.L.icpsymm_estimate_normals(self.h, k, vp))

    def set_normals(self, which, normals4):
        n4 = _f32(normals4)
        self._check(self.L.icpsymm_set_normals(self.h, which, _fp(n4)))

    def get_normals(self, which):
        n = (self.n_tgt if which else self.n_src) * self.n_pairs
        out = np.empty((n, 4), np.float32)
        self._check(self.L.icpsymm_get_normals(self.h, which, _fp(out)))
        return out if self.n_pairs == 1 else out.reshape(self.n_pairs, -1, 4)

    # ---- registration ----
    def register_symm(self, max_iters=10, diff_threshold=1.0):
        """MyICP::RegisterSymm (myicp.cpp:100-150) with the reference defaults 10 / 1.0f."""
        T = np.empty((self.n_pairs, 16), np.float32)
        st = (Stats * self.n_pairs)()
        rc = self._check(self.L.icpsymm_register_symm(self.h, max_iters, diff_threshold, _fp(T), st))
        stats = [s.as_dict() for s in st]
        if self.n_pairs == 1:
            return T.reshape(4, 4), stats[0], rc
        return T.reshape(-1, 4, 4), stats, rc

    def register_p2p(self, max_iters=10, diff_threshold=1.0):
        """point-to-point ICP (MyICP::RegisterP2P): correspondences + closed-form fit, all on the device"""
        T = np.empty(16, np.float32)
        st = Stats()
        rc = self._check(self.L.icpsymm_register_p2p(self.h, max_iters, diff_threshold, _fp(T), C.byref(st)))
        return T.reshape(4, 4), st.as_dict(), rc

    def register_symm_async(self, max_iters=10, diff_threshold=1.0):
        """enqueue the loop and return; wait() hands out the result"""
        self._check(self.L.icpsymm_register_symm_async(self.h, max_iters, diff_threshold))

    def wait(self):
        T = np.empty((self.n_pairs, 16), np.float32)
        st = (Stats * self.n_pairs)()
        rc = self._check(self.L.icpsymm_wait(self.h, _fp(T), st))
        stats = [s.as_dict() for s in st]
        if self.n_pairs == 1:
            return T.reshape(4, 4), stats[0], rc
        return T.reshape(-1, 4, 4), stats, rc

    def find_correspondences(self, T=None):
        T = _T16(T)
        n = self.n_src * self.n_pairs
        idx = np.empty(n, np.int32)
        d2 = np.empty(n, np.float32)
        self._check(self.L.icpsymm_find_correspondences(self.h, _fp(T), idx.ctypes.data_as(C.POINTER(C.c_int32)), _fp(d2)))
        return idx, d2

    def get_correspondences(self):
        n = self.n_src * self.n_pairs
        idx = np.empty(n, np.int32)
        d2 = np.empty(n, np.float32)
        self._check(self.L.icpsymm_get_correspondences(self.h, idx.ctypes.data_as(C.POINTER(C.c_int32)), _fp(d2)))
        return idx, d2

    def eval_diff(self, T=None):
        T = _T16(T)
        d = C.c_double()
        self._check(self.L.icpsymm_eval_diff(self.h, _fp(T), C.byref(d)))
        return d.value

    def normal_equations(self, T=None, origin=(0.0, 0.0, 0.0)):
        T = _T16(T)
        o = np.ascontiguousarray(origin, np.float64)
        S = np.empty(NSUMS, np.float64)
        dp = C.POINTER(C.c_double)
        self._check(self.L.icpsymm_normal_equations(self.h, _fp(T), o.ctypes.data_as(dp), S.ctypes.data_as(dp)))
        return S

    def solve_compose(self, sums, origin=(0.0, 0.0, 0.0)):
        S = np.ascontiguousarray(sums, np.float64)
        o = np.ascontiguousarray(origin, np.float64)
        T = np.empty(16, np.float64)
        x = np.empty(6, np.float64)
        r = C.c_double()
        dp = C.POINTER(C.c_double)
        rc = self._check(self.L.icpsymm_solve_compose(self.h, S.ctypes.data_as(dp), o.ctypes.data_as(dp),
                                                      T.ctypes.data_as(dp), x.ctypes.data_as(dp), C.byref(r)))
        return T.reshape(4, 4), x, rc, r.value

    def transformed_source(self, T=None):
        T = _T16(T)
        out = np.empty((self.n_src * self.n_pairs, 3), np.float32)
        self._check(self.L.icpsymm_get_transformed_source(self.h, _fp(T), _fp(out), 12))
        return out

    def trace(self, k, pair=0):
        T = np.empty(16, np.float32)
        d = C.c_double()
        self._check(self.L.icpsymm_get_trace(self.h, pair, k, _fp(T), C.byref(d)))
        return T.reshape(4, 4), d.value

    def search_times(self, n=64):
        """per-launch ms of the last registration timed with OPT_KERNEL_TIMING: (search[k], solve[k])"""
        a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
        k = self._check(self.L.icpsymm_get_search_times(self.h, _fp(a), _fp(b), n))
        return a[:k].copy(), b[:k].copy()

    def register_npoint(self, src, dst):
        """registrateNPoint (regist.h:8-72): dst = R src + T"""
        src = np.ascontiguousarray(src, np.float64)
        dst = np.ascontiguousarray(dst, np.float64)
        R = np.empty(9, np.float64)
        T = np.empty(3, np.float64)
        dp = C.POINTER(C.c_double)
        self._check(self.L.icpsymm_register_npoint(self.h, src.ctypes.data_as(dp), dst.ctypes.data_as(dp),
                                                   min(len(src), len(dst)), R.ctypes.data_as(dp), T.ctypes.data_as(dp)))
        return R.reshape(3, 3), T

    # ---- multi-GPU ----
    def set_query_shard(self, rank, world):
        self._check(self.L.icpsymm_set_query_shard(self.h, rank, world))

    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.L.icpsymm_peer_export(self.h, buf))
        return buf.raw

    def peer_attach(self, handles: bytes, world, rank):
        buf = C.create_string_buffer(handles, 64 * world)
        self._check(self.L.icpsymm_peer_attach(self.h, buf, world, rank))

    def peer_detach(self):
        self._check(self.L.icpsymm_peer_detach(self.h))

    def comm_destroy(self):
        self._check(self.L.icpsymm_comm_destroy(self.h))

    def comm_init(self, unique_id: bytes, world, rank):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.L.icpsymm_comm_init(self.h, buf, world, rank))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = load().icpsymm_comm_unique_id(buf)
    if rc != 0:
        raise IcpError(rc, "ncclGetUniqueId failed")
    return buf.raw
