import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import numpy as np
import icpsymm as icp
import workloads as wl
n = 1_000_000
src, tgt, T = wl.config_synth(n)
with icp.Context(0) as ctx:
    ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    L = ctx.L
    L.icpsymm_debug_sorted_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.icpsymm_debug_set_tile_order.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    pb = np.empty((n, 4), np.float32)
    L.icpsymm_debug_sorted_points(ctx.h, 0, pb.ctypes.data_as(C.c_void_p), n)
    ids = pb[:, 3].copy().view(np.uint32)
    idx, d2 = ctx.find_correspondences()
    d_sorted = np.sqrt(d2[ids])
    ntiles = (n + 255) // 256
    pad = np.zeros(ntiles * 256); pad[:n] = d_sorted
    tile_mean = pad.reshape(ntiles, 256).mean(1)
    def run(tag):
        ms = []
        for _ in range(6):
            Tm, st, rc = ctx.register_symm(50, 1.0)
            ms.append(st["loop_ms"])
        ctx.set_option(icp.OPT_USE_GRAPH, 0); ctx.set_option(icp.OPT_KERNEL_TIMING, 1)
        s1 = min(ctx.register_symm(0, 1.0)[1]["search_ms"] for _ in range(3))
        s2 = min(ctx.register_symm(1, 1.0)[1]["search_ms"] for _ in range(3))
        ctx.set_option(icp.OPT_KERNEL_TIMING, 0); ctx.set_option(icp.OPT_USE_GRAPH, 1)
        print(f"{tag}: loop_ms best {min(ms):.3f} median {sorted(ms)[3]:.3f}; first search {s1:.3f} ms, second {s2 - s1:.3f} ms, iters {st['iters']}", flush=True)
    run("natural order")
    order = np.argsort(-tile_mean, kind="stable").astype(np.uint32)
    L.icpsymm_debug_set_tile_order(ctx.h, order.ctypes.data_as(C.c_void_p), ntiles)
    run("exact first-search distance, descending")
    rng = np.random.default_rng(0)
    order = rng.permutation(ntiles).astype(np.uint32)
    L.icpsymm_debug_set_tile_order(ctx.h, order.ctypes.data_as(C.c_void_p), ntiles)
    run("random permutation")
    # coarse classes (32 linear classes of the mean distance), stable inside a class
    cls = np.minimum((tile_mean / tile_mean.max() * 32).astype(int), 31)
    order = np.argsort(-cls, kind="stable").astype(np.uint32)
    L.icpsymm_debug_set_tile_order(ctx.h, order.ctypes.data_as(C.c_void_p), ntiles)
    run("32 distance classes, descending")
