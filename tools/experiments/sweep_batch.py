import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import numpy as np
import icpsymm as icp
import workloads as wl
n, pairs, distinct = 32768, 256, 32
S = np.empty((pairs, n, 3), np.float32); T = np.empty((pairs, n, 3), np.float32)
for j in range(distinct):
    s, t, Tt = wl.config_batch_pair(j, n); S[j], T[j] = s, t
for j in range(distinct, pairs): S[j], T[j] = S[j % distinct], T[j % distinct]
with icp.Context(0) as ctx:
    ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
    L = ctx.L
    ctx._check(L.icpsymm_set_source_batch(ctx.h, S.ctypes.data, None, pairs, n, 12))
    ctx._check(L.icpsymm_set_target_batch(ctx.h, T.ctypes.data, None, pairs, n, 12))
    ctx.n_pairs, ctx.n_src, ctx.n_tgt = pairs, n, n
    ctx.estimate_normals(10)
    for gb in [0, 4, 10, 16, 24, 32, 64, 128]:
        ctx.set_option(icp.OPT_GRID_BLOCKS, gb)
        ms = []
        for _ in range(4):
            Tm, st, _ = ctx.register_symm(50, 1e-3)
            ms.append(st[0]["loop_ms"])
        print("blocks per pair", gb, "loop_ms best %.3f median %.3f -> %.0f pairs/s" % (min(ms), sorted(ms)[2], pairs / (min(ms) * 1e-3)), flush=True)
