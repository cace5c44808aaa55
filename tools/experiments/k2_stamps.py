import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import numpy as np
import icpsymm as icp
import workloads as wl
src, tgt, T = wl.config_synth(1_000_000)
with icp.Context(0) as ctx:
    ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
    ctx.set_option(icp.OPT_USE_GRAPH, 0)
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    for rep in range(3):
        Tm, st, rc = ctx.register_symm(1, 1.0)   # last K2 launch = a solving one? (3 searches: the third K2 stops at max_iters)
        buf = (C.c_ulonglong * 8)()
        ctx.L.icpsymm_debug_k2_stamps(buf)
        a = np.array(list(buf), dtype=np.float64)
        print("last K2 (stops): reduce %.0f, +cluster combine %.0f ns | first K2 (solves): until solved %.0f, until arrived %.0f ns" % (a[1]-a[0], a[2]-a[0], a[3]-a[5], a[4]-a[5]))
