import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import numpy as np
import icpsymm as icp
import workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
src, tgt, T = wl.config_synth(n)
for on in (1, 0):
    with icp.Context(0) as ctx:
        L = ctx.L
        ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
        ctx.set_option(icp.OPT_USE_GRAPH, 0)
        ctx.set_option(icp.OPT_KERNEL_TIMING, 1)
        ctx.set_option(icp.OPT_ORIENTED_BOUNDS, on)
        ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
        buf = (C.c_ulonglong * 32)()
        for iters in (1, 2, 3, 4, 7):
            L.icpsymm_debug_visit_counters(buf, 1)
            Tm, st, rc = ctx.register_symm(iters - 1, 1.0) if iters > 1 else ctx.register_symm(0, 1.0)
            L.icpsymm_debug_visit_counters(buf, 1)
            c = np.array(list(buf), dtype=np.float64)
            w = max(c[22], 1)
            print(f"oriented={on} searches={st['searches']} search_ms={st['search_ms']:.3f} warps={int(c[22])} per-warp: far nodes L1..6 {np.round(c[1:7]/w,1)} near nodes L1..6 {np.round(c[9:15]/w,1)} leaves far {c[16]/w:.1f} near {c[17]/w:.1f} rechecks {c[20]/w:.1f} pass {c[21]/w:.1f}")
