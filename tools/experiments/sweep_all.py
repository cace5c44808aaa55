import os, sys, glob, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
    import numpy as np
    import icpsymm as icp
    icp.LIB_PATH = sys.argv[1]
    import workloads as wl
    src, tgt, T = wl.config_synth(1_000_000)
    with icp.Context(0) as ctx:
        ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
        tb, tn = [], []
        for _ in range(8):
            ctx.synchronize(); t0 = time.perf_counter()
            ctx.set_source(src); ctx.set_target(tgt); ctx.synchronize(); t1 = time.perf_counter()
            ctx.estimate_normals(10); ctx.synchronize(); t2 = time.perf_counter()
            tb.append(t1 - t0); tn.append(t2 - t1)
        ms = []
        for _ in range(6):
            Tm, st, rc = ctx.register_symm(50, 1.0)
            ms.append(st["loop_ms"])
        print(os.path.basename(sys.argv[1]), "upload+index both %.3f ms, normals both %.3f ms, loop_ms best %.3f median %.3f iters %d" % (min(tb)*1e3, min(tn)*1e3, min(ms), sorted(ms)[3], st["iters"]), flush=True)
else:
    libs = sorted(glob.glob(os.path.join(ROOT, "tools/_scratch/variants/*.so"))) + [os.path.join(ROOT, "icp-symm_b200/libicpsymm_b200.so")]
    for l in libs:
        subprocess.run([sys.executable, __file__, l])
