import os, sys, glob, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
    import icpsymm as icp
    icp.LIB_PATH = sys.argv[1]
    import workloads as wl
    src, tgt, T = wl.config_synth(1_000_000)
    with icp.Context(0) as ctx:
        ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
        ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
        ms = []
        for _ in range(8):
            Tm, st, rc = ctx.register_symm(50, 1.0)
            ms.append(st["loop_ms"])
        ctx.set_option(icp.OPT_USE_GRAPH, 0); ctx.set_option(icp.OPT_KERNEL_TIMING, 1)
        Tm, st, rc = ctx.register_symm(50, 1.0)
        Tm, st, rc = ctx.register_symm(50, 1.0)
        print(os.path.basename(sys.argv[1]), "loop_ms best %.3f median %.3f; solve kernel avg %.2f us" % (min(ms), sorted(ms)[4], 1e3 * st["solve_ms"] / st["solve_launches"]), flush=True)
else:
    libs = sorted(glob.glob(os.path.join(ROOT, "tools/_scratch/variants/*.so"))) + [os.path.join(ROOT, "icp-symm_b200/libicpsymm_b200.so")]
    for l in libs:
        subprocess.run([sys.executable, __file__, l])
