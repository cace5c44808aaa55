import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import icpsymm as icp
import workloads as wl
for n in (4_000_000, 250_000):
    src, tgt, T = wl.config_synth(n)
    with icp.Context(0) as ctx:
        ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
        ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
        nt = (n + 255) // 256
        for gb in [0]:
            if gb > nt: continue
            ctx.set_option(icp.OPT_GRID_BLOCKS, gb)
            ms = []
            for _ in range(5):
                Tm, st, rc = ctx.register_symm(50, 1.0)
                ms.append(st["loop_ms"])
            print("n", n, "tiles", nt, "grid_blocks", gb, "loop_ms best %.3f median %.3f iters %d" % (min(ms), sorted(ms)[2], st["iters"]), flush=True)
