import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import icpsymm as icp
import workloads as wl
src, tgt, T = wl.config_synth(1_000_000)
with icp.Context(0) as ctx:
    ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    for gb in [0]:
        ctx.set_option(icp.OPT_GRID_BLOCKS, gb)
        ms = []
        for _ in range(6):
            Tm, st, rc = ctx.register_symm(50, 1.0)
            ms.append(st["loop_ms"])
        print("grid_blocks", gb, "loop_ms best %.3f median %.3f" % (min(ms), sorted(ms)[3]), flush=True)
