"""Small driver for ncu: one 1M-point registration (config 3) after setup; prints the loop stats."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import numpy as np
import icpsymm as icp
import workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
src, tgt, T = wl.config_synth(n)
with icp.Context(0) as ctx:
    ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
    ctx.set_option(icp.OPT_USE_GRAPH, 0)
    ctx.set_option(icp.OPT_KERNEL_TIMING, 1)
    if os.environ.get('GRID_BLOCKS'): ctx.set_option(icp.OPT_GRID_BLOCKS, int(os.environ['GRID_BLOCKS']))
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    for _ in range(reps):
        Tm, st, rc = ctx.register_symm(50, 1.0)
    print(st, wl.rotation_angle(Tm, T))
