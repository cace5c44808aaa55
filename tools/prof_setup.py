"""Small driver for ncu: index build + normal estimation of the 1M-point pair (config 3)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import icpsymm as icp
import workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
src, tgt, T = wl.config_synth(n)
with icp.Context(0) as ctx:
    for _ in range(2):
        ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    print("ok")
