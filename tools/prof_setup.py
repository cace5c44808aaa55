"""Small driver for ncu: index build + normal estimation of the 1M-point pair (config 3); prints host-clock times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import icpsymm as icp
import workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
src, tgt, T = wl.config_synth(n)
with icp.Context(0) as ctx:
    tb, tn = [], []
    for _ in range(reps):
        ctx.synchronize(); t0 = time.perf_counter()
        ctx.set_source(src); ctx.set_target(tgt); ctx.synchronize(); t1 = time.perf_counter()
        ctx.estimate_normals(10); ctx.synchronize(); t2 = time.perf_counter()
        tb.append(t1 - t0); tn.append(t2 - t1)
    print(f"ok  upload+index of both clouds {min(tb)*1e3:.3f} ms, normals of both clouds {min(tn)*1e3:.3f} ms (host clock, best of {reps})")
