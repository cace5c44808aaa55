import os, sys, glob, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
    import icpsymm as icp
    icp.LIB_PATH = sys.argv[1]
    import workloads as wl
    src, tgt, T = wl.config_synth(1_000_000)
    with icp.Context(0) as ctx:
        ctx.set_source(src); ctx.set_target(tgt)
        tn = []
        for _ in range(10):
            ctx.synchronize(); t1 = time.perf_counter()
            ctx.estimate_normals(10); ctx.synchronize(); t2 = time.perf_counter()
            tn.append(t2 - t1)
        print(os.path.basename(sys.argv[1]), "normals both clouds best %.3f ms median %.3f" % (min(tn)*1e3, sorted(tn)[5]*1e3), flush=True)
else:
    libs = sorted(glob.glob(os.path.join(ROOT, "tools/_scratch/variants/*.so"))) + [os.path.join(ROOT, "icp-symm_b200/libicpsymm_b200.so")]
    for l in libs:
        subprocess.run([sys.executable, __file__, l])
