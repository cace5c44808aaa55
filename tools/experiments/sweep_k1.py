"""Per-search kernel times and whole-loop time of the 1M registration for every ICP_EXPERIMENT build in
tools/_scratch/variants/*.so (make -C icp-symm_b200 exp TAG=_x EXTRA=-D...), plus a checksum of the final
correspondences and of the transform bits: every variant must print the same two checksums.

  python tools/experiments/sweep_k1.py [n_points] [lib ...]
"""
import ctypes as C
import glob
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(lib, n):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
    import numpy as np
    import torch
    import icpsymm as icp
    icp.LIB_PATH = lib
    import workloads as wl
    src, tgt, T = wl.config_synth(n)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with icp.Context(0) as ctx:
        L = ctx.L
        stream = torch.cuda.ExternalStream(ctx.stream)
        ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
        if os.environ.get("PROBE"):
            ctx.set_option(icp.OPT_PROBE, int(os.environ["PROBE"]))
        if os.environ.get("TILE_ORDER"):
            ctx.set_option(icp.OPT_TILE_ORDER, int(os.environ["TILE_ORDER"]))
        import time
        tn = []
        for _ in range(5):
            ctx.set_source(src); ctx.set_target(tgt); ctx.synchronize(); t1 = time.perf_counter()
            ctx.estimate_normals(10); ctx.synchronize(); tn.append(time.perf_counter() - t1)
        nrm = ctx.get_normals(1)
        h3 = hashlib.md5(nrm.tobytes()).hexdigest()[:8]
        ms = []
        for _ in range(8):
            with torch.cuda.stream(stream):
                flush.zero_()
            Tm, st, rc = ctx.register_symm(50, 1.0)
            ms.append(st["loop_ms"])
        idx, d2 = ctx.get_correspondences()
        h1 = hashlib.md5(idx.tobytes() + d2.tobytes()).hexdigest()[:8]
        h2 = hashlib.md5(np.ascontiguousarray(Tm).tobytes()).hexdigest()[:8]
        per = ""
        ctx.set_option(icp.OPT_KERNEL_TIMING, 1)
        runs = []
        for _ in range(int(os.environ.get("REPS", "10"))):
            with torch.cuda.stream(stream):
                flush.zero_()
            ctx.register_symm(50, 1.0)
            cur, sol = ctx.search_times()
            runs.append(cur)
        runs = np.array(runs)
        best, med = runs.min(axis=0), np.median(runs, axis=0)
        per = (" searches[" + " ".join(f"{x:.3f}" for x in best) + "] med[" + " ".join(f"{x:.3f}" for x in med)
               + f"] sum {best.sum():.3f}/{med.sum():.3f} solve {float(sol.mean())*1e3:.1f}us")
        ms = sorted(ms[2:])
        print(f"{os.path.basename(lib):28s} loop best {ms[0]:.3f} med {ms[len(ms)//2]:.3f} iters {st['iters']}{per} normals(both) {min(tn)*1e3:.3f} corr {h1} T {h2} nrm {h3}",
              flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
    else:
        n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1_000_000
        libs = [a for a in sys.argv[1:] if a.endswith(".so")] or sorted(glob.glob(os.path.join(ROOT, "tools/_scratch/variants/*.so")))
        for l in libs:
            subprocess.run([sys.executable, __file__, "--child", os.path.abspath(l), str(n)])
