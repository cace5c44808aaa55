import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import numpy as np
import icpsymm as icp
import workloads as wl
src, tgt, T = wl.config_synth(1_000_000)
with icp.Context(0) as ctx:
    ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
    ctx.set_option(icp.OPT_USE_GRAPH, 0)
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    nb = 2960
    pitch = 1_000_000
    pb = np.empty((pitch, 4), np.float32)
    ctx.L.icpsymm_debug_sorted_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    ctx.L.icpsymm_debug_sorted_points(ctx.h, 0, pb.ctypes.data_as(C.c_void_p), pitch)
    for iters in (0, 1, 6):
        for rep in range(2):
            Tm, st, rc = ctx.register_symm(iters, 1.0)
        buf = (C.c_ulonglong * (3 * nb))()
        ctx.L.icpsymm_debug_k1_times(buf, nb)
        a = np.array(list(buf), dtype=np.float64).reshape(nb, 3)
        t0 = a[:, 0].min(); stt = (a[:, 0] - t0) / 1e3; en = (a[:, 1] - t0) / 1e3; dur = en - stt
        two = np.arange(nb) < (3907 - nb)
        sm = a[:, 2].astype(int)
        fin = np.array([en[sm == s].max() for s in np.unique(sm)])
        print(f"search #{iters+1}: span {en.max():.0f} us; per-SM last finish min {fin.min():.0f} median {np.median(fin):.0f} max {fin.max():.0f}")
        for name, sel in (("2-tile", two), ("1-tile", ~two)):
            d = dur[sel]
            print(f"   {name} blocks: n {sel.sum()} dur us median {np.median(d):.0f} p90 {np.percentile(d,90):.0f} p99 {np.percentile(d,99):.0f} max {d.max():.0f}; last start {stt[sel].max():.0f}")
        worst = np.argsort(-en)[:6]
        def diag(b):
            out = []
            for tile in ([b, b + nb] if b + nb < 3907 else [b]):
                for w in range(8):
                    q = pb[tile * 256 + w * 32: tile * 256 + w * 32 + 32, :3]
                    if len(q): out.append(float(np.linalg.norm(q.max(0) - q.min(0))))
            return max(out)
        print("   last finishing blocks:", [(int(b), f"start {stt[b]:.0f} dur {dur[b]:.0f} maxwarpdiag {diag(int(b)):.3f}") for b in worst])
