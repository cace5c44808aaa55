import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import numpy as np
import icpsymm as icp
import workloads as wl
src, tgt, T = wl.config_synth(1_000_000)
with icp.Context(0) as ctx:
    ctx.set_source(src); ctx.set_target(tgt)
    for rep in range(2):
        ctx.estimate_normals(10); ctx.synchronize()
    nb = 7813
    buf = (C.c_ulonglong * (3 * nb))()
    ctx.L.icpsymm_debug_block_times(buf, nb)
    a = np.array(list(buf), dtype=np.float64).reshape(nb, 3)
    t0 = a[:, 0].min(); st = (a[:, 0] - t0) / 1e3; en = (a[:, 1] - t0) / 1e3; dur = en - st
    print("kernel span %.1f us; block duration us: min %.1f median %.1f p90 %.1f p99 %.1f max %.1f" % (en.max(), dur.min(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
    # per-SM finish time
    sm = a[:, 2].astype(int)
    fin = np.array([en[sm == s].max() for s in np.unique(sm)])
    print("per-SM last finish us: min %.1f median %.1f max %.1f; SMs %d" % (fin.min(), np.median(fin), fin.max(), len(fin)))
    busy = np.array([dur[sm == s].sum() for s in np.unique(sm)])
    print("per-SM sum of block durations us: min %.0f median %.0f max %.0f" % (busy.min(), np.median(busy), busy.max()))
    # duration by block index decile
    for d in range(10):
        sl = slice(d * nb // 10, (d + 1) * nb // 10)
        print("blocks %5d..%5d: mean dur %.1f us, start %.0f..%.0f" % (sl.start, sl.stop, dur[sl].mean(), st[sl].min(), st[sl].max()))
    worst = np.argsort(-dur)[:10]
    print("slowest blocks:", [(int(b), round(float(dur[b]), 1)) for b in worst])
    n = 1_000_000
    pitch = (n + 7) // 8 * 8
    pb = np.empty((pitch, 4), np.float32)
    ctx.L.icpsymm_debug_sorted_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    ctx.L.icpsymm_debug_sorted_points(ctx.h, 1, pb.ctypes.data_as(C.c_void_p), pitch)   # last estimate ran on target
    def warp_diag(b):
        out = []
        for w in range(4):
            q = pb[b * 128 + w * 32: b * 128 + w * 32 + 32, :3]
            out.append(float(np.linalg.norm(q.max(0) - q.min(0))))
        return out
    alld = np.array([warp_diag(b) for b in range(nb - 1)])
    print("warp patch diag: median %.4f p99 %.4f max %.4f" % (np.median(alld), np.percentile(alld, 99), alld.max()))
    for b in worst:
        print("block", int(b), "dur %.1f" % dur[b], "warp diags", np.round(warp_diag(int(b)), 4))
    big = np.argsort(-alld.max(1))[:10]
    print("blocks with the largest warp diag:", [(int(b), round(float(alld[b].max()), 3), round(float(dur[b]), 1)) for b in big])
