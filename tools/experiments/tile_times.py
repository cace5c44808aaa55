"""Per-tile duration of the first (far) search of the 1M pair against the tile's mean NN distance.
Needs a build with -DICP_TILE_TIMES -DICP_BLOCK_TIMES:  python tools/experiments/tile_times.py <lib.so>"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import numpy as np
import icpsymm as icp
icp.LIB_PATH = os.path.abspath(sys.argv[1])
import workloads as wl
n = 1_000_000
src, tgt, T = wl.config_synth(n)
with icp.Context(0) as ctx:
    L = ctx.L
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    ctx.set_option(icp.OPT_TILE_ORDER, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    for _ in range(2):
        idx, d2 = ctx.find_correspondences()
    tiles = (n + 255) // 256
    if hasattr(L, "icpsymm_debug_heavy"):
        L.icpsymm_debug_heavy.argtypes = [C.c_void_p]
        print("heavy tiles (walked in shares):", L.icpsymm_debug_heavy(ctx.h))
    tt = np.zeros(2 * tiles, np.uint64)
    L.icpsymm_debug_tile_times(tt.ctypes.data_as(C.c_void_p), tiles)
    sp = np.zeros((((n + 7) // 8) * 8, 4), np.float32)
    L.icpsymm_debug_sorted_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.icpsymm_debug_sorted_points(ctx.h, 0, sp.ctypes.data_as(C.c_void_p), len(sp))
    orig = sp[:n, 3].copy().view(np.uint32)
    d = np.sqrt(d2[orig].astype(np.float64))
    t0, t1 = tt[0::2].astype(np.float64), tt[1::2].astype(np.float64)
    dur = (t1 - t0) * 1e-3
    start = (t0 - t0.min()) * 1e-3
    md = np.array([d[k * 256:(k + 1) * 256].mean() for k in range(tiles)])
    mx = np.array([d[k * 256:(k + 1) * 256].max() for k in range(tiles)])
    mn = np.array([d[k * 256:(k + 1) * 256].min() for k in range(tiles)])
    xyz = sp[:n, :3]
    diag = np.array([np.linalg.norm(xyz[k * 256:(k + 1) * 256].max(0) - xyz[k * 256:(k + 1) * 256].min(0)) for k in range(tiles)])
    udiag = np.array([max(np.linalg.norm(xyz[u * 32:(u + 1) * 32].max(0) - xyz[u * 32:(u + 1) * 32].min(0)) for u in range(k * 8, min(k * 8 + 8, (n + 31) // 32))) for k in range(tiles)])
    for name, f in (("mean d", md), ("max d", mx), ("min d", mn), ("tile diag", diag), ("max unit diag", udiag), ("max d * unit diag", mx * udiag)):
        print(f"corr(duration, {name}) = {np.corrcoef(dur, f)[0, 1]:.3f}")
    top = np.argsort(-dur)[:12]
    print("slowest tiles: idx start dur mean_d max_d tile_diag max_unit_diag")
    for k in top:
        print(k, round(start[k], 1), round(dur[k], 1), round(md[k], 3), round(mx[k], 3), round(diag[k], 3), round(udiag[k], 3))
    print("medians:", np.median(dur), np.median(md), np.median(diag), np.median(udiag))
    print("kernel span us", (t1.max() - t0.min()) * 1e-3, "sum of tile durations / 592 us", dur.sum() / 592)
    print("corr(duration, mean NN distance)", np.corrcoef(dur, md)[0, 1])
    print("tile  start_us  dur_us  mean_dist   (every 100th tile)")
    for k in range(0, tiles, 100):
        print(f"{k:5d} {start[k]:8.1f} {dur[k]:7.1f} {md[k]:.4f}")
    q = np.quantile(dur, [0.05, 0.25, 0.5, 0.75, 0.95, 1.0])
    print("duration quantiles us", q)
    # idle tail: time after which fewer than 592 tiles are in flight
    ends = np.sort(t1 - t0.min()) * 1e-3
    print("last tile starts at us", start.max(), " kernel end", ends[-1])
    np.save(os.path.join(ROOT, "gpurun_out", "tile_times.npy"), np.stack([start, dur, md]))
