"""How much of the converged search is caused by 'fat' boxes (groups that straddle a jump of the curve)?"""
import numpy as np
from scipy.spatial import cKDTree
tgt = np.load('/root/repo/gpurun_out/sorted_tgt.npy').astype(np.float64)
src = np.load('/root/repo/gpurun_out/sorted_src.npy').astype(np.float64)
n = len(tgt)
import sys; sys.path.insert(0, '/root/repo/icp-symm_b200')
import workloads as wl
_, _, T = wl.config_synth(n)
srcT = src @ T[:3, :3].T + T[:3, 3]     # aligned source (converged pose)
tree = cKDTree(tgt)
rng = np.random.default_rng(0)
for gs in (8, 64, 512):
    ng = n // gs
    G = tgt[:ng * gs].reshape(ng, gs, 3)
    lo = G.min(1); hi = G.max(1)
    diag = np.linalg.norm(hi - lo, axis=1)
    med = np.median(diag)
    fat = diag > 3 * med
    print(f"group size {gs}: {ng} groups, median diag {med:.4f}, fat (>3x median) {fat.sum()} = {100*fat.mean():.2f} %, p99 {np.percentile(diag,99):.4f}, max {diag.max():.3f}")
    # warp-level: 32 consecutive source points = one patch; ball radius = each query's NN distance
    nw = 1500
    ws = rng.choice(n // 32 - 1, nw, replace=False)
    cnt_norm = cnt_fat = 0
    for w in ws:
        Q = srcT[w * 32:(w + 1) * 32]
        d, _ = tree.query(Q)
        # a group is opened by the warp when any lane's ball reaches its box
        ex = np.maximum(np.maximum(lo[None] - Q[:, None], Q[:, None] - hi[None]), 0)   # 32 x ng x 3
        reach = (np.sqrt((ex * ex).sum(2)) <= d[:, None]).any(0)
        cnt_fat += (reach & fat).sum(); cnt_norm += (reach & ~fat).sum()
    print(f"   per warp (converged pose, exact bounds): groups reached {(cnt_norm+cnt_fat)/nw:.2f}, of which fat {cnt_fat/nw:.2f}")
