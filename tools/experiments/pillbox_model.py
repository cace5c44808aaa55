"""Model: how many leaf groups does an exact search have to open with AABB bounds vs oriented pillbox bounds?"""
import sys, numpy as np
sys.path.insert(0, '/root/repo/icp-symm_b200')
import workloads as wl
from scipy.spatial import cKDTree
N = 1_000_000
src, tgt, Ttrue = wl.config_synth(N)
tgt = tgt.astype(np.float64); src = src.astype(np.float64)
# k-d partition of the target into leaves of 8 (2^17 leaves after padding)
def kd_leaves(P, leaf=8):
    idx = np.arange(len(P)); out = []
    stack = [idx]
    while stack:
        ids = stack.pop()
        if len(ids) <= leaf: out.append(ids); continue
        pts = P[ids]; ax = np.argmax(pts.max(0) - pts.min(0))
        h = len(ids) // 2
        part = np.argpartition(pts[:, ax], h)
        stack.append(ids[part[:h]]); stack.append(ids[part[h:]])
    return out
def groups(P, size):
    return kd_leaves(P, size)
tree = cKDTree(tgt)
rng = np.random.default_rng(1)
for size in (8, 64, 512):
    L = groups(tgt, size)
    nl = len(L)
    lo = np.array([tgt[i].min(0) for i in L]); hi = np.array([tgt[i].max(0) for i in L])
    c = np.empty((nl, 3)); nrm = np.empty((nl, 3)); w = np.empty(nl); R = np.empty(nl)
    for j, i in enumerate(L):
        X = tgt[i]; m = X.mean(0); D = X - m
        ev, evec = np.linalg.eigh(D.T @ D)
        n = evec[:, 0]; h = D @ n
        # recentre along the normal so the slab is symmetric
        mid = 0.5 * (h.max() + h.min()); m2 = m + mid * n; D = X - m2; h = D @ n
        lat = np.sqrt(np.maximum((D * D).sum(1) - h * h, 0))
        c[j] = m2; nrm[j] = n; w[j] = np.abs(h).max(); R[j] = lat.max()
    print(f"group size {size}: {nl} groups, mean AABB diag {np.linalg.norm(hi-lo,axis=1).mean():.4f}, mean w {w.mean():.5f}, mean R {R.mean():.4f}")
    for deg in (30, 10, 3, 1, 0.3, 0.0):
        # residual pose: rotation deg about (1,2,3) through centroid + proportional translation
        ax = np.array([1., 2, 3]); ax /= np.linalg.norm(ax); th = np.radians(deg)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        Rm = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        cen = src.mean(0)
        # queries = true-aligned source (tgt without noise ~ use tgt itself moved by residual)
        qi = rng.choice(N, 600, replace=False)
        base = tgt[qi] + rng.normal(0, 1e-4 * 2.74, (600, 3))
        Q = (base - cen) @ Rm.T + cen + np.array([0.1 * 2.74 * deg / 30, 0, 0])
        dnn, _ = tree.query(Q)
        ca = cp = cb = 0
        for q, d in zip(Q, dnn):
            ex = np.maximum(np.maximum(lo - q, q - hi), 0); la = np.sqrt((ex * ex).sum(1))
            D = q - c; h = (D * nrm).sum(1); l2 = np.maximum((D * D).sum(1) - h * h, 0)
            lp = np.sqrt(np.maximum(np.abs(h) - w, 0) ** 2 + np.maximum(np.sqrt(l2) - R, 0) ** 2)
            ca += (la <= d).sum(); cp += (lp <= d).sum(); cb += (np.maximum(la, lp) <= d).sum()
        print(f"  residual {deg:5.1f} deg: mean d_nn {dnn.mean():.4f}  groups reached: AABB {ca/600:7.1f}  pillbox {cp/600:7.1f}  both {cb/600:7.1f}")
