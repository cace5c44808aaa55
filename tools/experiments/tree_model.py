"""CPU model of the search cost for alternative orders of the target (no GPU needed).

Input: gpurun_out/sorted_tgt.npy / sorted_src.npy (dumped on the GPU box with the debug build: the
curve order + 512-point k-d refinement).  For every candidate order the implicit 8-ary hierarchy is
built (groups of 8, 64, 512, 4096, ... consecutive points) and, for a sample of 32-query warps of
the source at a few poses, the number of groups of every level whose box is reached by ANY lane's
exact ball (radius = true NN distance) is counted -- the quantity the warp-synchronous walk pays for."""
import sys
import numpy as np
from scipy.spatial import cKDTree
sys.path.insert(0, '/root/repo/icp-symm_b200')
import workloads as wl

tgt0 = np.load('/root/repo/gpurun_out/sorted_tgt.npy').astype(np.float64)
src = np.load('/root/repo/gpurun_out/sorted_src.npy').astype(np.float64)
n = len(tgt0)
_, _, T = wl.config_synth(n)
tree = cKDTree(tgt0)
rng = np.random.default_rng(0)


def kd_partition(P, window, stop, axis_rule="bbox"):
    """median splits of every aligned window down to parts of `stop` points (stable), like kd_coarse_kernel"""
    P = P.copy()
    for base in range(0, len(P) - window + 1, window):
        stack = [(base, window)]
        while stack:
            b, s = stack.pop()
            if s <= stop:
                continue
            X = P[b:b + s]
            if axis_rule == "bbox":
                ax = int(np.argmax(X.max(0) - X.min(0)))
            else:
                ax = int(np.argmax(X.var(0)))
            order = np.argsort(X[:, ax], kind="stable")
            low = np.sort(order[:s // 2]); high = np.sort(order[s // 2:])      # stable partition
            P[b:b + s] = np.concatenate([X[low], X[high]])
            stack.append((b, s // 2)); stack.append((b + s // 2, s // 2))
    return P


def refine512(P):
    """full k-d refinement of every 512 window down to 8 (what kd_refine_kernel produces)"""
    P = P.copy()
    for base in range(0, len(P) - 511, 512):
        stack = [(base, 512)]
        while stack:
            b, s = stack.pop()
            if s <= 8:
                continue
            X = P[b:b + s]
            ax = int(np.argmax(X.max(0) - X.min(0)))
            order = np.argsort(X[:, ax], kind="stable")
            P[b:b + s] = X[order]
            stack.append((b, s // 2)); stack.append((b + s // 2, s // 2))
    return P


def level_boxes(P):
    out = []
    gs = 8
    while gs <= len(P):
        ng = len(P) // gs
        G = P[:ng * gs].reshape(ng, gs, 3)
        out.append((gs, G.min(1), G.max(1)))
        gs *= 8
    return out


def cost(P, poses, nw=400):
    boxes = level_boxes(P)
    res = {}
    for name, Q_all in poses.items():
        ws = rng.choice(n // 32 - 1, nw, replace=False)
        tot = np.zeros(len(boxes))
        for w in ws:
            Q = Q_all[w * 32:(w + 1) * 32]
            d, _ = tree.query(Q)
            for li, (gs, lo, hi) in enumerate(boxes):
                ex = np.maximum(np.maximum(lo[None] - Q[:, None], Q[:, None] - hi[None]), 0)
                tot[li] += (np.sqrt((ex * ex).sum(2)) <= d[:, None]).any(0).sum()
        res[name] = tot / nw
    return res


def pose(deg):
    ax = np.array([1., 2, 3]); ax /= np.linalg.norm(ax); th = np.radians(deg)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    aligned = src @ T[:3, :3].T + T[:3, 3]
    c = aligned.mean(0)
    return (aligned - c) @ R.T + c + np.array([0.1 * 2.74 * deg / 30, 0, 0])


poses = {"converged": pose(0.0), "3deg": pose(3.0), "30deg": pose(30.0)}
cands = {"curve + kd512 (round-1 start)": tgt0}
for window in (8192, 65536):
    cands[f"kd partition {window} (bbox axis) + kd512"] = refine512(kd_partition(tgt0, window, 512))
cands["kd partition 8192 (variance axis) + kd512"] = refine512(kd_partition(tgt0, 8192, 512, "var"))
for name, P in cands.items():
    r = cost(P, poses)
    print(name)
    for pn, v in r.items():
        print(f"   {pn:10s} groups reached per warp by level (8,64,512,4096,...): {np.round(v, 2)}  weighted ~{v[0]*90+v[1:].sum()*120:.0f} instr")
