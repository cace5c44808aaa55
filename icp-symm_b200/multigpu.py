"""Host-side plumbing of the two multi-GPU modes, shared by bench.py and tests/test_gpu_multi.py.

One process per GPU; `torch.distributed` only carries the 64-byte IPC handles / the NCCL id between the
ranks and the comparisons of the results -- the data path is the library's own (peer stores fused into
the reduce+solve kernel, or ncclAllReduce of exact-zero-padded part sums).
"""
from __future__ import annotations

import numpy as np


def attach_peers(ctx, dist, torch, rank, world):
    """NVLink peer exchange: export my handle, gather all, attach; collective, ends with a barrier"""
    mine = torch.frombuffer(bytearray(ctx.peer_export()), dtype=torch.uint8).cuda()
    table = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(table, mine)
    ctx.peer_attach(b"".join(bytes(t.cpu().numpy().tobytes()) for t in table), world, rank)
    dist.barrier()          # nobody starts a sharded registration before every rank is attached


def detach_peers(ctx, dist):
    dist.barrier()          # nobody unmaps a buffer a peer may still store into
    ctx.peer_detach()
    dist.barrier()


def init_nccl(ctx, dist, torch, rank, world):
    import icpsymm as icp
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(icp.comm_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), world, rank)
    dist.barrier()


def same_on_all_ranks(dist, torch, arr) -> bool:
    """bit-equality of a small array across ranks"""
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).copy()).cuda()
    ref = t.clone()
    dist.broadcast(ref, 0)
    flag = torch.tensor([1.0 if torch.equal(ref, t) else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item() == 1.0)


def union_of_shards(dist, torch, idx):
    """every rank holds original target indices for its own shard and -1 elsewhere: element-wise max = union;
    also returns how many queries were claimed by more than one rank (must be 0)"""
    t = torch.from_numpy(idx).cuda()
    owned = (t >= 0).to(torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    return t.cpu().numpy(), int((owned != 1).sum().item())


def sharded_vs_single(ctx, dist, torch, rank, world, src, tgt, max_iters, thr, rel_tol, steps=1, warmup=1,
                      check_nccl=True):
    """Register one pair query-sharded over all ranks (peer exchange, and NCCL) and unsharded on every rank;
    returns a dict with timings and the parity verdict.  Clouds must be identical on all ranks."""
    import icpsymm as icp
    ctx.set_option(icp.OPT_REL_DIFF_TOL, rel_tol)
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    out = {}

    def run(n_warm, n_steps):
        ms = []
        for k in range(n_warm + n_steps):
            if world > 1:
                dist.barrier()
            T, st, _ = ctx.register_symm(max_iters, thr)
            if k >= n_warm:
                ms.append(st["loop_ms"])
        return T, st, ms

    # unsharded: every rank registers the whole pair on its own (also the 1-GPU time of this box)
    ctx.set_query_shard(0, 1)
    T1, st1, ms1 = run(warmup, steps)
    idx1, d21 = ctx.get_correspondences()
    out["single"] = {"ms": float(np.mean(ms1)), "iters": st1["iters"], "searches": st1["searches"], "diff": st1["diff"],
                     "rmse": st1["rmse"]}
    out["T_single"] = T1
    if world == 1:
        return out
    # peer exchange fused into the reduce+solve kernel, device-side loop graph
    attach_peers(ctx, dist, torch, rank, world)
    Tp, stp, msp = run(warmup, steps)
    idxp, _ = ctx.get_correspondences()
    detach_peers(ctx, dist)
    union, clashes = union_of_shards(dist, torch, idxp)
    t = torch.tensor([float(np.mean(msp))], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["peer"] = {"ms": float(t.item()), "iters": stp["iters"], "diff": stp["diff"], "rmse": stp["rmse"]}
    checks = {
        "peer_T_equals_single_gpu_T_bitwise": bool(np.array_equal(Tp, T1)),
        "peer_T_identical_on_all_ranks": same_on_all_ranks(dist, torch, Tp),
        "union_of_shard_indices_equals_single_gpu_indices": bool(np.array_equal(union, idx1)) and clashes == 0,
        "iters_diff_rmse_equal": bool(stp["iters"] == st1["iters"] and stp["diff"] == st1["diff"] and stp["rmse"] == st1["rmse"]),
    }
    if check_nccl:
        init_nccl(ctx, dist, torch, rank, world)
        Tn, stn, msn = run(warmup, steps)
        dist.barrier()
        ctx.comm_destroy()
        t = torch.tensor([float(np.mean(msn))], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["nccl"] = {"ms": float(t.item()), "iters": stn["iters"]}
        checks["nccl_T_equals_single_gpu_T_bitwise"] = bool(np.array_equal(Tn, T1))
    ok = torch.tensor([1.0 if all(checks.values()) else 0.0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["checks"] = checks
    out["parity"] = "bit-equal" if ok.item() == 1.0 else "MISMATCH: " + ", ".join(k for k, v in checks.items() if not v)
    return out
