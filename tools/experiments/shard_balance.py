"""Per-rank, per-search kernel times of the query-sharded 16M registration (stream mode with kernel timing):
how much of the sharded run's fixed cost is imbalance between the ranks (every iteration waits for the slowest).
  torchrun --nproc-per-node N tools/experiments/shard_balance.py [points]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import numpy as np, torch, torch.distributed as dist
import icpsymm as icp, workloads as wl, multigpu as mg
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_777_216
src, tgt, T = wl.config_synth(n)
with icp.Context(lr) as ctx:
    ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    res = {}
    for mode in ("single", "sharded"):
        if mode == "sharded":
            mg.attach_peers(ctx, dist, torch, rank, world)
        else:
            ctx.set_query_shard(0, 1)
        for timing in (0, 1):
            ctx.set_option(icp.OPT_KERNEL_TIMING, timing)
            for _ in range(2):
                dist.barrier()
                Tm, st, _ = ctx.register_symm(50, 1.0)
            if timing:
                a, b = ctx.search_times()
                res[mode] = (st["loop_ms"], a, b)
            else:
                res[mode + "_graph"] = st["loop_ms"]
    mg.detach_peers(ctx, dist)
    s_ms, s_a, s_b = res["single"]; p_ms, p_a, p_b = res["sharded"]
    t = torch.tensor(np.concatenate([[p_ms, res["sharded_graph"]], p_a, p_b]), device="cuda", dtype=torch.float64)
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    if rank == 0:
        k = len(p_a)
        A = np.stack([x.cpu().numpy() for x in allr])
        print(f"single GPU: graph {res['single_graph']:.2f} ms, stream {s_ms:.2f} ms; searches sum {s_a.sum():.2f}, solves sum {s_b.sum():.3f}")
        print(f"sharded over {world}: graph {A[:,1].max():.2f} ms, stream {A[:,0].max():.2f} ms")
        sa = A[:, 2:2 + k]; sb = A[:, 2 + k:]
        print("search k: single/world | per-rank sharded ms | max-min")
        for j in range(k):
            print(f"{j:2d}: {s_a[j]/world:7.3f} | " + " ".join(f"{x:7.3f}" for x in sa[:, j]) + f" | {sa[:, j].max()-sa[:, j].min():6.3f}   solve " + " ".join(f"{x*1e3:5.0f}us" for x in sb[:, j]))
        print(f"sum of max-over-ranks searches {sa.max(axis=0).sum():.2f} ms vs single/world {s_a.sum()/world:.2f}; sum of per-iteration imbalance {(sa.max(axis=0)-sa.min(axis=0)).sum():.2f} ms; solves (max) {sb.max(axis=0).sum():.3f} ms")
dist.destroy_process_group()
