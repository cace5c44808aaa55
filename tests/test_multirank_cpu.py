"""world_size-2 gloo tests (CPU) of the multi-rank host logic:
query sharding -- per-shard normal-equation vectors all-reduce to the single-rank vector and give
the same solve on every rank; pair batches -- contiguous, disjoint, complete partitions; the
128-byte communicator id travels from rank 0 to the others."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "icp-symm_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import workloads as wl  # noqa: E402


def test_shard_ranges_partition():
    for n in (0, 1, 7, 4096, 1_000_003):
        for world in (1, 2, 3, 8):
            parts = [wl.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gold = os.path.join(ROOT, "tests", "golden")
        cat, cat_out = np.load(os.path.join(gold, "cat_xyz.npy")), np.load(os.path.join(gold, "cat_out_xyz.npy"))
        g = dict(np.load(os.path.join(gold, "cat_golden.npz")))
        ns, nt = g["normals_cat"], g["normals_cat_out"]
        # rank 0 makes the communicator id, the others receive it (what bench.py does for NCCL)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(np.random.default_rng(5).integers(0, 256, 128, dtype=np.uint8))
        dist.broadcast(uid, 0)
        # query sharding: each rank owns a contiguous slice of the source
        lo, hi = wl.shard_range(len(cat), rank, world)
        idx, _ = orc.nn_brute(cat[lo:hi], cat_out)
        origin = (0.0, 0.0, 0.0)
        S = orc.symm_accumulate(cat[lo:hi], ns[lo:hi], cat_out[idx], nt[idx], origin)
        t = torch.from_numpy(S.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        T, x, deg = orc.solve_compose(t.numpy(), origin)
        # pair-batch data parallel: disjoint blocks of pairs, results gathered on every rank
        plo, phi = wl.shard_range(10, rank, world)
        mine = torch.tensor([float(b) for b in range(plo, phi)], dtype=torch.float64)
        pad = torch.full((5,), -1.0, dtype=torch.float64); pad[: len(mine)] = mine
        allp = [torch.zeros(5, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allp, pad)
        q.put((rank, uid.numpy().tobytes(), t.numpy().copy(), T.copy(), torch.cat(allp).numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_gloo_allreduce_of_normal_equations():
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    gold = os.path.join(ROOT, "tests", "golden")
    cat, cat_out = np.load(os.path.join(gold, "cat_xyz.npy")), np.load(os.path.join(gold, "cat_out_xyz.npy"))
    g = dict(np.load(os.path.join(gold, "cat_golden.npz")))
    idx, _ = orc.nn_brute(cat, cat_out)
    whole = orc.symm_accumulate(cat, g["normals_cat"], cat_out[idx], g["normals_cat_out"][idx], (0, 0, 0))
    assert res[0][1] == res[1][1]                                   # same communicator id everywhere
    assert np.array_equal(res[0][2], res[1][2])                     # bit-identical reduced vector on every rank
    assert np.array_equal(res[0][3], res[1][3])                     # hence a bit-identical transform
    assert np.allclose(res[0][2], whole, rtol=1e-12, atol=1e-9 * np.abs(whole).max())
    assert res[0][2][36] == len(cat)
    pairs = res[0][4]
    assert sorted(int(v) for v in pairs if v >= 0) == list(range(10))
