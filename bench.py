#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 symmetric-ICP path (contract: see README / DESIGN.md section 7).

Metric (BASELINE.json): correspondences/s of the registration loop (NN search + symmetric
accumulate + solve) on the synthetic 1M-point pair (config 3: bumpy ellipsoid, 30 deg /
0.1-extent perturbation, 1e-4-extent noise).  One "step" = one complete registration of that pair
from the initial perturbation until the loop stops (max 50 iterations, threshold 1.0, relative
diff tolerance 1e-5).  Counted work = N x (solves applied); the final convergence-check search is
timed but not counted.

  value : loop only, clouds + index + normals already resident in HBM (CUDA-graph mode)
  e2e   : the full MyICP::RegisterSymm-equivalent through the C ABI with HOST buffers --
          H2D of both clouds, index build, normal estimation, loop, D2H of the transform
  N > 1 : every rank registers its own copy of the pair (pair-batch data parallel, no collective);
          value = all ranks' correspondences / max-over-ranks time.  "scaling": "weak".

The same JSON line carries the two multi-GPU configurations of BASELINE.json as sub-objects, measured at
every N (so that SCALE holds them for N = 1, 2, 4, 8):
  shard16m : config 4 -- ONE 16M-point pair query-sharded over the N ranks (replicated clouds, part sums
             exchanged by NVLink peer stores inside the reduce+solve kernel), strong scaling, with the
             parity verdict of the sharded run against the unsharded run on the same GPU
  batch32k : config 5 -- 4096 independent 32K-point pairs, 4096/N per rank, strong scaling, pairs/s

`--impl reference` times the CPU oracle port (the reference itself cannot be built here: no
PCL/Eigen) on the same workload with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "icp-symm_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import workloads as wl  # noqa: E402

MAX_ITERS, DIFF_THR, REL_TOL = 50, 1.0, 1e-5
BYTES_PER_CORR_FUSED = 64          # fused NN+accumulate, M = N
BYTES_PER_POINT_NORMALS = 48       # 16 read + 16 write + 16 (sorted copy read once)
BYTES_PER_ROW_K2 = 320             # one fp64 row of 40 values per 256-query tile
METRIC, UNIT = "correspondences/sec (NN+symm solve)", "correspondences/s"
SHARD_POINTS, BATCH_PAIRS, BATCH_POINTS = 16 * 2 ** 20, 4096, 32768


def workload_name(n):
    return f"synth-{n} bumpy ellipsoid, 30deg/0.1-extent, noise 1e-4 (config 3)"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_profile():
    """committed summary of the `ncu --set full` capture of the dominant kernel (profiles/r02_nn_accum.json, else
    round 1's): dram bytes per launch and executed warp instructions per launch"""
    for name in ("r02_nn_accum.json", "r01_nn_accum.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
                d["file"] = "profiles/" + name
                return d
        except Exception:
            continue
    return {}


class ClockSampler:
    """samples nvidia-smi SM clocks + throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_threads():
    """all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which is not the machine)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle port on the host cores
# ------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """the reference's CPU path (oracle port: kd-tree leaf 15 + canonical accumulate/solve, OpenMP over queries).
    value = the loop (kd-tree build + iterations, normals given) -- like for like with the GPU arm's `value`;
    e2e = the whole MyICP::RegisterSymm (k=10 normals of both clouds + loop) -- like for like with its `e2e`."""
    if rank != 0:
        return
    from oracle import oracle as orc
    n = args.points
    src, tgt, _ = wl.config_synth(n)
    sample = min(n, args.cpu_sample or n)
    sel = np.sort(np.random.default_rng(7).choice(n, sample, replace=False)) if sample < n else np.arange(n)
    threads = host_threads()
    loop_s, whole_s, iters = [], [], 0
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        ns, _ = orc.estimate_normals(src[sel], 10, threads=threads)
        nt, _ = orc.estimate_normals(tgt, 10, threads=threads)
        t1 = time.perf_counter()
        r = orc.register_symm(src[sel], ns, tgt, nt, max_iters=MAX_ITERS, diff_threshold=DIFF_THR * sample / n,
                              rel_diff_tol=REL_TOL, threads=threads)
        t2 = time.perf_counter()
        if s >= args.warmup:
            loop_s.append(t2 - t1); whole_s.append(t2 - t0)
            iters = r["iters"]
    value = sample * iters * len(loop_s) / float(np.sum(loop_s))
    e2e = sample * iters * len(whole_s) / float(np.sum(whole_s))
    smp = (f"{sample} of {n} source points against the full {n}-point target; value = registration loop incl. kd-tree "
           "build (normals given), e2e = whole RegisterSymm (k=10 normals of both clouds + loop)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(loop_s)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n), "points": n, "max_iters": MAX_ITERS, "diff_threshold": DIFF_THR,
                       "rel_diff_tol": REL_TOL, "iters_run": iters},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": smp},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": 1e3 * float(np.mean(whole_s)), "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "note": "PCL-equivalent restatement (exact kd-tree, leaf 15, OpenMP); PCL itself is not installable here"}
    print(json.dumps(line), flush=True)


def cpu_baseline(args, src, tgt, ctx):
    """the oracle port on the host cores beside the GPU number, on bounded samples of the same
    workload (normals from the device run).  B1 = all cores (the headline `value` of this object), B0 = one thread
    (the reference is single-threaded, myicp.cpp:121-142), B2 = third-party search-only sanity legs."""
    from oracle import oracle as orc
    n = len(src)
    ns_all, nt = ctx.get_normals(0), ctx.get_normals(1)
    threads = host_threads()
    legs = []

    def oracle_leg(name, sample, thr):
        sel = np.sort(np.random.default_rng(7).choice(n, sample, replace=False)) if sample < n else np.arange(n)
        t0 = time.perf_counter()
        r = orc.register_symm(src[sel], ns_all[sel], tgt, nt, max_iters=MAX_ITERS, diff_threshold=DIFF_THR * sample / n,
                              rel_diff_tol=REL_TOL, threads=thr)
        dt = time.perf_counter() - t0
        return {"name": name, "value": sample * r["iters"] / dt, "unit": UNIT, "cores": thr, "seconds": dt, "iters": r["iters"],
                "sample": f"{sample} of {n} source points (random subset) against the full {n}-point target; whole "
                          "registration loop incl. kd-tree build (leaf 15); normals given"}

    b1 = oracle_leg("B1 oracle port, all host threads (OpenMP over queries)", min(n, args.cpu_sample or n), threads)
    legs.append(oracle_leg("B0 oracle port, 1 thread", min(n, max(1000, n // 8)), 1))
    q = np.ascontiguousarray(src[:: max(1, n // 250_000)])
    try:
        import cv2
        t0 = time.perf_counter()
        fl = cv2.flann_Index(tgt, dict(algorithm=4, leaf_max_size=15))       # KDTREE_SINGLE: pcl::KdTreeFLANN's index
        fl.knnSearch(q, 1, params=dict(checks=-1, eps=0.0, sorted=True))
        dt = time.perf_counter() - t0
        legs.append({"name": "B2 OpenCV FLANN KDTREE_SINGLE, search only (build + one exact 1-NN pass, unregistered pose)",
                     "value": len(q) / dt, "unit": UNIT, "cores": 1, "seconds": dt,
                     "sample": f"{len(q)} of {n} queries against the full target"})
    except Exception as e:   # third-party leg: absence is reported, not fatal
        legs.append({"name": "B2 OpenCV FLANN", "unavailable": str(e)[:100]})
    try:
        from scipy.spatial import cKDTree
        t0 = time.perf_counter()
        cKDTree(tgt, leafsize=15).query(np.ascontiguousarray(src[::2]), k=1, workers=-1)
        dt = time.perf_counter() - t0
        legs.append({"name": "B2 scipy cKDTree(workers=-1), search only (build + one exact 1-NN pass, unregistered pose)",
                     "value": len(src[::2]) / dt, "unit": UNIT, "cores": threads, "seconds": dt,
                     "sample": f"{len(src[::2])} of {n} queries against the full target"})
    except Exception as e:
        legs.append({"name": "B2 scipy cKDTree", "unavailable": str(e)[:100]})
    out = dict(b1)
    out["kind"] = "port"
    out["legs"] = legs
    return out


# ------------------------------------------------------------------------------------------------------
# sub-benchmarks: config 4 (one 16M pair, query-sharded) and config 5 (4096 x 32K pairs)
# ------------------------------------------------------------------------------------------------------
def shard_sub(args, torch, dist, icp, rank, world, local_rank):
    """config 4: strong scaling of ONE pair over the ranks + parity of the sharded run against the unsharded one"""
    import multigpu as mg
    n = args.shard_points
    t0 = time.perf_counter()
    src, tgt, Ttrue = wl.config_synth(n)               # identical on every rank (same seed)
    gen_s = time.perf_counter() - t0
    ctx = icp.Context(local_rank)
    try:
        r = mg.sharded_vs_single(ctx, dist, torch, rank, world, src, tgt, MAX_ITERS, DIFF_THR, REL_TOL,
                                 steps=max(2, args.steps // 3), warmup=2, check_nccl=True)
    finally:
        ctx.close()
    single = r["single"]
    T = r["T_single"]
    out = {"metric": METRIC, "unit": UNIT, "scaling": "strong", "points": n, "iters": single["iters"],
           "searches": single["searches"],
           "workload": f"synth-{n} pair (config 4), replicated clouds, queries sharded over {world} rank(s)",
           "single_gpu_ms": single["ms"], "single_gpu_value": float(n) * single["iters"] / (single["ms"] * 1e-3),
           "rot_err_rad": wl.rotation_angle(T, Ttrue), "trans_err_over_extent": float(np.abs(T[:3, 3] - Ttrue[:3, 3]).max() / wl.extent(tgt)),
           "host_generation_s": gen_s}
    # the DRAM-bound configuration (working set 1.2 GB >> 126 MB L2): algorithmic bytes against the measured peak, beside
    # the dram bytes of the committed ncu capture of this very workload (profiles/r02_nn_accum_16m.json)
    peak, peak_kind = measured_peak_gbs()
    try:
        with open(os.path.join(ROOT, "profiles", "r02_nn_accum_16m.json")) as f:
            prof16 = json.load(f)
    except Exception:
        prof16 = {}
    alg = BYTES_PER_CORR_FUSED * float(n) * single["searches"]
    out["roofline"] = {"bound": "hbm", "kernel": "nn_accum_kernel, all searches of the unsharded 16M registration (incl. reduce+solve launches)",
                       "achieved": alg / (single["ms"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                       "frac": alg / (single["ms"] * 1e-3) / 1e9 / peak, "peak_kind": f"of {peak_kind}",
                       "traffic": prof16.get("dram_bytes_per_launch_mean"), "algorithmic_bytes_per_launch": BYTES_PER_CORR_FUSED * n,
                       "traffic_source": "profiles/r02_nn_accum_16m.json (first 10 searches)" if prof16 else None}
    if world == 1:
        out.update({"value": out["single_gpu_value"], "ms_per_step": single["ms"], "parity": "n/a (one rank: nothing is sharded)"})
        return out
    peer, nccl = r["peer"], r.get("nccl")
    out.update({"value": float(n) * peer["iters"] / (peer["ms"] * 1e-3), "ms_per_step": peer["ms"],
                "efficiency_vs_1gpu": single["ms"] / (world * peer["ms"]),
                "exchange": "part sums stored into every peer's buffer over NVLink inside the reduce+solve kernel (device-side WHILE graph)",
                "parity": r["parity"], "checks": r["checks"]})
    if nccl:
        searches = max(1, single["searches"])
        out["nccl_variant"] = {"ms_per_step": nccl["ms"], "value": float(n) * nccl["iters"] / (nccl["ms"] * 1e-3),
                               "exchange": "ncclAllReduce of zero-padded part sums between a reduce-only and a solve-only launch (stream mode)",
                               "extra_us_per_iteration_vs_peer": 1e3 * (nccl["ms"] - peer["ms"]) / searches}
    return out


def batch_sub(args, torch, dist, icp, rank, world, local_rank, barrier):
    """config 5: 4096 independent 32K-point pairs, 4096/N per rank (strong scaling), no collective on the data path"""
    n, total = BATCH_POINTS, args.pairs
    lo, hi = wl.shard_range(total, rank, world)
    local = hi - lo
    distinct = min(local, 32)
    S = np.empty((local, n, 3), np.float32)
    T = np.empty((local, n, 3), np.float32)
    truth = []
    for j in range(distinct):
        s, t, Tt = wl.config_batch_pair(lo + j, n)
        S[j], T[j] = s, t
        truth.append(Tt)
    for j in range(distinct, local):          # the remaining pairs repeat the distinct ones (own memory, same work)
        S[j], T[j] = S[j % distinct], T[j % distinct]
    hS, hT = torch.from_numpy(S).pin_memory(), torch.from_numpy(T).pin_memory()
    thr = 1e-3
    ctx = icp.Context(local_rank)
    try:
        ctx.set_option(icp.OPT_REL_DIFF_TOL, REL_TOL)
        stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

        def upload():
            ctx.set_source_batch_ptr(hS.data_ptr(), local, n, 12)
            ctx.set_target_batch_ptr(hT.data_ptr(), local, n, 12)
            ctx.estimate_normals(10)

        upload()
        for _ in range(2):
            Tm, st, _ = ctx.register_symm(MAX_ITERS, thr)
        steps = max(2, args.steps // 3)
        barrier()
        loop_ms = []
        for _ in range(steps):
            Tm, st, _ = ctx.register_symm(MAX_ITERS, thr)
            loop_ms.append(st[0]["loop_ms"])
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_ms = []
        for _ in range(steps):
            ev0.record(stream)
            upload()
            Tm, st, _ = ctx.register_symm(MAX_ITERS, thr)
            ev1.record(stream)
            ev1.synchronize()
            e2e_ms.append(ev0.elapsed_time(ev1))
        barrier()
        # the same end-to-end work pipelined over two contexts with the asynchronous calls: the rank's pairs in chunks,
        # chunk k+1 is uploaded / indexed / given normals on the second context while chunk k iterates on the first
        chunk = max(1, min(512, local // 4))
        nchunks = (local + chunk - 1) // chunk
        pool = [ctx, icp.Context(local_rank)]
        pool[1].set_option(icp.OPT_REL_DIFF_TOL, REL_TOL)
        for c in pool:
            c.set_option(icp.OPT_ASYNC_SETUP, 1)

        def pipelined():
            for k in range(nchunks):
                c = pool[k % 2]
                if k >= 2:
                    c.wait()
                lo_k, n_k = k * chunk, min(chunk, local - k * chunk)
                c.set_source_batch_ptr(hS.data_ptr() + lo_k * n * 12, n_k, n, 12)
                c.set_target_batch_ptr(hT.data_ptr() + lo_k * n * 12, n_k, n, 12)
                c.estimate_normals(10)
                c.register_symm_async(MAX_ITERS, thr)
            last = None
            for k in range(max(0, nchunks - 2), nchunks):
                last = pool[k % 2].wait()
            return last

        pipe_ms = []
        try:
            pipelined()
            barrier()
            for _ in range(steps):
                torch.cuda.synchronize()
                ev0.record()
                pipelined()
                ev1.record()
                ev1.synchronize()
                pipe_ms.append(ev0.elapsed_time(ev1))
            barrier()
        finally:
            pool[1].close()
    finally:
        ctx.close()
    errs = [wl.rotation_angle(Tm[j], truth[j]) for j in range(distinct)]
    iters = [s["iters"] for s in st]
    v = torch.tensor([float(np.mean(loop_ms)), float(np.mean(e2e_ms)), float(np.max(errs)), float(np.max(iters)),
                      float(np.mean(pipe_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    ms, ms2, max_err, max_it, ms3 = [float(x) for x in v.tolist()]
    return {"metric": "registered pairs/sec", "unit": "pairs/s", "scaling": "strong", "pairs": total, "points_per_cloud": n,
            "workload": f"batch of {total} independent {n}-point pairs (config 5), {total}/{world} per rank, {distinct} distinct per rank",
            "value": total / (ms * 1e-3), "pairs_per_s": total / (ms * 1e-3), "ms_per_step": ms,
            "e2e_pairs_per_s": total / (ms2 * 1e-3), "e2e_ms_per_step": ms2,
            "e2e_pipelined_pairs_per_s": total / (ms3 * 1e-3), "e2e_pipelined_ms_per_step": ms3,
            "e2e_pipelined": f"same work in chunks of {chunk} pairs over two contexts with icpsymm_register_symm_async / icpsymm_wait "
                             "(upload + index + normals of chunk k+1 overlap the loop of chunk k); device-clock wall time",
            "e2e_includes": "H2D of both cloud batches from pinned memory, index build, kNN-10 normals, loop, D2H of transforms + stats",
            "h2d_bytes_per_step_per_rank": int(2 * local * n * 12), "d2h_bytes_per_step_per_rank": int(local * (64 + 72)),
            "diff_threshold": thr, "rel_diff_tol": REL_TOL, "iters_mean_rank0": float(np.mean(iters)), "iters_max": int(max_it),
            "max_rot_err_rad": max_err}


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--pairs", type=int, default=BATCH_PAIRS)
    ap.add_argument("--shard-points", type=int, default=SHARD_POINTS)
    ap.add_argument("--cpu-sample", type=int, default=0, help="source points per CPU step (0 = all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the shard16m / batch32k sub-benchmarks")
    ap.add_argument("--seed-per-rank", action="store_true", help="rank r registers the pair of seed 1234+r instead of the same pair")
    ap.add_argument("--seed-offset", type=int, default=0, help="added to the generator seed of the headline pair")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import icpsymm as icp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.points
    src, tgt, T_true = wl.config_synth(n, seed=1234 + args.seed_offset + (rank if args.seed_per_rank else 0))
    # pinned host buffers: the e2e path copies from these every step
    h_src = torch.from_numpy(src).pin_memory()
    h_tgt = torch.from_numpy(tgt).pin_memory()
    ctx = icp.Context(local_rank)
    ctx.set_option(icp.OPT_REL_DIFF_TOL, REL_TOL)
    ctx.set_option(icp.OPT_ASYNC_SETUP, 1)      # per-cloud streams: the target's H2D overlaps the source's build
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def l2_flush():
        with torch.cuda.stream(stream):
            flush.zero_()

    def setup_resident():
        ctx.set_source_ptr(h_src.data_ptr(), n, 12)
        ctx.set_target_ptr(h_tgt.data_ptr(), n, 12)
        ctx.estimate_normals(10)

    def step_resident():
        return ctx.register_symm(MAX_ITERS, DIFF_THR)

    def step_e2e():
        setup_resident()
        return ctx.register_symm(MAX_ITERS, DIFF_THR)

    # ---------------- value: loop only, resident inputs, graph mode ----------------
    setup_resident()
    for _ in range(args.warmup):
        l2_flush()
        T, st, _ = step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    loop_ms, iters, launches = [], [], 0
    for _ in range(args.steps):
        l2_flush()
        T, st, _ = step_resident()           # stats.loop_ms = CUDA events on the launching stream
        loop_ms.append(st["loop_ms"])
        iters.append(st["iters"])
        launches += 2 + 2 * st["searches"]   # set_lp + init_state, then [search, solve] per WHILE-body trip
    barrier()
    my_ms = float(np.sum(loop_ms))
    my_corr = float(n) * float(np.sum(iters))

    # ---------------- roofline: same steps with per-kernel CUDA events (stream mode) ----------------
    ctx.set_option(icp.OPT_KERNEL_TIMING, 1)
    per_search, per_solve = [], []
    for _ in range(args.steps):
        l2_flush()
        _, stk, _ = step_resident()
        a, b = ctx.search_times()
        per_search.append(a); per_solve.append(b)
        stream_loop_ms = stk["loop_ms"]
    ctx.set_option(icp.OPT_KERNEL_TIMING, 0)
    per_search = np.mean(np.stack(per_search), axis=0)          # mean duration of search k over the steps
    per_solve = np.mean(np.stack(per_solve), axis=0)
    # set-up kernels on their own (synchronous set-up, CUDA events on the context stream)
    ctx.set_option(icp.OPT_ASYNC_SETUP, 0)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    build_ms, normals_ms = [], []
    for _ in range(max(3, args.steps // 2)):
        l2_flush()
        evs[0].record(stream)
        ctx.set_source_ptr(h_src.data_ptr(), n, 12); ctx.set_target_ptr(h_tgt.data_ptr(), n, 12)
        evs[1].record(stream)
        ctx.estimate_normals(10)
        evs[2].record(stream)
        evs[2].synchronize()
        build_ms.append(evs[0].elapsed_time(evs[1])); normals_ms.append(evs[1].elapsed_time(evs[2]))
    ctx.set_option(icp.OPT_ASYNC_SETUP, 1)

    # ---------------- e2e: host buffers in, transform out, every step ----------------
    for _ in range(2):
        step_e2e()
    barrier()
    e2e_ms, e2e_iters = [], []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.steps):
        l2_flush()
        ev0.record(stream)
        T, st, _ = step_e2e()
        ev1.record(stream)
        ev1.synchronize()
        e2e_ms.append(ev0.elapsed_time(ev1))
        e2e_iters.append(st["iters"])
    barrier()
    # the same end-to-end step as a stream of registrations over two contexts (icpsymm_register_symm_async /
    # icpsymm_wait): pair k+1 is uploaded, indexed and given normals while pair k iterates.  Reported beside `e2e`
    # (which stays the strictly serial call), not instead of it.
    pipe_ms, pipe_iters = None, 0
    try:
        pool = [ctx, icp.Context(local_rank)]
        pool[1].set_option(icp.OPT_REL_DIFF_TOL, REL_TOL)
        pool[1].set_option(icp.OPT_ASYNC_SETUP, 1)

        def pipelined(k_steps):
            its = 0
            for k in range(k_steps):
                c = pool[k % 2]
                if k >= 2:
                    its += c.wait()[1]["iters"]
                c.set_source_ptr(h_src.data_ptr(), n, 12)
                c.set_target_ptr(h_tgt.data_ptr(), n, 12)
                c.estimate_normals(10)
                c.register_symm_async(MAX_ITERS, DIFF_THR)
            for k in range(max(0, k_steps - 2), k_steps):
                its += pool[k % 2].wait()[1]["iters"]
            return its

        pipelined(4)
        barrier()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_pipe = max(4, args.steps)
        pe0.record()
        pipe_iters = pipelined(k_pipe)
        pe1.record()
        pe1.synchronize()
        pipe_ms = pe0.elapsed_time(pe1) / k_pipe
        barrier()
        pool[1].close()
    except Exception as e:   # an extra: never takes the line with it
        pipe_ms = None
        sys.stderr.write(f"pipelined e2e skipped: {type(e).__name__}: {e}\n")
    clocks = sampler.stop()     # sampled across the value, roofline and e2e timed regions
    my_e2e_ms = float(np.sum(e2e_ms))
    my_e2e_corr = float(n) * float(np.sum(e2e_iters))

    # accuracy of what was timed (not part of the timed region)
    rot_err = wl.rotation_angle(T, T_true)
    trans_err = float(np.abs(T[:3, 3] - T_true[:3, 3]).max() / wl.extent(tgt))

    # ---------------- reduce over ranks: sum of work, max of time; per-rank detail ----------------
    vec = torch.tensor([my_corr, my_e2e_corr, my_ms, my_e2e_ms], dtype=torch.float64, device="cuda")
    mine = torch.tensor([my_ms / args.steps, my_e2e_ms / args.steps, float(iters[-1]), float(st["searches"])],
                        dtype=torch.float64, device="cuda")
    per_rank = [mine]
    if world > 1:
        work = vec[:2].clone(); tmax = vec[2:].clone()
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        vec = torch.cat([work, tmax])
        per_rank = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(per_rank, mine)
    tot_corr, tot_e2e_corr, max_ms, max_e2e_ms = [float(x) for x in vec.tolist()]

    line = None
    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        prof = ncu_profile()
        k_launch = int(len(per_search))
        k_ms = float(per_search.sum())
        achieved = (BYTES_PER_CORR_FUSED * float(n) * k_launch / (k_ms * 1e-3)) / 1e9 if k_ms > 0 else 0.0
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        slots = 148 * 4 * sm_mhz * 1e6       # warp-instruction issue slots per second: 148 SMs x 4 schedulers x clock
        issue = {"issue_slots_per_s": slots, "how": "warp instructions per launch from the committed ncu capture "
                 "(smsp__inst_executed.sum) / this run's CUDA-event duration of the same launch / (148 SMs x 4 schedulers x SM clock)"}
        winst = prof.get("warp_inst_per_launch")
        if winst and n == prof.get("points", 1_000_000):
            for name, k in (("far", 0), ("converged", k_launch - 1)):
                wi = float(winst[min(k, len(winst) - 1)])
                issue[name] = {"search": k, "ms": float(per_search[k]), "warp_inst_per_launch": wi, "inst_per_query": wi * 32.0 / n,
                               "frac": wi / (float(per_search[k]) * 1e-3) / slots}
        line = {
            "metric": METRIC, "value": tot_corr / (max_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n), "points": n, "max_iters": MAX_ITERS, "diff_threshold": DIFF_THR,
                       "rel_diff_tol": REL_TOL, "iters_run": int(iters[-1]), "searches_run": int(st["searches"]),
                       "pairs": "one pair per GPU" + (" (seed 1234 + rank)" if args.seed_per_rank else " (the same pair on every GPU)"),
                       "l2": "flushed between steps (256 MB memset)",
                       "mode": "one CUDA graph per step: WHILE node with body [search+accumulate, reduce+solve], condition set on device",
                       "rot_err_rad": rot_err, "trans_err_over_extent": trans_err},
            "clocks": clocks,
            "e2e": {"value": tot_e2e_corr / (max_e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": max_e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(2 * n * 12), "d2h_bytes_per_step": 64 + 2 * 128,
                    "includes": "H2D of both clouds, Hilbert sort + k-d refinement + box hierarchy, kNN-10 normals, loop, D2H transform",
                    "pipelined": None if pipe_ms is None else {
                        "ms_per_step": pipe_ms, "value_per_gpu": float(n) * pipe_iters / max(4, args.steps) / (pipe_ms * 1e-3),
                        "how": "the same step as a stream of registrations over two contexts with icpsymm_register_symm_async / "
                               "icpsymm_wait: upload + index + normals of pair k+1 overlap the loop of pair k (this rank; device clock)"}},
            "gpu_launches": int(launches),
            "per_rank": [{"rank": r, "loop_ms": float(t[0]), "e2e_ms": float(t[1]), "iters": int(t[2]), "searches": int(t[3])}
                         for r, t in enumerate(per_rank)],
            "roofline": {"bound": "hbm", "kernel": "nn_accum_kernel (fused NN search + symmetric accumulate)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "peak_kind": f"of {peak_kind}", "traffic": prof.get("dram_bytes_per_launch_mean"),
                         "traffic_source": prof.get("file"),
                         "bytes_per_unit": BYTES_PER_CORR_FUSED, "units_per_launch": n,
                         "avg_launch_ms": k_ms / k_launch if k_launch else None, "launches_timed": k_launch * args.steps,
                         "per_search_ms": [float(x) for x in per_search],
                         "per_search_frac": [BYTES_PER_CORR_FUSED * float(n) / (float(x) * 1e-3) / 1e9 / peak for x in per_search],
                         "stream_mode_loop_ms": stream_loop_ms,
                         "issue": issue,
                         "kernels": [
                             {"kernel": "solve_kernel (reduce + 6x6 solve + compose)", "ms": float(per_solve.mean()),
                              "bytes": BYTES_PER_ROW_K2 * icp.tile_rows(n),
                              "achieved": BYTES_PER_ROW_K2 * icp.tile_rows(n) / (float(per_solve.mean()) * 1e-3) / 1e9,
                              "frac": BYTES_PER_ROW_K2 * icp.tile_rows(n) / (float(per_solve.mean()) * 1e-3) / 1e9 / peak,
                              "note": "latency-bound by design: includes the launch gap; one 320-byte row per 256-query tile"},
                             {"kernel": "normals_kernel<10> x 2 clouds (kNN-10 + covariance + eigensolve)", "ms": float(np.min(normals_ms)),
                              "bytes": 2 * BYTES_PER_POINT_NORMALS * n,
                              "achieved": 2 * BYTES_PER_POINT_NORMALS * n / (float(np.min(normals_ms)) * 1e-3) / 1e9,
                              "frac": 2 * BYTES_PER_POINT_NORMALS * n / (float(np.min(normals_ms)) * 1e-3) / 1e9 / peak},
                             {"kernel": "H2D + index build x 2 clouds (bbox, Hilbert keys, radix sort, k-d refinement, boxes)",
                              "ms": float(np.min(build_ms))},
                         ]},
        }
    if not args.no_extras:
        try:
            sh = shard_sub(args, torch, dist, icp, rank, world, local_rank)
        except Exception as e:   # a failed sub-benchmark must not take the headline line with it
            sh = {"error": f"{type(e).__name__}: {e}"[:300]}
        try:
            ba = batch_sub(args, torch, dist, icp, rank, world, local_rank, barrier)
        except Exception as e:
            ba = {"error": f"{type(e).__name__}: {e}"[:300]}
        if rank == 0:
            line["shard16m"], line["batch32k"] = sh, ba
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, src, tgt, ctx)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
