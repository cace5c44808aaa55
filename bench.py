#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 symmetric-ICP path (contract: see README / DESIGN.md).

Metric (BASELINE.json): correspondences/s of the registration loop (NN search + symmetric
accumulate + solve) on the synthetic 1M-point pair (config 3: bumpy ellipsoid, 30 deg /
0.1-extent perturbation, 1e-4-extent noise).  One "step" = one complete registration of that pair
from the initial perturbation until the loop stops (max 50 iterations, threshold 1.0, relative
diff tolerance 1e-5).  Counted work = N x (solves applied); the final convergence-check search is
timed but not counted.

  value : loop only, clouds + index + normals already resident in HBM (CUDA-graph mode)
  e2e   : the full MyICP::RegisterSymm-equivalent through the C ABI with HOST buffers --
          H2D of both clouds, index build, normal estimation, loop, D2H of the transform
  N > 1 : every rank registers its own independent pair (pair-batch data parallel, no collective);
          value = all ranks' correspondences / max-over-ranks time.  "scaling": "weak".

`--impl reference` times the CPU oracle port (the reference itself cannot be built here: no
PCL/Eigen) on a bounded sample of the same workload with all host threads.
Other workloads for manual runs: --workload batch32k (config 5), --workload shard16m (config 4).
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
METRIC, UNIT = "correspondences/sec (NN+symm solve)", "correspondences/s"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/r01_nn_accum.json); None when no capture is committed"""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_nn_accum.json")) as f:
            return json.load(f)["dram_bytes_per_launch_mean"]
    except Exception:
        return None


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


def make_pair(n, rank):
    src, tgt, T = wl.config_synth(n, seed=1234 + rank)
    return src, tgt, T


def run_reference(args, rank, world):
    """CPU arm: the oracle port (kd-tree leaf 15 + canonical accumulate/solve) on the host cores."""
    if rank != 0:
        return
    from oracle import oracle as orc
    n = args.points
    src, tgt, _ = make_pair(n, 0)
    sample = min(n, args.cpu_sample or n)
    sel = np.sort(np.random.default_rng(7).choice(n, sample, replace=False)) if sample < n else np.arange(n)
    threads = host_threads()
    times, iters = [], 0
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        # the whole MyICP::RegisterSymm of the reference (myicp.cpp:100-150): normals of both clouds
        # (kd-tree + kNN-10 + PCA), kd-tree of the target, then the loop -- what our e2e arm times
        ns, _ = orc.estimate_normals(src[sel], 10, threads=threads)
        nt, _ = orc.estimate_normals(tgt, 10, threads=threads)
        r = orc.register_symm(src[sel], ns, tgt, nt, max_iters=MAX_ITERS, diff_threshold=DIFF_THR * sample / n,
                              rel_diff_tol=REL_TOL, threads=threads)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
            iters = r["iters"]
    total = float(np.sum(times))
    value = sample * iters * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synth-{n} bumpy ellipsoid, 30deg/0.1-extent, noise 1e-4 (config 3)",
                       "max_iters": MAX_ITERS, "diff_threshold": DIFF_THR, "rel_diff_tol": REL_TOL, "iters_run": iters},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} of {n} source points against the full {n}-point target; every step is the "
                                       "whole RegisterSymm: k=10 normals of both clouds, kd-tree build, loop"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "PCL-equivalent restatement (exact kd-tree, leaf 15, OpenMP); PCL itself is not installable here"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synth1m", choices=["synth1m", "batch32k", "shard16m"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--pairs", type=int, default=4096)
    ap.add_argument("--cpu-sample", type=int, default=0, help="source points per CPU step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"], help="shard16m: exchange of the normal equations")
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

    if args.workload != "synth1m":
        import bench_extra
        bench_extra.run(args, rank, world, local_rank, barrier)
        return

    n = args.points
    src, tgt, T_true = make_pair(n, rank)
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
    k_ms, k_launch, k_units, s_ms = 0.0, 0, 0.0, 0.0
    for _ in range(args.steps):
        l2_flush()
        _, stk, _ = step_resident()
        k_ms += stk["search_ms"]; k_launch += stk["search_launches"]; k_units += float(n) * stk["search_launches"]
        s_ms += stk["solve_ms"]
        stream_loop_ms = stk["loop_ms"]
    ctx.set_option(icp.OPT_KERNEL_TIMING, 0)

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
    clocks = sampler.stop()     # sampled across the value, roofline and e2e timed regions
    my_e2e_ms = float(np.sum(e2e_ms))
    my_e2e_corr = float(n) * float(np.sum(e2e_iters))

    # accuracy of what was timed (not part of the timed region)
    rot_err = wl.rotation_angle(T, T_true)
    trans_err = float(np.abs(T[:3, 3] - T_true[:3, 3]).max() / wl.extent(tgt))

    # ---------------- reduce over ranks: sum of work, max of time ----------------
    vec = torch.tensor([my_corr, my_e2e_corr, my_ms, my_e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        work = vec[:2].clone(); tmax = vec[2:].clone()
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        vec = torch.cat([work, tmax])
    tot_corr, tot_e2e_corr, max_ms, max_e2e_ms = [float(x) for x in vec.tolist()]

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        achieved = (BYTES_PER_CORR_FUSED * k_units / (k_ms * 1e-3)) / 1e9 if k_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": tot_corr / (max_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synth-{n} bumpy ellipsoid, 30deg/0.1-extent, noise 1e-4 (config 3), one pair per GPU",
                       "points": n, "max_iters": MAX_ITERS, "diff_threshold": DIFF_THR, "rel_diff_tol": REL_TOL,
                       "iters_run": int(iters[-1]), "searches_run": int(st["searches"]),
                       "l2": "flushed between steps (256 MB memset)",
                       "mode": "one CUDA graph per step: WHILE node with body [search+accumulate, reduce+solve], condition set on device",
                       "rot_err_rad": rot_err, "trans_err_over_extent": trans_err},
            "clocks": clocks,
            "e2e": {"value": tot_e2e_corr / (max_e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": max_e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(2 * n * 12), "d2h_bytes_per_step": 64 + 2 * 128,
                    "includes": "H2D of both clouds, Hilbert sort + k-d refinement + box hierarchy, kNN-10 normals, loop, D2H transform"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "nn_accum_kernel (fused NN search + symmetric accumulate)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "peak_kind": f"of {peak_kind}", "traffic": ncu_traffic(),
                         "bytes_per_unit": BYTES_PER_CORR_FUSED, "units_per_launch": n,
                         "avg_launch_ms": k_ms / k_launch if k_launch else None, "launches_timed": k_launch,
                         "solve_kernel_avg_ms": s_ms / k_launch if k_launch else None,
                         "stream_mode_loop_ms": stream_loop_ms},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, src, tgt, ctx)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, src, tgt, ctx):
    """oracle port timed on the host cores, bounded sample of the same workload (normals from the device run)"""
    from oracle import oracle as orc
    n = len(src)
    sample = min(n, args.cpu_sample or n)
    sel = np.sort(np.random.default_rng(7).choice(n, sample, replace=False))
    ns, nt = ctx.get_normals(0)[sel], ctx.get_normals(1)
    threads = host_threads()
    t0 = time.perf_counter()
    r = orc.register_symm(src[sel], ns, tgt, nt, max_iters=MAX_ITERS, diff_threshold=DIFF_THR * sample / n,
                          rel_diff_tol=REL_TOL, threads=threads)
    dt = time.perf_counter() - t0
    return {"value": sample * r["iters"] / dt, "unit": UNIT, "cores": threads, "kind": "port", "seconds": dt,
            "iters": r["iters"],
            "sample": f"{sample} of {n} source points (random subset) against the full {n}-point target; whole "
                      "registration loop incl. kd-tree build (leaf 15, OpenMP over queries); normals given"}


if __name__ == "__main__":
    main()
