"""How much of each search of the 1M registration is the price of not knowing the answer?

For every search k of the registration (transform T_k from the trace) the search is timed alone
  base    : as the loop runs it (k = 0: no seed; k > 0: seeded with the correspondences of search k-1)
  mode m  : seeded by icpsymm_debug_seed(m) from the TRUE neighbours of this very search
            (1 own neighbour = perfect bound, 2 unit centre's, 3 same lane of the previous unit, 4 previous tile,
             5 centre of the previous unit, 6 one probe per 256 points, 7 one per 1024, 8 one per leaf group)
Needs an ICP_EXPERIMENT build: python tools/experiments/seed_potential.py tools/_scratch/variants/lib_base.so
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import numpy as np
import torch
import icpsymm as icp
icp.LIB_PATH = os.path.abspath(sys.argv[1])
import workloads as wl

n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
src, tgt, T = wl.config_synth(n)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with icp.Context(0) as ctx:
    L = ctx.L
    L.icpsymm_debug_seed.argtypes = [C.c_void_p, C.c_int]
    L.icpsymm_debug_single_search.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(icp.Stats)]
    stream = torch.cuda.ExternalStream(ctx.stream)

    def single(Tk, reps=3):
        ms = []
        for _ in range(reps):
            with torch.cuda.stream(stream):
                flush.zero_()
            st = icp.Stats()
            t16 = np.ascontiguousarray(Tk.reshape(16), np.float32)
            rc = L.icpsymm_debug_single_search(ctx.h, icp._fp(t16), C.byref(st))
            assert rc >= 0, rc
            ms.append(st.loop_ms)
        return min(ms)

    ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
    ctx.set_option(icp.OPT_TRACE, 1)
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    Tm, st, rc = ctx.register_symm(50, 1.0)
    Ts = [ctx.trace(k)[0] for k in range(st["searches"])]
    print("searches", st["searches"], "loop_ms", st["loop_ms"])
    ctx.set_option(icp.OPT_TRACE, 0)
    modes = [1, 2, 3, 4, 5, 6, 7, 8]
    print("k   base   " + "  ".join(f"m{m:<5d}" for m in modes))
    for k, Tk in enumerate(Ts):
        # baseline: seeds as the loop has them
        ctx.set_option(icp.OPT_WARM_START, 1)
        if k > 0:
            ctx.find_correspondences(Ts[k - 1])
            ctx.set_option(icp.OPT_WARM_START, 2)
        base = []
        for _ in range(3):
            if k > 0:
                ctx.set_option(icp.OPT_WARM_START, 1)
                ctx.find_correspondences(Ts[k - 1])
                ctx.set_option(icp.OPT_WARM_START, 2)
            base.append(single(Tk, 1))
        ctx.set_option(icp.OPT_WARM_START, 1)
        idx_true, _ = ctx.find_correspondences(Tk)
        L.icpsymm_debug_seed(ctx.h, -1)
        ctx.set_option(icp.OPT_WARM_START, 2)
        row = []
        for m in modes:
            t = []
            for _ in range(3):
                L.icpsymm_debug_seed(ctx.h, m)
                t.append(single(Tk, 1))
            row.append(min(t))
            idx, _ = ctx.get_correspondences()
            assert np.array_equal(idx, idx_true), f"seeded search changed the result (mode {m})"
        print(f"{k}  {min(base):6.3f}  " + "  ".join(f"{x:6.3f}" for x in row), flush=True)
