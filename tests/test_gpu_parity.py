"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): nearest-neighbour indices bit-exact (ties -> lowest index),
fp32 squared distances bit-exact, final transform rotation <= 1e-5 rad, translation <= 1e-5 x
cloud extent, RMSE relative <= 1e-5.
"""
import zlib

import numpy as np
import pytest

import icpsymm as icp
import workloads as wl
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-5          # rad
TRANS_TOL = 1e-5        # x extent
RMSE_RTOL = 1e-5


@pytest.fixture(scope="module")
def ctx():
    c = icp.Context(0)
    yield c
    c.close()


def _rand_T(rng, scale=1.0, max_deg=40.0):
    R = wl.rotation(rng.standard_normal(3), rng.uniform(-max_deg, max_deg))
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = rng.standard_normal(3) * scale
    return T.astype(np.float32)


def _check_nn(ctx, src, tgt, T=None):
    """device search under T == oracle search of the oracle-transformed points, bit for bit"""
    ctx.set_source(src)
    ctx.set_target(tgt)
    idx, d2 = ctx.find_correspondences(T)
    Tf = np.eye(4, dtype=np.float32) if T is None else np.asarray(T, np.float32)
    p = orc.transform_points(Tf, src)
    oi, od = (orc.nn_brute(p, tgt) if len(tgt) * len(src) <= 4e7 else orc.KdTree(tgt).nn(p))
    assert np.array_equal(idx, oi), f"{(idx != oi).sum()} of {len(idx)} indices differ"
    assert np.array_equal(d2, od)
    return idx, d2


def test_device_is_blackwell():
    info = icp.device_info(0)
    assert info["cc"][0] >= 10, info


def test_nn_cat_golden(ctx, cat, cat_out, gold):
    idx, d2 = _check_nn(ctx, cat, cat_out)
    assert np.array_equal(idx, gold["nn_idx"]) and np.array_equal(d2, gold["nn_d2"])
    assert np.array_equal(idx, gold["flann_idx"])        # OpenCV FLANN KDTREE_SINGLE agrees too
    assert ctx.eval_diff() == pytest.approx(np.sqrt(gold["nn_d2"].astype(np.float64)).sum(), rel=1e-6)


def test_nn_under_transforms(ctx, cat, cat_out, bunny):
    rng = np.random.default_rng(11)
    for _ in range(4):
        _check_nn(ctx, cat, cat_out, _rand_T(rng, 20.0))
    tgt, _ = wl.config_bunny(bunny)
    _check_nn(ctx, bunny, tgt)
    _check_nn(ctx, bunny, tgt, _rand_T(rng, 0.1, 10))


def test_nn_ties_lowest_index(ctx):
    g = np.stack(np.meshgrid(*[np.arange(9)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    tgt = np.concatenate([g, g[::-1], g[:100]])            # every point duplicated: ties everywhere
    q = np.concatenate([g + 0.5, g, g * 0.5 + 0.25]).astype(np.float32)
    idx, _ = _check_nn(ctx, q, tgt)
    assert (idx < len(g)).all()                           # the first copy always wins


@pytest.mark.parametrize("n,m", [(1, 1), (1, 9), (7, 3), (8, 8), (9, 65), (63, 511), (64, 513), (257, 4097),
                                 (1000, 37), (5000, 20000)])
def test_nn_ragged_sizes(ctx, n, m):
    rng = np.random.default_rng(n * 100003 + m)
    _check_nn(ctx, rng.standard_normal((n, 3)).astype(np.float32), rng.standard_normal((m, 3)).astype(np.float32))


def test_nn_far_queries_and_clusters(ctx):
    rng = np.random.default_rng(5)
    tgt = np.concatenate([rng.standard_normal((30000, 3)) * 0.01 + c for c in ([0, 0, 0], [5, 5, 5], [-3, 8, 1])]).astype(np.float32)
    q = (rng.standard_normal((20000, 3)) * 6).astype(np.float32)
    _check_nn(ctx, q, tgt)
    # pcl::PointXYZ-style 16-byte stride (myicp.h:5)
    q4 = np.zeros((len(q), 4), np.float32); q4[:, :3] = q
    ctx.set_source(q4[:, :3])                              # strided view, stride 16
    idx, _ = ctx.find_correspondences()
    assert np.array_equal(idx, orc.KdTree(tgt).nn(q)[0])


def test_nn_synth_200k(ctx):
    src, tgt, T = wl.config_synth(200_000)
    _check_nn(ctx, src, tgt)                               # far queries (30 deg / 0.1 extent)
    _check_nn(ctx, src, tgt, T.astype(np.float32))         # near queries (aligned)


def _angle(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.arctan2(np.linalg.norm(np.cross(a, b), axis=1), np.abs(np.einsum("ni,ni->n", a, b)))


def test_normals_parity(ctx, cat, cat_out):
    src, tgt, _ = wl.config_synth(50_000)
    for a, b in ((cat, cat_out), (src, tgt)):
        ctx.set_source(a)
        ctx.set_target(b)
        ctx.estimate_normals(10)
        for which, xyz in ((0, a), (1, b)):
            dn = ctx.get_normals(which)
            on, gap = orc.estimate_normals(xyz, 10)
            ok = gap > 1e-3
            ang = _angle(dn[ok, :3], on[ok, :3])
            assert ang.max() < 1e-4, f"max normal angle {ang.max():.3e} rad"
            assert (np.einsum("ni,ni->n", dn[ok, :3], on[ok, :3]) > 0).mean() > 0.9999
            assert np.allclose(dn[ok, 3], on[ok, 3], atol=2e-5)
            assert np.allclose(np.linalg.norm(dn[:, :3], axis=1), 1.0, atol=1e-5)


def test_normals_k_variants_and_small_clouds(ctx):
    rng = np.random.default_rng(3)
    P = np.c_[rng.uniform(-1, 1, (3000, 2)), np.zeros(3000)].astype(np.float32)
    P[:, 2] = 0.05 * np.sin(3 * P[:, 0])
    ctx.set_source(P)
    ctx.set_target(P[:5])                                  # fewer points than k
    for k in (5, 10, 16, 24):
        ctx.estimate_normals(k, viewpoint=(0, 0, 10))
        dn = ctx.get_normals(0)
        on, gap = orc.estimate_normals(P, k, viewpoint=(0, 0, 10))
        ok = gap > 1e-3
        assert _angle(dn[ok, :3], on[ok, :3]).max() < 1e-4
        assert (dn[:, 2] > 0).all()                        # flipped towards the viewpoint above the sheet
        tn = ctx.get_normals(1)
        assert np.isfinite(tn).all()


def test_normal_equations_match_oracle_and_literal_rows(ctx, cat, cat_out):
    ctx.set_source(cat)
    ctx.set_target(cat_out)
    ctx.estimate_normals(10)
    ns, nt = ctx.get_normals(0), ctx.get_normals(1)
    rng = np.random.default_rng(2)
    for T, origin in ((None, (0.0, 0.0, 0.0)), (_rand_T(rng, 5.0, 10), (3.0, 20.0, 30.0))):
        S = ctx.normal_equations(T, origin)
        Tf = np.eye(4, dtype=np.float32) if T is None else T
        p, npn = orc.transform_points(Tf, cat), orc.rotate_vectors(Tf, ns)
        idx, _ = orc.nn_brute(p, cat_out)
        So = orc.symm_accumulate(p, npn, cat_out[idx], nt[idx], origin)
        scale = np.abs(So).max()
        assert np.allclose(S, So, rtol=2e-5, atol=1e-6 * scale), np.abs(S - So).max() / scale
        assert S[36] == len(cat)
    # identity pairing + origin 0 is exactly the sum of calculateMatrixNotation's rows (func.cpp:43-60)
    ctx.set_option(icp.OPT_IDENTITY_PAIRS, 1)
    try:
        S = ctx.normal_equations(None, (0, 0, 0))
    finally:
        ctx.set_option(icp.OPT_IDENTITY_PAIRS, 0)
    M, N, c = orc.matrix_notation(cat, ns[:, :3], cat_out, nt[:, :3])
    M, N, c = M.astype(np.float64), N.astype(np.float64), c.astype(np.float64)
    iu = np.triu_indices(3)
    assert np.allclose(S[0:6], (M.T @ M)[iu], rtol=2e-5)
    assert np.allclose(S[15:21], (N.T @ N)[iu], rtol=2e-5)
    assert np.allclose(S[21:24], M.T @ c, rtol=2e-5) and np.allclose(S[24:27], N.T @ c, rtol=2e-5)
    assert S[33] == pytest.approx(orc.eval_diff(cat, cat_out), rel=1e-6)    # evalDiff, func.cpp:19-32


def test_solve_compose_device_vs_oracle(ctx, cat, cat_out, gold):
    ns, nt = gold["normals_cat"], gold["normals_cat_out"]
    for origin in ((0.0, 0.0, 0.0), (5.0, 20.0, 30.0)):
        S = orc.symm_accumulate(cat, ns, cat_out, nt, origin)
        To, xo, dego = orc.solve_compose(S, origin)
        Td, xd, rc, ratio = ctx.solve_compose(S, origin)
        assert rc == dego == 0 and np.allclose(xd, xo, rtol=1e-10, atol=1e-12) and np.allclose(Td, To, atol=1e-10)
        assert ratio == pytest.approx(orc.solve_compose.last_min_pivot_ratio, rel=1e-9)
    # rank-deficient: a plane
    rng = np.random.default_rng(1)
    P = np.c_[rng.uniform(-1, 1, (500, 2)), np.zeros(500)].astype(np.float32)
    n4 = np.tile(np.array([0, 0, 1, 0], np.float32), (500, 1))
    S = orc.symm_accumulate(P, n4, P + np.float32([0, 0, 0.01]), n4)
    Td, xd, rc, _ = ctx.solve_compose(S)
    To, xo, dego = orc.solve_compose(S)
    assert rc == icp.WARN_DEGENERATE and dego == 1 and np.isfinite(Td).all() and np.allclose(Td, To, atol=1e-12)


def test_solve_compose_matches_independent_lstsq_golden(ctx, cat, cat_out, gold, solve_gold):
    """device normal equations (fp32 products, fp64 sums) + device solve against the independent numpy lstsq golden
    (tests/golden/make_golden.py: canonical_solve), identity pairs and exact-NN pairs, golden PCA normals"""
    ctx.set_source(cat); ctx.set_target(cat_out)
    ctx.set_normals(0, gold["normals_cat"]); ctx.set_normals(1, gold["normals_cat_out"])
    ext = wl.extent(cat)
    for name, ident in (("nn", 0), ("identity", 1)):
        ctx.set_option(icp.OPT_IDENTITY_PAIRS, ident)
        try:
            for origin in ((0.0, 0.0, 0.0), (10.0, 20.0, 30.0)):
                S = ctx.normal_equations(None, origin)
                T, x, rc, _ = ctx.solve_compose(S, origin)
                assert rc == 0 and S[36] == len(cat)
                assert np.allclose(x, solve_gold["x_" + name], rtol=1e-5, atol=1e-6)
                assert wl.rotation_angle(T, solve_gold["T_" + name]) <= 1e-6
                assert np.abs(T[:3, 3] - solve_gold["T_" + name][:3, 3]).max() <= 1e-6 * ext
        finally:
            ctx.set_option(icp.OPT_IDENTITY_PAIRS, 0)
    # the device solve alone, fed with exact fp64 sums
    idx = gold["nn_idx"]
    S = orc.symm_accumulate(cat, gold["normals_cat"], cat_out[idx], gold["normals_cat_out"][idx], (0, 0, 0))
    T, x, rc, _ = ctx.solve_compose(S, (0, 0, 0))
    assert np.allclose(x, solve_gold["x_nn"], rtol=0, atol=1e-7) and wl.rotation_angle(T, solve_gold["T_nn"]) <= 1e-8


def _register_and_compare(ctx, src, tgt, max_iters, thr, rel_tol=0.0, T_true=None, check_trace=True, match_iters=False):
    ctx.set_source(src)
    ctx.set_target(tgt)
    ctx.estimate_normals(10)
    ns, nt = ctx.get_normals(0), ctx.get_normals(1)
    ctx.set_option(icp.OPT_TRACE, 1)
    ctx.set_option(icp.OPT_REL_DIFF_TOL, rel_tol)
    try:
        T, st, rc = ctx.register_symm(max_iters, thr)
        if match_iters:   # a relative-change stop can fall either side of a rounding error: pin the trip count
            r = orc.register_symm(src, ns, tgt, nt, max_iters=st["iters"], diff_threshold=-1.0)
            assert 0 < st["iters"] < max_iters
        else:
            r = orc.register_symm(src, ns, tgt, nt, max_iters=max_iters, diff_threshold=thr, rel_diff_tol=rel_tol)
            assert st["converged"] == r["converged"], (st, r["converged"])
        assert st["iters"] == r["iters"], (st, r["iters"])
        ext = wl.extent(tgt)
        assert wl.rotation_angle(T, r["T"]) <= ROT_TOL
        assert np.abs(T[:3, 3].astype(np.float64) - r["T"][:3, 3]).max() <= TRANS_TOL * ext
        assert st["rmse"] == pytest.approx(r["rmse"], rel=RMSE_RTOL, abs=1e-7 * ext)
        assert st["diff"] == pytest.approx(r["diff"], rel=1e-4, abs=1e-6 * ext * len(src))
        # correspondences of the final search: bit-exact against the oracle fed with the device transform
        idx, d2 = ctx.get_correspondences()
        p = orc.transform_points(T, src)
        oi, od = orc.KdTree(tgt).nn(p)
        assert np.array_equal(idx, oi) and np.array_equal(d2, od)
        if check_trace:
            # EVERY search of the loop, not only the last: the transform the loop held at search k, the device
            # search under it (indices and fp32 distances bit-exact against the oracle) and the diff the loop saw
            tree = orc.KdTree(tgt)
            traced = [ctx.trace(k) for k in range(st["searches"])]
            for k, (Tk, dk) in enumerate(traced):
                oi, od = tree.nn(orc.transform_points(Tk, src))
                assert dk == pytest.approx(np.sqrt(od.astype(np.float64)).sum(), rel=1e-6), k
                ik, d2k = ctx.find_correspondences(Tk)
                assert np.array_equal(ik, oi) and np.array_equal(d2k, od), k
                if k < len(r["trace_T"]):   # the oracle's own transform at search k
                    assert wl.rotation_angle(Tk, r["trace_T"][k]) <= ROT_TOL, k
        if T_true is not None:
            assert wl.rotation_angle(T, T_true) < 1e-4
    finally:
        ctx.set_option(icp.OPT_TRACE, 0)
        ctx.set_option(icp.OPT_REL_DIFF_TOL, 0.0)
    return T, st


def test_register_cat_matches_oracle_and_truth(ctx, cat, cat_out, gold):
    T, st = _register_and_compare(ctx, cat, cat_out, 50, 1.0)
    assert st["iters"] == 5 and st["converged"] == 1 and st["degenerate"] == 0
    # the reference's only known answer: Rz(45 deg), t = (2.5, 0, 0)  (main.cpp:44-52)
    assert wl.rotation_angle(T, gold["truth_R"]) < 1e-5
    assert np.abs(T[:3, 3] - gold["truth_t"]).max() < 1e-5 * wl.extent(cat)
    # reference defaults (myicp.cpp:6): 10 iterations, threshold 1.0
    T10, st10, _ = ctx.register_symm()
    assert st10["iters"] == 5 and np.array_equal(T10, T)


def test_register_stream_mode_equals_graph_mode(ctx, cat, cat_out):
    ctx.set_source(cat); ctx.set_target(cat_out); ctx.estimate_normals(10)
    Tg, sg, _ = ctx.register_symm(50, 1.0)
    ctx.set_option(icp.OPT_USE_GRAPH, 0)
    ctx.set_option(icp.OPT_KERNEL_TIMING, 1)
    try:
        Ts, ss, _ = ctx.register_symm(50, 1.0)
    finally:
        ctx.set_option(icp.OPT_USE_GRAPH, 1)
        ctx.set_option(icp.OPT_KERNEL_TIMING, 0)
    assert np.array_equal(Tg, Ts) and sg["iters"] == ss["iters"] and sg["diff"] == ss["diff"]
    assert ss["search_launches"] == ss["searches"] and ss["search_ms"] > 0
    # the unrolled graph ((max_iters+1) x [search, solve] with early exits) is a third way to run the same loop
    ctx.set_option(icp.OPT_USE_GRAPH, 2)
    try:
        Tu, su, _ = ctx.register_symm(50, 1.0)
    finally:
        ctx.set_option(icp.OPT_USE_GRAPH, 1)
    assert np.array_equal(Tg, Tu) and su["iters"] == sg["iters"] and su["diff"] == sg["diff"]
    # warm start off must not change any result
    ctx.set_option(icp.OPT_WARM_START, 0)
    try:
        Tc, sc, _ = ctx.register_symm(50, 1.0)
    finally:
        ctx.set_option(icp.OPT_WARM_START, 1)
    assert np.array_equal(Tg, Tc) and sc["diff"] == sg["diff"]


def test_scheduling_and_seeding_options_never_change_a_result(ctx):
    """tile order (most expensive first), probes (one seed per tile after a long step) and walk skipping only change HOW
    the exact neighbours are found: every combination gives the same transform bits, diff and correspondences"""
    src, tgt, _ = wl.config_synth(300_000)          # > 2 x 592 tiles: the ordered first search is active
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
    out = []
    try:
        for order, probe, skip in ((1, 1, 1), (0, 1, 1), (1, 0, 1), (1, 1, 0), (0, 0, 0)):
            ctx.set_option(icp.OPT_TILE_ORDER, order); ctx.set_option(icp.OPT_PROBE, probe); ctx.set_option(icp.OPT_SKIP_WALK, skip)
            T, st, _ = ctx.register_symm(50, 1.0)
            out.append((T.copy(), st["iters"], st["diff"], st["rmse"], ctx.get_correspondences()[0].copy()))
    finally:
        ctx.set_option(icp.OPT_TILE_ORDER, 1); ctx.set_option(icp.OPT_PROBE, 1); ctx.set_option(icp.OPT_SKIP_WALK, 1)
        ctx.set_option(icp.OPT_REL_DIFF_TOL, 0.0)
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0]) and o[1:4] == out[0][1:4] and np.array_equal(o[4], out[0][4])


@pytest.mark.parametrize("case", ["cat", "synth200k", "noise_free", "duplicates"])
def test_skipping_walks_never_changes_a_result(ctx, cat, cat_out, case):
    """ICPSYMM_OPT_SKIP_WALK: a query whose winner provably stands keeps it without walking the index.  The proof is a
    margin argument about fp32 values, so the result must be IDENTICAL to always walking: transform bits, diff, rmse,
    iteration count and every correspondence of the last search -- and the walks must actually be skipped."""
    if case == "cat":
        src, tgt, thr, rel, iters = cat, cat_out, 1e-3, 0.0, 12
    elif case == "synth200k":
        src, tgt, _ = wl.config_synth(200_000)
        thr, rel, iters = 1.0, 1e-7, 14          # runs on long after convergence: many skipped searches
    elif case == "noise_free":
        src, tgt, _ = wl.config_synth(60_000, noise_frac=0.0)
        thr, rel, iters = 1e-9, 0.0, 10
    else:
        # every 5th target point exists twice (exact ties, the later copy must never win) and a block of source points
        # coincides exactly with target points: zero distances, zero margins
        src, tgt, Tt = wl.config_synth(40_000, deg=8.0, shift_frac=0.02)
        tgt = np.concatenate([tgt, tgt[::5]])[np.random.default_rng(3).permutation(len(tgt) + len(tgt[::5]))]
        src = src.copy(); src[:500] = orc.transform_points(np.linalg.inv(Tt).astype(np.float32), tgt[:500])
        thr, rel, iters = 1e-9, 0.0, 9
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    ctx.set_option(icp.OPT_REL_DIFF_TOL, rel)
    ctx.set_option(icp.OPT_KERNEL_TIMING, 1)
    out = []
    try:
        for skip in (1, 0):
            ctx.set_option(icp.OPT_SKIP_WALK, skip)
            T, st, _ = ctx.register_symm(iters, thr)
            idx, d2 = ctx.get_correspondences()
            out.append((T.copy(), st, idx.copy(), d2.copy(), ctx.search_times()[0]))
    finally:
        ctx.set_option(icp.OPT_SKIP_WALK, 1)
        ctx.set_option(icp.OPT_REL_DIFF_TOL, 0.0)
        ctx.set_option(icp.OPT_KERNEL_TIMING, 0)
    (Ta, sa, ia, da, ta), (Tb, sb, ib, db, tb) = out
    assert np.array_equal(Ta, Tb) and np.array_equal(ia, ib) and np.array_equal(da, db)
    assert sa["iters"] == sb["iters"] and sa["diff"] == sb["diff"] and sa["rmse"] == sb["rmse"] and sa["pairs"] == sb["pairs"]
    oi, od = orc.KdTree(tgt).nn(orc.transform_points(Ta, src))
    assert np.array_equal(ia, oi) and np.array_equal(da, od)
    if case == "synth200k":
        assert sa["searches"] >= 8 and ta[-1] < 0.6 * tb[-1], (ta, tb)   # the late searches are mostly skipped


def test_register_identity_pairs_and_zero_trip(ctx, cat, cat_out, gold):
    ctx.set_source(cat); ctx.set_target(cat_out); ctx.estimate_normals(10)
    ctx.set_option(icp.OPT_IDENTITY_PAIRS, 1)
    try:
        T, st, _ = ctx.register_symm(10, 1.0)
        assert st["iters"] == 1 and wl.rotation_angle(T, gold["truth_R"]) < 1e-5
        # the documented distance to the LITERAL reference loop: same pairing
        # (myicp.cpp:128-131), but the reference solves two alternating 3-unknown systems on un-centred points and
        # composes in the reverse order (func.cpp:51-58,85-99).  Its first iterate is 0.107 rad from the fixture's
        # known answer, its tenth 2e-3 rad with a translation that never settles; ours is the known answer after one.
        assert wl.rotation_angle(T, gold["lit_T1"]) == pytest.approx(0.1066, abs=2e-3)
        assert wl.rotation_angle(T, gold["lit_T_final"]) < 3e-3
        assert np.abs(gold["lit_T_final"][:3, 3] - gold["truth_t"]).max() > 5.0 > 1e-3 > np.abs(T[:3, 3] - gold["truth_t"]).max()
        ctx.set_target(cat_out[:-1])
        with pytest.raises(icp.IcpError) as e:
            ctx.register_symm(10, 1.0)
        assert e.value.status == icp.ERR_SIZE                                        # func.cpp:21 assert
    finally:
        ctx.set_option(icp.OPT_IDENTITY_PAIRS, 0)
    # already aligned: zero trips, identity result (myicp.cpp:123)
    ctx.set_source(cat); ctx.set_target(cat); ctx.estimate_normals(10)
    T, st, _ = ctx.register_symm(10, 1.0)
    assert st["iters"] == 0 and st["converged"] == 1 and np.array_equal(T, np.eye(4, dtype=np.float32))


def test_register_bunny_config2(ctx, bunny):
    tgt, Ttrue = wl.config_bunny(bunny)
    ctx.set_source(bunny); ctx.set_target(tgt); ctx.estimate_normals(10)
    T, st, rc = ctx.register_symm(10, 1e-6)
    assert np.isfinite(T).all() and np.isfinite(st["diff"])
    idx, d2 = ctx.get_correspondences()
    oi, od = orc.nn_brute(orc.transform_points(T, bunny), tgt)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    assert st["rmse"] == pytest.approx(np.sqrt(od.astype(np.float64).mean()), rel=1e-5)
    # exactly collinear input with consistent normals: flagged, finite
    line = np.c_[np.linspace(0, 1, 200), np.zeros(200), np.zeros(200)].astype(np.float32)
    n4 = np.tile(np.array([0, 0, 1, 0], np.float32), (200, 1))
    ctx.set_source(line); ctx.set_target(line + np.float32([0, 0, 0.01]))
    ctx.set_normals(0, n4); ctx.set_normals(1, n4)
    T, st, rc = ctx.register_symm(3, 1e-9)
    assert rc == icp.WARN_DEGENERATE and st["degenerate"] == 1 and np.isfinite(T).all()


@pytest.mark.parametrize("case", ["cat", "synth50k"])
def test_end_to_end_oracle_with_its_own_normals(ctx, cat, cat_out, case):
    """whole RegisterSymm on both sides, nothing shared: oracle normals -> oracle loop against device normals -> device
    loop (every other registration test hands the device's normals to the oracle to isolate the loop)"""
    if case == "cat":
        src, tgt, thr, rel = cat, cat_out, 1.0, 0.0
    else:
        src, tgt, _ = wl.config_synth(50_000)
        thr, rel = 1.0, 1e-5
    ns, _ = orc.estimate_normals(src, 10)
    nt, _ = orc.estimate_normals(tgt, 10)
    ctx.set_source(src); ctx.set_target(tgt); ctx.estimate_normals(10)
    ctx.set_option(icp.OPT_REL_DIFF_TOL, rel)
    try:
        T, st, _ = ctx.register_symm(50, thr)
    finally:
        ctx.set_option(icp.OPT_REL_DIFF_TOL, 0.0)
    r = orc.register_symm(src, ns, tgt, nt, max_iters=st["iters"], diff_threshold=thr if case == "cat" else -1.0)
    assert r["iters"] == st["iters"]
    ext = wl.extent(tgt)
    assert wl.rotation_angle(T, r["T"]) <= ROT_TOL
    assert np.abs(T[:3, 3].astype(np.float64) - r["T"][:3, 3]).max() <= TRANS_TOL * ext
    assert st["rmse"] == pytest.approx(r["rmse"], rel=1e-4, abs=1e-7 * ext)


def test_register_synth_100k(ctx):
    src, tgt, Ttrue = wl.config_synth(100_000)
    _register_and_compare(ctx, src, tgt, 50, 1.0, rel_tol=1e-5, T_true=Ttrue, check_trace=True, match_iters=True)


def test_max_pair_distance_rejection(ctx, cat, cat_out):
    ctx.set_source(cat); ctx.set_target(cat_out); ctx.estimate_normals(10)
    ns, nt = ctx.get_normals(0), ctx.get_normals(1)
    ctx.set_option(icp.OPT_MAX_PAIR_DIST, 25.0)
    try:
        T, st, _ = ctx.register_symm(3, 1.0)
    finally:
        ctx.set_option(icp.OPT_MAX_PAIR_DIST, 0.0)
    r = orc.register_symm(cat, ns, cat_out, nt, max_iters=3, diff_threshold=1.0, max_pair_dist=25.0)
    assert st["iters"] == r["iters"] and st["pairs"] == r["pairs"]
    assert wl.rotation_angle(T, r["T"]) <= 1e-4 and np.allclose(T[:3, 3], r["T"][:3, 3], atol=1e-3)
    # normal-compatibility rejection (the other half of SURVEY 8f row 3)
    ctx.set_option(icp.OPT_MIN_NORMAL_DOT, 0.5)
    try:
        T, st, _ = ctx.register_symm(2, 1.0)
    finally:
        ctx.set_option(icp.OPT_MIN_NORMAL_DOT, -2.0)
    r = orc.register_symm(cat, ns, cat_out, nt, max_iters=2, diff_threshold=1.0, min_normal_dot=0.5)
    assert st["iters"] == r["iters"] and 0 < st["pairs"] < len(cat) and abs(int(st["pairs"]) - int(r["pairs"])) <= 2
    assert wl.rotation_angle(T, r["T"]) <= 1e-3


def test_batch_equals_single_pairs(ctx):
    pairs = [wl.config_batch_pair(b, 4096) for b in range(6)]
    counts = np.array([4096, 4000, 4096, 1234, 4096, 9], np.uint32)
    S = np.zeros((6, 4096, 3), np.float32); Tg = np.zeros((6, 4096, 3), np.float32)
    for b, (s, t, _) in enumerate(pairs):
        S[b, :counts[b]] = s[:counts[b]]; Tg[b, :counts[b]] = t[:counts[b]]
    ctx.set_source_batch(S, counts); ctx.set_target_batch(Tg, counts)
    ctx.estimate_normals(10)
    Tb, stb, _ = ctx.register_symm(30, 1e-3)
    idxb, _ = ctx.get_correspondences()
    idxb = idxb.reshape(6, 4096)
    for b in range(6):
        c = int(counts[b])
        ctx.set_source(S[b, :c]); ctx.set_target(Tg[b, :c]); ctx.estimate_normals(10)
        T1, st1, _ = ctx.register_symm(30, 1e-3)
        i1, _ = ctx.get_correspondences()
        assert np.array_equal(Tb[b], T1), b
        assert stb[b]["iters"] == st1["iters"] and stb[b]["diff"] == st1["diff"]
        assert np.array_equal(idxb[b, :c], i1)
        if c > 1000:
            assert wl.rotation_angle(T1, pairs[b][2]) < 5e-3


def test_query_shards_sum_to_whole(ctx, cat, cat_out):
    """the per-shard normal-equation vectors add up to the single-GPU vector"""
    ctx.set_source(cat); ctx.set_target(cat_out); ctx.estimate_normals(10)
    whole = ctx.normal_equations(None, (0, 0, 0))
    idx_whole, _ = ctx.find_correspondences()
    parts, idx = [], np.full(len(cat), -1, np.int32)
    for world in (2, 3):
        parts = []
        idx[:] = -1
        for r in range(world):
            ctx.set_query_shard(r, world)
            parts.append(ctx.normal_equations(None, (0, 0, 0)))
            i, _ = ctx.get_correspondences()
            assert ((idx >= 0) & (i >= 0)).sum() == 0       # shards are disjoint
            idx = np.where(i >= 0, i, idx)
        ctx.set_query_shard(0, 1)
        assert np.array_equal(idx, idx_whole)
        assert np.allclose(np.sum(parts, axis=0), whole, rtol=1e-5, atol=1e-6 * np.abs(whole).max())  # fp32 reassociation
    # canonical order (csrc/common.cuh): 8 ranks own one part each, and the whole is the 8 part sums added in
    # part order starting from +0.0 -- the very expression every rank evaluates after the exchange
    parts = []
    for r in range(8):
        ctx.set_query_shard(r, 8)
        parts.append(ctx.normal_equations(None, (0, 0, 0)))
    ctx.set_query_shard(0, 1)
    acc = np.zeros(icp.NSUMS)
    for prt in parts:
        acc = acc + prt
    assert np.array_equal(acc, whole)
    with pytest.raises(icp.IcpError):
        ctx.set_query_shard(0, 9)


def test_result_does_not_depend_on_the_batch_a_pair_sits_in(ctx):
    """the tile -> block map and the order of the sums follow from the pair's own point count: a pair gives the
    same bits alone, as the small member of a ragged batch, and next to a much larger pair"""
    pairs = [wl.config_batch_pair(b, n) for b, n in ((0, 20000), (1, 70000), (2, 3000))]
    res = []
    for s, t, _ in pairs:
        ctx.set_source(s); ctx.set_target(t); ctx.estimate_normals(10)
        T1, st1, _ = ctx.register_symm(30, 1e-3)
        res.append((T1.copy(), st1, ctx.get_correspondences()[0].copy()))
    m = max(len(p[0]) for p in pairs)
    S = np.zeros((3, m, 3), np.float32); Tg = np.zeros((3, m, 3), np.float32)
    counts = np.array([len(p[0]) for p in pairs], np.uint32)
    for b, (s, t, _) in enumerate(pairs):
        S[b, :len(s)], Tg[b, :len(t)] = s, t
    ctx.set_source_batch(S, counts); ctx.set_target_batch(Tg, counts); ctx.estimate_normals(10)
    Tb, stb, _ = ctx.register_symm(30, 1e-3)
    idxb = ctx.get_correspondences()[0].reshape(3, m)
    for b in range(3):
        assert np.array_equal(Tb[b], res[b][0]), b
        assert stb[b]["iters"] == res[b][1]["iters"] and stb[b]["diff"] == res[b][1]["diff"]
        assert np.array_equal(idxb[b, :counts[b]], res[b][2])
    # the same pairs twice in one batch of six: more than four segments take the 64-bit sort keys of the index
    # build (segment bits + 30 curve bits no longer fit 32), single pairs and small batches the 32-bit ones
    S6, T6, c6 = np.concatenate([S, S]), np.concatenate([Tg, Tg]), np.concatenate([counts, counts])
    ctx.set_source_batch(S6, c6); ctx.set_target_batch(T6, c6); ctx.estimate_normals(10)
    T6b, st6, _ = ctx.register_symm(30, 1e-3)
    idx6 = ctx.get_correspondences()[0].reshape(6, m)
    for b in range(6):
        assert np.array_equal(T6b[b], res[b % 3][0]), b
        assert st6[b]["iters"] == res[b % 3][1]["iters"] and st6[b]["diff"] == res[b % 3][1]["diff"]
        assert np.array_equal(idx6[b, :c6[b]], res[b % 3][2])


def test_register_npoint_regist_h(ctx, cat, cat_out, gold):
    R, T = ctx.register_npoint(cat, cat_out)
    rc, Ro, To = orc.registrate_npoint(cat, cat_out)
    assert np.allclose(R, Ro, atol=1e-10) and np.allclose(T, To, atol=1e-8)
    assert wl.rotation_angle(R, gold["truth_R"]) < 1e-7 and np.abs(T - gold["truth_t"]).max() < 1e-6
    with pytest.raises(icp.IcpError):
        ctx.register_npoint(cat[:2], cat_out[:2])
    rng = np.random.default_rng(0)
    P = rng.standard_normal((100000, 3)); Rt = wl.rotation((1, 2, 3), 77.0); t = np.array([1.0, -2.0, 3.0])
    R, T = ctx.register_npoint(P, P @ Rt.T + t)
    assert np.allclose(R, Rt, atol=1e-12) and np.allclose(T, t, atol=1e-12)


def test_transformed_source_is_apply_transform(ctx, cat):
    ctx.set_source(cat)
    T = _rand_T(np.random.default_rng(9), 3.0)
    assert np.array_equal(ctx.transformed_source(T), orc.transform_points(T, cat))


def test_errors_are_reported_not_hidden(ctx, cat):
    c = icp.Context(0)
    with pytest.raises(icp.IcpError) as e:
        c.register_symm()
    assert e.value.status == icp.ERR_NOT_READY
    c.set_source(cat); c.set_target(cat)
    with pytest.raises(icp.IcpError) as e:
        c.register_symm()                                   # normals missing
    assert e.value.status == icp.ERR_NOT_READY
    bad = cat.copy(); bad[7, 1] = np.nan
    with pytest.raises(icp.IcpError) as e:
        c.set_target(bad)
    assert e.value.status == icp.ERR_NONFINITE
    with pytest.raises(icp.IcpError):
        c.set_source(np.zeros((0, 3), np.float32))
    c.close()


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes: bit-exact index parity against the oracle kd-tree plus the
# size-independent properties of the domain
# ---------------------------------------------------------------------------------------------
def test_full_size_1m_config3(ctx):
    src, tgt, Ttrue = wl.config_synth(1_000_000)
    ctx.set_source(src); ctx.set_target(tgt)
    tree = orc.KdTree(tgt)
    # far first search (30 deg / 0.1 extent): every one of the 1M indices and fp32 distances identical
    idx, d2 = ctx.find_correspondences()
    oi, od = tree.nn(src)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    ctx.estimate_normals(10)
    ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
    try:
        T, st, _ = ctx.register_symm(50, 1.0)
        assert 0 < st["iters"] <= 50 and st["converged"] == 2 and st["degenerate"] == 0
        assert wl.rotation_angle(T, Ttrue) < 2e-5 and np.abs(T[:3, 3] - Ttrue[:3, 3]).max() < 2e-5 * wl.extent(tgt)
        # the north-star bars at the headline size, against the ORACLE run on the same inputs (trip count pinned: a
        # relative-change stop can fall either side of a rounding error)
        r = orc.register_symm(src, ctx.get_normals(0), tgt, ctx.get_normals(1), max_iters=st["iters"], diff_threshold=-1.0)
        assert r["iters"] == st["iters"]
        assert wl.rotation_angle(T, r["T"]) <= ROT_TOL
        assert np.abs(T[:3, 3].astype(np.float64) - r["T"][:3, 3]).max() <= TRANS_TOL * wl.extent(tgt)
        assert st["rmse"] == pytest.approx(r["rmse"], rel=RMSE_RTOL)
        assert np.array_equal(ctx.get_correspondences()[0], r["idx"]) or \
            (ctx.get_correspondences()[0] != r["idx"]).mean() < 1e-4     # pairs that sit on a tie between the two transforms
        # final correspondences: bit-exact against the oracle fed with the device transform
        idx, d2 = ctx.get_correspondences()
        oi, od = tree.nn(orc.transform_points(T, src))
        assert np.array_equal(idx, oi) and np.array_equal(d2, od)
        assert st["rmse"] == pytest.approx(np.sqrt(od.astype(np.float64).mean()), rel=1e-6)
        assert st["diff"] == pytest.approx(np.sqrt(od.astype(np.float64)).sum(), rel=1e-6)
        # determinism: a second run is bit-identical
        T2, st2, _ = ctx.register_symm(50, 1.0)
        assert np.array_equal(T, T2) and st2["diff"] == st["diff"]
        # idempotence: registering the aligned source changes (almost) nothing
        ctx.set_source(orc.transform_points(T, src)); ctx.estimate_normals(10)
        T3, st3, _ = ctx.register_symm(50, 1.0)
        assert wl.rotation_angle(T3, np.eye(4)) < 1e-5 and np.abs(T3[:3, 3]).max() < 1e-5 * wl.extent(tgt)
    finally:
        ctx.set_option(icp.OPT_REL_DIFF_TOL, 0.0)


def test_batch_of_32k_pairs_config5(ctx):
    B, n = 48, 32768
    pairs = [wl.config_batch_pair(b, n) for b in range(B)]
    S = np.stack([p[0] for p in pairs]); Tg = np.stack([p[1] for p in pairs])
    ctx.set_source_batch(S); ctx.set_target_batch(Tg)
    ctx.estimate_normals(10)
    ctx.set_option(icp.OPT_REL_DIFF_TOL, 1e-5)
    try:
        T, st, _ = ctx.register_symm(50, 1e-3)
    finally:
        ctx.set_option(icp.OPT_REL_DIFF_TOL, 0.0)
    idx, d2 = ctx.get_correspondences()
    idx = idx.reshape(B, n)
    for b in range(B):
        assert wl.rotation_angle(T[b], pairs[b][2]) < 1e-4, b
        assert np.abs(T[b][:3, 3] - pairs[b][2][:3, 3]).max() < 1e-4 * wl.extent(Tg[b])
        assert st[b]["iters"] <= 50 and st[b]["degenerate"] == 0
    for b in (0, B // 2, B - 1):      # index parity of the final search, per pair
        oi, _ = orc.KdTree(Tg[b]).nn(orc.transform_points(T[b], S[b]))
        assert np.array_equal(idx[b], oi)


def test_device_resident_inputs_on_caller_stream(cat, cat_out):
    """icpsymm_create_on_stream + set_*_dev: clouds already in HBM (torch tensors), caller's stream"""
    torch = pytest.importorskip("torch")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        d_src = torch.zeros((len(cat), 4), dtype=torch.float32, device="cuda")      # 16-byte pcl::PointXYZ rows
        d_src[:, :3] = torch.from_numpy(cat).cuda()
        d_tgt = torch.from_numpy(cat_out).cuda().contiguous()
    stream.synchronize()
    c = icp.Context(0, stream=stream.cuda_stream)
    try:
        assert c.stream == stream.cuda_stream
        c.set_source_dev(d_src.data_ptr(), len(cat), 16)
        c.set_target_dev(d_tgt.data_ptr(), len(cat_out), 12)
        idx, d2 = c.find_correspondences()
        oi, od = orc.nn_brute(cat, cat_out)
        assert np.array_equal(idx, oi) and np.array_equal(d2, od)
        c.estimate_normals(10)
        T, st, _ = c.register_symm(50, 1.0)
        assert st["iters"] == 5 and st["converged"] == 1
    finally:
        c.close()
    # the legacy default stream cannot be captured: the loop must fall back to plain launches, same result
    c0 = icp.Context(0, stream=0)
    try:
        c0.set_source(cat); c0.set_target(cat_out); c0.estimate_normals(10)
        T0, st0, _ = c0.register_symm(50, 1.0)
        assert np.array_equal(T0, T) and st0["iters"] == 5
    finally:
        c0.close()


@pytest.mark.parametrize("case", ["identical", "line", "two_clusters", "huge_offset", "tiny_scale", "grid_plane", "mixed_scales"])
def test_nn_degenerate_distributions(ctx, case):
    """edge distributions for the implicit hierarchy: every one must stay bit-exact against brute force"""
    rng = np.random.default_rng(hash(case) % 2**32)
    n, m = 3000, 5000
    if case == "identical":                       # all target points coincide: every index ties -> index 0
        tgt = np.tile(np.float32([1.5, -2.0, 0.25]), (m, 1)); src = rng.standard_normal((n, 3)).astype(np.float32)
    elif case == "line":                          # zero extent in two axes
        tgt = np.c_[np.linspace(-5, 5, m), np.zeros(m), np.zeros(m)].astype(np.float32)
        src = np.c_[rng.uniform(-6, 6, n), rng.normal(0, .1, n), rng.normal(0, .1, n)].astype(np.float32)
    elif case == "two_clusters":
        tgt = np.concatenate([rng.normal(0, 1e-3, (m // 2, 3)), rng.normal(0, 1e-3, (m - m // 2, 3)) + 1000.0]).astype(np.float32)
        src = np.concatenate([rng.normal(0, 1.0, (n // 2, 3)), rng.normal(0, 1.0, (n - n // 2, 3)) + 1000.0]).astype(np.float32)
    elif case == "huge_offset":                   # fp32 spacing at 1e6 is 0.06: many exact ties
        tgt = (rng.uniform(0, 50, (m, 3)) + 1e6).astype(np.float32); src = (rng.uniform(0, 50, (n, 3)) + 1e6).astype(np.float32)
    elif case == "tiny_scale":
        tgt = (rng.standard_normal((m, 3)) * 1e-12).astype(np.float32); src = (rng.standard_normal((n, 3)) * 1e-12).astype(np.float32)
    elif case == "grid_plane":                    # integer lattice in a plane: equidistant neighbours everywhere
        g = np.stack(np.meshgrid(np.arange(70), np.arange(70), indexing="ij"), -1).reshape(-1, 2)
        tgt = np.c_[g, np.zeros(len(g))].astype(np.float32)
        src = np.c_[rng.integers(0, 139, (n, 2)) * 0.5, rng.integers(-2, 3, n)].astype(np.float32)
    else:                                         # a dense blob inside a sparse shell
        tgt = np.concatenate([rng.normal(0, 1e-4, (m // 2, 3)), rng.normal(0, 100.0, (m - m // 2, 3))]).astype(np.float32)
        src = np.concatenate([rng.normal(0, 1e-3, (n // 2, 3)), rng.normal(0, 150.0, (n - n // 2, 3))]).astype(np.float32)
    idx, _ = _check_nn(ctx, src, tgt)
    if case == "identical":
        assert (idx == 0).all()
    # the same clouds must also survive normal estimation and a registration without NaNs
    ctx.estimate_normals(10)
    assert np.isfinite(ctx.get_normals(0)).all() and np.isfinite(ctx.get_normals(1)).all()
    T, st, rc = ctx.register_symm(3, 0.0)
    assert np.isfinite(T).all() and np.isfinite(st["diff"])


@pytest.mark.parametrize("case", ["signs", "duplicates", "lattice", "huge_range", "tiny", "constant_axis", "two_values"])
def test_nn_large_targets_with_awkward_coordinates(ctx, case):
    """targets above 8192 points go through the coarse k-d partition (radix-select medians, stable
    partition): whatever the key distribution, it must only permute the points -- searches stay bit-exact,
    every original index is returned by some query when the queries are the targets themselves"""
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    m, n = 3 * 8192 + 777, 3000
    if case == "signs":                 # coordinates on both sides of zero: ordered keys span more than 2^31
        tgt = rng.uniform(-3, 3, (m, 3))
    elif case == "duplicates":          # 64 distinct positions only: every median is a tie
        tgt = rng.integers(0, 4, (m, 3)).astype(np.float64)
    elif case == "lattice":
        g = np.stack(np.meshgrid(np.arange(30), np.arange(30), np.arange(30), indexing="ij"), -1).reshape(-1, 3)
        tgt = g[rng.permutation(len(g))[:m]] * 0.5
    elif case == "huge_range":          # exponents from 1e-20 to 1e+18 with both signs
        tgt = rng.standard_normal((m, 3)) * 10.0 ** rng.uniform(-20, 18, (m, 1))
    elif case == "tiny":
        tgt = rng.standard_normal((m, 3)) * 1e-30
    elif case == "constant_axis":       # zero extent along two axes
        tgt = np.c_[rng.standard_normal(m), np.full(m, 7.25), np.zeros(m)]
    else:                               # each coordinate takes two values: 8 distinct points
        tgt = rng.integers(0, 2, (m, 3)) * 1e3 - 500.0
    tgt = tgt.astype(np.float32)
    q = (tgt[rng.integers(0, m, n)] * rng.uniform(0.9, 1.1, (n, 1)) + rng.standard_normal((n, 3)) * 1e-3 * (np.abs(tgt).max() if case != "huge_range" else 1.0)).astype(np.float32)
    _check_nn(ctx, q, tgt)
    # the targets as queries: with distinct points every index comes back as its own neighbour
    if case == "signs":
        ctx.set_source(tgt)
        idx, d2 = ctx.find_correspondences()
        assert np.array_equal(idx, np.arange(m)) and not d2.any()
    if case != "huge_range":            # (squared distances across 38 decades overflow fp32: searches stay exact, moments do not)
        ctx.estimate_normals(10)
        assert np.isfinite(ctx.get_normals(1)).all()


def test_normals_knn_ties_by_lowest_index(ctx):
    """lattice cloud: the 10th neighbour is one of many equidistant points; the kNN set must follow the oracle's
    (d2, index) order, so the covariance and hence the normal agree wherever the spectrum is not degenerate"""
    g = np.stack(np.meshgrid(np.arange(40), np.arange(40), indexing="ij"), -1).reshape(-1, 2)
    P = np.c_[g, 0.01 * (g[:, 0] % 3)].astype(np.float32)
    P = P[np.random.default_rng(3).permutation(len(P))]
    ctx.set_source(P); ctx.set_target(P)
    ctx.estimate_normals(10, viewpoint=(0, 0, 100))
    dn = ctx.get_normals(0)
    on, gap = orc.estimate_normals(P, 10, viewpoint=(0, 0, 100))
    ok = gap > 1e-3
    assert ok.mean() > 0.5
    assert _angle(dn[ok, :3], on[ok, :3]).max() < 1e-4
    assert np.allclose(dn[ok, 3], on[ok, 3], atol=2e-5)


def test_async_setup_overlaps_without_changing_results(cat, cat_out):
    """ICPSYMM_OPT_ASYNC_SETUP: per-cloud streams, no host waits in set_* / estimate_normals; same bits out,
    and a NaN input is reported by the next call that returns data"""
    a, b = icp.Context(0), icp.Context(0)
    try:
        b.set_option(icp.OPT_ASYNC_SETUP, 1)
        for c in (a, b):
            c.set_source(cat); c.set_target(cat_out); c.estimate_normals(10)
        Ta, sa, _ = a.register_symm(50, 1.0)
        Tb, sb, _ = b.register_symm(50, 1.0)
        assert np.array_equal(Ta, Tb) and sa["diff"] == sb["diff"] and sa["iters"] == sb["iters"] == 5
        assert np.array_equal(a.get_normals(1), b.get_normals(1))
        for _ in range(3):                                    # re-upload while the previous result is still being read
            b.set_source(cat); b.set_target(cat_out); b.estimate_normals(10)
            Tc, sc, _ = b.register_symm(50, 1.0)
            assert np.array_equal(Tc, Ta)
        bad = cat_out.copy(); bad[11, 2] = np.inf
        b.set_target(bad)                                     # returns at once: nothing has been checked yet
        b.estimate_normals(10)
        with pytest.raises(icp.IcpError) as e:
            b.register_symm(5, 1.0)
        assert e.value.status == icp.ERR_NONFINITE
        b.set_target(cat_out); b.estimate_normals(10)         # the context recovers
        Td, sd, _ = b.register_symm(50, 1.0)
        assert np.array_equal(Td, Ta)
    finally:
        a.close(); b.close()


def test_async_registration_pipelines_two_contexts(cat, cat_out):
    """icpsymm_register_symm_async / icpsymm_wait: the loop is only enqueued, so one host thread keeps two contexts busy
    (pair k+1 is uploaded, indexed and started while pair k iterates); results are those of the blocking call"""
    s2, t2, _ = wl.config_synth(50_000)
    pairs = [(cat, cat_out), (s2, t2), (cat_out, cat)]
    ref = []
    with icp.Context(0) as c:
        for s, t in pairs:
            c.set_source(s); c.set_target(t); c.estimate_normals(10)
            T, st, _ = c.register_symm(50, 1.0)
            ref.append((T.copy(), st["iters"], st["diff"]))
    pool = [icp.Context(0), icp.Context(0)]
    try:
        for c in pool:
            c.set_option(icp.OPT_ASYNC_SETUP, 1)
        got = [None] * len(pairs)
        for k, (s, t) in enumerate(pairs):
            c = pool[k % 2]
            if k >= 2:
                got[k - 2] = c.wait()
            c.set_source(s); c.set_target(t); c.estimate_normals(10)
            c.register_symm_async(50, 1.0)
            with pytest.raises(icp.IcpError) as e:          # one registration in flight per context
                c.register_symm_async(50, 1.0)
            assert e.value.status == icp.ERR_INVALID
        for k in range(max(0, len(pairs) - 2), len(pairs)):
            got[k] = pool[k % 2].wait()
        for (T, st, _), (Tr, it, diff) in zip(got, ref):
            assert np.array_equal(T, Tr) and st["iters"] == it and st["diff"] == diff
        with pytest.raises(icp.IcpError) as e:
            pool[0].wait()                                   # nothing in flight
        assert e.value.status == icp.ERR_NOT_READY
    finally:
        for c in pool:
            c.close()


def test_register_p2p_matches_the_oracle_loop(ctx, cat, cat_out, gold):
    """icpsymm_register_p2p (MyICP::RegisterP2P): exact correspondences + registrateNPoint per iteration, on the device,
    against the same loop made of oracle pieces (kd-tree NN, regist.h restatement)"""
    ctx.set_source(cat); ctx.set_target(cat_out)
    T, st, _ = ctx.register_p2p(60, 0.5)
    Tt, tree = np.eye(4), orc.KdTree(cat_out)
    for it in range(61):
        p = orc.transform_points(Tt.astype(np.float32), cat)
        idx, d2 = tree.nn(p)
        diff = np.sqrt(d2.astype(np.float64)).sum()
        if not diff > 0.5 or it >= 60:
            break
        rc, R, t = orc.registrate_npoint(p.astype(np.float64), cat_out[idx].astype(np.float64))
        inc = np.eye(4); inc[:3, :3] = R; inc[:3, 3] = t
        Tt = inc @ Tt
    assert st["iters"] == it and st["searches"] == it + 1
    assert wl.rotation_angle(T, Tt) < 1e-5 and np.abs(T[:3, 3] - Tt[:3, 3]).max() < 1e-5 * wl.extent(cat)
    assert st["diff"] == pytest.approx(diff, rel=1e-5)
    di, dd = ctx.get_correspondences()
    oi, od = tree.nn(orc.transform_points(T, cat))
    assert np.array_equal(di, oi) and np.array_equal(dd, od)
    with pytest.raises(icp.IcpError):
        ctx.set_source(cat[:2]); ctx.register_p2p(5, 0.5)            # regist.h:9: fewer than 3 pairs
