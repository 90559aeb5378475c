"""Pins of the oracle against the REFERENCE's own code.

`oracle/_ref/libref.so` is the reference's plan_env/src/raycast.cpp, traj_utils/sdlp.hpp, sfc_gen/lbfgs.hpp,
sfc_gen/firi.hpp, traj_utils/src/bernstein.cpp, traj_opt/src/bezier_optimizer.cpp and
path_searching/src/{fake_,}risk_hybrid_a_star.cpp, traj_coordinator/src/particles.cpp compiled unmodified (oracle/Makefile; Eigen replaced by the stand-in
oracle/ref_build/Eigen/Eigen, the OSQP wrapper by a recorder, ROS and the map class by minimal stand-ins).
Every test runs the oracle's restatement on seeded cases and compares it with

  * the committed outputs of that library (tests/golden/ref_pins.npz, scripts/make_golden_ref.py) — always;
  * the library itself on a larger sweep — where it exists (this container, and the GPU box it travels to).

What is exact and what is not: the DDA is integer stepping on scalar doubles — identical voxel sequences.  costMVIE,
the Lewis-Overton line search and the L-BFGS loop are deterministic given their inputs — identical to the last bit
(status, cost, all nine variables).  The LP shuffles its constraints (the reference with a function-static
mt19937_64, the oracle with a caller-owned LCG), so optima agree to 1e-7, not bitwise.  maxVolInsEllipsoid starts
L-BFGS from the LP's interior point and stops it on a 1e-7 relative cost decrease, so two correct runs agree in
volume to ~1e-3 and differ more along flat directions; firi's first polytope (before any ellipsoid) is identical,
later ones are compared as sets (both valid corridors, intersection-over-union).  Bezier evaluation is three small
matrix products whose summation order the oracle fixes one way and the stand-in another: 1e-14 relative (99 % of the
samples are identical).  BezierOpt::setup only places and scales constants: Q, A, lb, ub identical.  The OSQP solver
itself is third-party and absent; its restatement in oracle/o_qp.cpp is NOT pinned here (tests/test_golden_reference.py
holds the reference's own QP fixture).  The hybrid A* is a deterministic search over doubles: return code, expansion
and node counts and every sampled path state identical (its occupancy queries are answered by the oracle's grid)."""
import os

import numpy as np
import pytest

from tests import ref_lib as rl

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_pins.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def orc(oracle):
    return rl.OrcHooks(oracle)


@pytest.fixture(scope="module")
def ref():
    return rl.load()


def test_golden_file_is_current(ref, gold):
    """Where the reference library can be built, the committed vectors are what it produces today."""
    if ref is None:
        pytest.skip("oracle/_ref/libref.so not present (no /root/reference here)")
    paths = rl.unpack(gold["ray_path"], gold["ray_offs"], 3)
    for (s, e), want in zip(rl.ray_cases(40), paths):
        assert np.array_equal(ref.raycaster(s, e)[0], want)
    for (A, x0), want in zip(rl.lbfgs_cases(10), gold["lbfgs"]):
        ret, x, f = ref.lbfgs_mvie(A, x0)
        assert ret == want[0] and f == want[1] and np.array_equal(x, want[2:])


# ---- RayCaster (plan_env/src/raycast.cpp:242-335) -----------------------------------------------------------------
def test_raycaster_voxel_sequences_identical(orc, ref, gold):
    cases = rl.ray_cases(len(gold["ray_ok"]))
    paths = rl.unpack(gold["ray_path"], gold["ray_offs"], 3)
    total = 0
    for (s, e), want, ok in zip(cases, paths, gold["ray_ok"]):
        got, gok = orc.raycaster(s, e)
        assert gok == ok and np.array_equal(got, want)
        total += len(got)
    assert total > 20000
    if ref is not None:
        for s, e in rl.ray_cases(5000, seed=17):
            a, b = ref.raycaster(s, e), orc.raycaster(s, e)
            assert a[1] == b[1] and np.array_equal(a[0], b[0])


# ---- sdlp::linprog (traj_utils/sdlp.hpp:708-785) ------------------------------------------------------------------
def _lp_check(v_ref, x_ref, v, x, A, b):
    unb_ref = not np.isfinite(v_ref) or abs(v_ref) > 1e12
    unb = not np.isfinite(v) or abs(v) > 1e12
    assert unb_ref == unb
    if unb:
        if np.isposinf(v_ref) or np.isposinf(v):      # infeasible is +inf on both sides
            assert v_ref == v
        return 0
    assert abs(v - v_ref) <= 1e-9 * max(1.0, abs(v_ref))
    assert np.abs(x - x_ref).max() <= 1e-7
    assert (A @ x - b).max() <= 1e-9
    return 1


def test_linprog_matches_reference(orc, ref, gold):
    finite = 0
    for i, (c, A, b) in enumerate(rl.lp_cases(len(gold["lp_val"]))):
        v, x = orc.linprog(c, A, b, seed=i + 1)
        finite += _lp_check(gold["lp_val"][i], gold["lp_x"][i][:len(c)], v, x, A, b)
    assert finite > 100
    if ref is not None:
        for i, (c, A, b) in enumerate(rl.lp_cases(1500, seed=23)):
            vr, xr = ref.linprog(c, A, b)
            v, x = orc.linprog(c, A, b, seed=i + 7)
            _lp_check(vr, xr, v, x, A, b)


# ---- firi::costMVIE, lbfgs::lbfgs_optimize (firi.hpp:74-137, lbfgs.hpp:276-720) -------------------------------------
def test_mvie_cost_and_gradient_bit_identical(orc, ref, gold):
    for (A, x), want in zip(rl.mvie_cost_cases(len(gold["cost"])), gold["cost"]):
        c, g = orc.cost_mvie(A, x)
        assert c == want[0] and np.array_equal(g, want[1:])
    if ref is not None:
        for A, x in rl.mvie_cost_cases(1000, seed=31):
            cr, gr = ref.cost_mvie(A, x)
            c, g = orc.cost_mvie(A, x)
            assert c == cr and np.array_equal(g, gr)


def test_lbfgs_runs_bit_identical(orc, ref, gold):
    """Same data, same start: the whole optimisation (every line search, every two-loop update, the stop test) lands on
    the same nine doubles, the same cost and the same status."""
    for (A, x0), want in zip(rl.lbfgs_cases(len(gold["lbfgs"])), gold["lbfgs"]):
        ret, x, f = orc.lbfgs_mvie(A, x0)
        assert ret == want[0] and f == want[1] and np.array_equal(x, want[2:])
    if ref is not None:
        for A, x0 in rl.lbfgs_cases(300, seed=37):
            a, b = ref.lbfgs_mvie(A, x0), orc.lbfgs_mvie(A, x0)
            assert a[0] == b[0] and a[2] == b[2] and np.array_equal(a[1], b[1])


# ---- firi::maxVolInsEllipsoid (firi.hpp:143-214) ------------------------------------------------------------------
def test_mvie_same_optimum(orc, ref, gold):
    rel = []
    for k, (h, want) in enumerate(zip(rl.poly_cases(len(gold["mvie"])), gold["mvie"])):
        ok, R, p, r = orc.mvie(h, np.eye(3), np.zeros(3), np.full(3, 0.3), seed=k + 1)
        vol, Q, viol = rl.ellipsoid(R, p, r, h)
        assert ok == want[0] == 1
        assert abs(np.linalg.det(R) - 1.0) < 1e-12                 # a proper rotation, as the reference enforces
        assert viol < 2e-3 and want[2] < 2e-3                       # inscribed up to the penalty's softness
        rel.append(abs(vol - want[1]) / want[1])
        assert np.abs(p - want[3:6]).max() < 5e-2
        assert np.abs(Q - want[6:].reshape(3, 3)).max() <= 5e-2 * np.abs(Q).max()
    assert max(rel) < 2e-3 and np.median(rel) < 2e-5


# ---- firi::firi (firi.hpp:216-365) --------------------------------------------------------------------------------
def test_firi_first_polytope_identical(orc, ref, gold):
    """One iteration = the polytope construction alone (tangent planes, boundary/obstacle ordering, back-transform),
    from the caller's radii: identical rows in identical order."""
    cases = rl.firi_cases(len(gold["firi1_offs"]) - 1)
    for k, ((bd, pc, a, b, r0, lo, hi), want) in enumerate(zip(cases, rl.unpack(gold["firi1"], gold["firi1_offs"], 4))):
        got, _ = orc.firi(bd, pc, a, b, 1, r0, seed=k + 1)
        assert got.shape == want.shape and np.array_equal(got, want)
    if ref is not None:
        for k, (bd, pc, a, b, r0, lo, hi) in enumerate(rl.firi_cases(200, seed=41)):
            assert np.array_equal(orc.firi(bd, pc, a, b, 1, r0, seed=k + 1)[0], ref.firi(bd, pc, a, b, 1, r0)[0])


@pytest.mark.parametrize("iterations", [2, 4])
def test_firi_iterated_same_corridor(orc, gold, iterations):
    cases = rl.firi_cases(len(gold["firi1_offs"]) - 1)
    wants = rl.unpack(gold[f"firi{iterations}"], gold[f"firi{iterations}_offs"], 4)
    ious = []
    for k, ((bd, pc, a, b, r0, lo, hi), want) in enumerate(zip(cases, wants)):
        got, _ = orc.firi(bd, pc, a, b, iterations, r0, seed=k + 1)
        for h in (got, want):                                       # both are valid corridors of the same problem
            assert (h @ np.r_[a, 1] <= 1e-9).all() and (h @ np.r_[b, 1] <= 1e-9).all()
            if len(pc):
                assert ((np.c_[pc, np.ones(len(pc))] @ h.T).max(1) > -1e-5).all()
        assert abs(len(got) - len(want)) <= 3
        ious.append(rl.poly_iou(got, want, lo, hi))
    assert np.median(ious) > 0.999 and min(ious) > 0.85


def test_firi_refuses_like_reference(orc, ref):
    """a or b outside the boundary polytope: firi returns false (firi.hpp:227-229)."""
    bd, pc, a, b, r0, lo, hi = rl.firi_cases(1)[0]
    out = a.copy()
    out[0] = hi[0] + 0.5
    assert len(orc.firi(bd, pc, out, b, 2)[0]) == 0
    if ref is not None:
        assert len(ref.firi(bd, pc, out, b, 2)[0]) == 0


# ---- Bernstein::Bezier (traj_utils/src/bernstein.cpp) ---------------------------------------------------------------
def test_bezier_evaluation_matches_reference(oracle, ref, gold):
    cases = rl.bezier_cases(len(gold["bezier_rates"]))
    wants = rl.unpack(gold["bezier"], gold["bezier_offs"])
    same = total = 0
    for (dur, cp, t), want, rates in zip(cases, wants, gold["bezier_rates"]):
        want = want.reshape(3, len(t), 3)
        for d in (0, 1, 2):
            got = oracle.bezier_eval(dur, cp, t, deriv=d)
            assert np.abs(got - want[d]).max() <= 1e-14 * max(1.0, np.abs(want[d]).max())
            same += int(np.array_equal(got, want[d]))
            total += 1
        for o in (1, 2):
            assert abs(oracle.bezier_max_rate(dur, cp, o) - rates[o - 1]) <= 1e-14 * max(1.0, rates[o - 1])
    assert same >= 0.9 * total
    if ref is not None:
        for dur, cp, t in rl.bezier_cases(300, seed=43):
            for d in (0, 1, 2):
                a, b = ref.bezier_eval(dur, cp, t, d), oracle.bezier_eval(dur, cp, t, deriv=d)
                assert np.abs(a - b).max() <= 1e-14 * max(1.0, np.abs(a).max())


# ---- traj_opt::BezierOpt::setup (traj_opt/src/bezier_optimizer.cpp:26-250) -----------------------------------------
def test_qp_assembly_identical(oracle, ref, gold):
    """Min-jerk cost, continuity / dynamical / corridor rows and both bound vectors exactly as the reference builds them
    (same row order, same scaling by the piece durations, -OSQP_INFTY = -1e30 on the one-sided rows)."""
    wants = rl.unpack(gold["qp"], gold["qp_offs"])
    for (init, goal, dur, polys, vmax, amax), want in zip(rl.qp_cases(len(wants), max_pieces=4), wants):
        got = oracle.qp_setup(init, goal, dur, polys, vmax=vmax, amax=amax)
        assert np.array_equal(np.concatenate([a.ravel() for a in got]), want)
    if ref is not None:
        for init, goal, dur, polys, vmax, amax in rl.qp_cases(60, seed=47):
            got, want = oracle.qp_setup(init, goal, dur, polys, vmax=vmax, amax=amax), ref.qp_setup(init, goal, dur, polys, vmax, amax)
            assert all(np.array_equal(a, b) for a, b in zip(got, want))


# ---- FakeRiskHybridAstar (path_searching/src/fake_risk_hybrid_a_star.cpp) -------------------------------------------
@pytest.mark.parametrize("risk", [False, True])
def test_hybrid_astar_identical(oracle, ref, gold, risk):
    """The reference's own search — motion primitives, pruning by (voxel, time slice), heuristic via the quartic,
    one-shot trajectories, open-set order — and the oracle's restatement, on the same grid with moving obstacles:
    same ASTAR_RET, same number of expansions and nodes, same retry, the same path to the last bit.  Both classes:
    FakeRiskHybridAstar (ground-truth map) and RiskHybridAstar (RiskBase map; its one-shot check looks at slice 0)."""
    cfg, (cloud, cyl) = rl.astar_world_risk() if risk else rl.astar_world()
    key = "astar_risk" if risk else "astar"
    heads = gold[key + "_head"]
    paths = rl.unpack(gold[key + "_path"], gold[key + "_offs"], 6)
    found = 0
    for (pose, start, goal, t0), head, want in zip(rl.astar_cases(len(heads), seed=9 if risk else 8), heads, paths):
        m = oracle.map(cfg)
        m.set_pose(pose)
        m.update_gt(cloud, cyl, 0.0)
        ret, pth, st = m.astar(start, goal, t0, cap=256)
        assert [ret, st[0], st[1], st[2]] == head.tolist()
        assert pth.shape == want.shape and np.array_equal(pth, want)
        found += int(ret != 0)
    assert found >= len(heads) // 2
    if ref is not None:
        for pose, start, goal, t0 in rl.astar_cases(60, seed=53):
            m = oracle.map(cfg)
            m.set_pose(pose)
            m.update_gt(cloud, cyl, 0.0)
            a, b = ref.astar(cfg, m, pose, start, goal, t0, risk=risk), m.astar(start, goal, t0, cap=256)
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2][:3], b[2][:3])


# ---- ParticleATC (traj_coordinator/src/particles.cpp:62-87, 131-191, 291-422) ---------------------------------------
def test_peer_particles_identical(oracle, ref, gold):
    """The swarm coordinator's side of the peer projection: body particles (own: exact, peers': float32 on the wire),
    trajectories as received, and for every (peer, time) the particles getParticlesWithRisk hands to the map — with
    its rules for a peer that has not published, has not started, has finished, or is the ego drone.  Identical."""
    import ctypes as C
    cfg, size, ego, offs, counts, recs, qa, qt = rl.peer_case()
    ego_pts = np.zeros((64, 3))
    n = oracle.lib.orc_ego_particles(rl.P(size), rl.P(ego_pts), 64)
    assert np.array_equal(ego_pts[:n], gold["peer_ego"])
    got = rl.peer_queries(oracle.lib.orc_particles_with_risk,
                          (C.byref(cfg), ego, rl.P(offs), counts.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(recs.ctypes.data)),
                          offs.shape[1], qa, qt)
    assert np.array_equal(got[2], gold["peer_n"]) and np.array_equal(got[0], gold["peer_pts"])
    assert np.array_equal(got[1], gold["peer_risks"])
    assert (got[2] > 0).sum() > 50 and (got[2] < 0).sum() > 50
    if ref is not None:
        cfg, size, ego, offs, counts, recs, qa, qt = rl.peer_case(seed=59, n=10, ego=7, nq=500)
        a = ref.particles_with_risk(10, ego, size, offs, counts, recs, 10.0, qa, qt)
        b = rl.peer_queries(oracle.lib.orc_particles_with_risk,
                            (C.byref(cfg), ego, rl.P(offs), counts.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(recs.ctypes.data)),
                            offs.shape[1], qa, qt)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_separation_check_point_sets_identical(oracle, ref, gold):
    """isSafeAfterOpt up to the LP (the GLPK solver is absent): which peers are checked at a given time and which of
    their control points are left — identical.  The LP itself: tests/test_golden_reference.py (the reference's own
    separable sets) and tests/test_oracle_corridor_qp.py."""
    import ctypes as C
    cfg, size, ego, offs, counts, recs, qa, qt = rl.peer_case()
    cpts, dur, M, nows = rl.sep_case()
    for k, now in enumerate(nows):
        n, sets = rl.sep_sets(oracle.lib.orc_safe_after_opt_sets,
                              (C.byref(cfg), rl.P(offs), counts.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(recs.ctypes.data),
                               ego, rl.P(cpts), M, rl.D(now)), ())
        assert n == gold["sep_n"][k] and np.array_equal(sets, gold["sep_sets"][k])
        if ref is not None:
            r = rl.ref_sep_sets(ref, cfg.n_agents_total, ego, size, offs, counts, recs, now, cpts, dur, M)
            assert r[0] == n and np.array_equal(r[1], sets)
    assert gold["sep_n"].max() >= 5 and gold["sep_n"].min() == 0


# ---- FakeBaselinePlanner::replan / BaselinePlanner::replan (plan_manager/src/baseline_fake.cpp, baseline.cpp) -------
@pytest.mark.parametrize("risk", [False, True])
def test_whole_replan_against_the_reference_planner(oracle, ref, gold, risk):
    """The reference's planners end to end — their own A*, corridor stage (bounding boxes, obstacle points per segment,
    FIRI, shrink, validity / intersection / goal LPs), QP assembly and trajectory hand-over — run on the oracle's grid
    with the oracle's ADMM behind their solver wrapper, against the oracle's replan on problems with moving obstacles:
    the same replans succeed and fail, with the same number of A* expansions and trajectory pieces; control points
    agree to the solver's own tolerance (eps 1e-3; FIRI's ellipsoid step stops L-BFGS on a 1e-7 cost decrease, so
    corridor faces of two correct runs differ by ~1e-4 and the QP optimum moves with them) — median below 1e-4.
    risk=False: FakeBaselinePlanner on the ground-truth map; risk=True: BaselinePlanner on a RiskBase-style map, whose
    rules differ (every face shrunk, corridor cut and count, goal search)."""
    from tests import oracle_cycle
    cfg, (cloud, cyl) = rl.astar_world_risk() if risk else rl.astar_world()
    key, seed = ("replan_risk", 79) if risk else ("replan", 67)
    heads = gold[key + "_head"]
    wants = rl.unpack(gold[key + "_cpts"], gold[key + "_offs"], 3)
    rel, n_ok = [], 0
    for (pose, start, goal, ts), head, want in zip(rl.astar_cases(len(heads), seed=seed), heads, wants):
        m = oracle.map(cfg)
        m.set_pose(pose)
        m.update_gt(cloud, cyl, 0.0)
        m.set_time(100.0)
        rs, rec, st = oracle_cycle.replan(m, None, 0, start, goal, 100.0 + ts)
        if head[0] < 0:                                                # not run: see scripts/make_golden_ref.py
            assert rs == 2
            continue
        assert (rs == 0) == bool(head[0]), (rs, head.tolist())
        assert st[1] == head[2]                                        # A* expansions
        if rs == 0:
            M = int(rec["n_pieces"])
            assert M == head[1]
            got = np.asarray(rec["cpts"]).reshape(-1, 3)[:5 * M]
            rel.append(np.abs(got - want).max() / max(1.0, np.abs(want).max()))
            n_ok += 1
    assert n_ok >= 10 and max(rel) < 1e-2 and np.median(rel) < 1e-4, (n_ok, max(rel), np.median(rel))
    if ref is not None:
        for pose, start, goal, ts in rl.astar_cases(15, seed=73):
            m = oracle.map(cfg)
            m.set_pose(pose)
            m.update_gt(cloud, cyl, 0.0)
            m.set_time(100.0)
            rs, rec, so = oracle_cycle.replan(m, None, 0, start, goal, 100.0 + ts)
            if rs == 2:
                continue
            ok, M, cp, st = ref.replan(cfg, m, pose, 100.0, 100.0 + ts, start, goal, risk=risk)
            assert (rs == 0) == bool(ok) and so[1] == st[0]
            if ok:
                assert int(rec["n_pieces"]) == M
                got = np.asarray(rec["cpts"]).reshape(-1, 3)[:5 * M]
                assert np.abs(got - cp).max() <= 1e-2 * max(1.0, np.abs(cp).max())


def test_is_traj_safe_equals_reference(oracle, ref, gold):
    """BaselinePlanner::isTrajSafe (baseline_fake.cpp:53-76; the same code in baseline.cpp:45-68): running-sum sample
    times, clamping of the horizon to the trajectory's duration, queries by time — identical booleans."""
    got, live = [], []
    for cfg, m, pose, t_map, rec, now, T in rl.traj_safe_cases(oracle):
        a, tu = m.traj_safe(rec, cfg.corridor_tau, now, T)
        assert ((tu >= 0) == (a == 0)).all()
        got.append(a)
        if ref is not None:
            live.append(ref.traj_safe(cfg, m, pose, t_map, rec, cfg.corridor_tau, now, T))
            assert np.array_equal(ref.traj_safe(cfg, m, pose, t_map, rec, cfg.corridor_tau, now, T, risk=True), live[-1])
    got = np.concatenate(got)
    assert np.array_equal(got, gold["traj_safe"])
    assert (got == 0).sum() > 10 and (got == 1).sum() > 50
    if live:
        assert np.array_equal(np.concatenate(live), got)
