"""Pins the oracle to the reference's own known-answer tests, read from the committed fixture
tests/golden/reference_kats.json (parsed out of the reference's test sources by
scripts/make_golden.py: traj_utils/test/test_bernstein.cpp, traj_opt/test/test_bezier_opt.cpp,
utils/separator/src/test_separator.cpp), and checks that the oracle still reproduces the committed
config-1 vectors (tests/golden/oracle_c1.npz) the GPU tests are compared with."""
import json
import os

import numpy as np
import pytest

import pred_occ_planner_b200 as p
from tests import cases, oracle_cycle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def kats():
    return json.load(open(os.path.join(GOLD, "reference_kats.json")))


def test_bernstein_piece_exact(oracle, kats):
    k = kats["bernstein"]
    c = np.array(k["piece_cpts"])
    deriv = {"getPos": 0, "getVel": 1, "getAcc": 2}
    assert len(k["piece_expect"]) == 4
    for e in k["piece_expect"]:   # EXPECT_EQ: exact equality
        got = oracle.piece_eval(c, k["piece_t0"], k["piece_t1"], e["t"], deriv[e["fn"]])
        assert np.array_equal(got, np.array(e["value"])), e
    assert np.array_equal(oracle.bernstein_coeff(), np.array(k["coeff_mat"]))
    # getVelCtrlPts == 4 everywhere, getAccCtrlPts == 0: rates over the piece duration
    dur = k["piece_t1"] - k["piece_t0"]
    assert np.isclose(oracle.bezier_max_rate([dur], c, 1), np.linalg.norm(k["vel_cpts"][0]) / dur, rtol=0, atol=1e-15)
    assert oracle.bezier_max_rate([dur], c, 2) == np.linalg.norm(k["acc_cpts"][0])


def test_bezier_curve_fixture(oracle, kats):
    k = kats["bernstein"]
    dur, cp = k["curve_durations"], np.array(k["curve_cpts"])
    assert sum(dur) == k["curve_expect"]["duration"] and len(dur) == k["curve_expect"]["pieces"]
    assert oracle.bezier_max_rate(dur, cp, 1) > k["curve_expect"]["max_vel_rate_gt"]
    knots = np.concatenate([[0.0], np.cumsum(dur)])
    for i, t in enumerate(knots[:-1]):
        assert np.allclose(oracle.bezier_eval(dur, cp, [t])[0], cp[5 * i])
    assert np.allclose(oracle.bezier_eval(dur, cp, [knots[-1]])[0], cp[-1])


def test_qp_fixture_properties(oracle, kats):
    """TestOpt / TestWaypoints / TestContinuity of test_bezier_opt.cpp: solved, boundary conditions and
    C0-C2 continuity within 1e-3; control points inside their cubes up to the ADMM accuracy."""
    k = kats["bezier_opt"]
    cfg = cases.cfg_c1(p.POP_CELL_F32)
    st, X, it = oracle.qp_solve(cfg, k["start"], k["end"], k["t"], k["polys"])
    assert st in (1, 2)   # BezierOpt::optimize() accepts solved and solved-inaccurate (bezier_optimizer.cpp:279)
    tol, T = k["tolerance"], k["t"]
    tot = sum(T)
    assert np.abs(oracle.bezier_eval(T, X, [0.0])[0] - k["start"][:3]).max() < tol
    assert np.abs(oracle.bezier_eval(T, X, [0.0], 1)[0] - k["start"][3:6]).max() < tol
    assert np.abs(oracle.bezier_eval(T, X, [tot])[0] - k["end"][:3]).max() < tol
    assert np.abs(oracle.bezier_eval(T, X, [tot], 1)[0] - k["end"][3:6]).max() < tol
    knots = np.cumsum(T)[:-1]
    for d in (0, 1, 2):
        left = oracle.bezier_eval(T, X, knots - 1e-9, d)
        right = oracle.bezier_eval(T, X, knots + 1e-9, d)
        assert np.abs(left - right).max() < 10 ** (d + 1) * tol
    for i, poly in enumerate(k["polys"]):
        h = np.array(poly)
        pts = X[5 * i:5 * i + 5]
        assert (pts @ h[:, :3].T + h[:, 3] <= 2e-2).all()   # not asserted by the reference; ADMM accuracy at eps 1e-3


def test_separator_instance(oracle, kats):
    k = kats["separator"]
    assert oracle.separable(k["points_a"], k["points_b"]) == k["expect_solved"]
    # mixing the sets makes them inseparable
    assert not oracle.separable(k["points_a"] + k["points_b"][:1], k["points_b"])


def test_oracle_reproduces_committed_c1_vectors(oracle):
    g = np.load(os.path.join(GOLD, "oracle_c1.npz"))
    cfg = cases.cfg_c1(p.POP_CELL_F32)
    cloud, cyl = cases.world_c1()
    om = oracle.map(cfg)
    om.set_pose(g["start"][:3].astype(np.float32))
    om.update_gt(cloud, cyl, float(g["t_map"]))
    occ = np.argwhere(om.risk_maps() > 0).astype(np.int32)
    assert np.array_equal(occ, g["occupied"])
    assert np.array_equal(om.query(g["query_pos"], g["query_t"]), g["query_out"])
    ret, path, _ = om.astar(g["start"], g["goal"], float(g["t_plan"] - g["t_map"]))
    assert ret == int(g["astar_ret"]) and np.array_equal(path, g["astar_path"])
    st, rec, _ = oracle_cycle.replan(om, None, 0, g["start"], g["goal"], float(g["t_plan"]))
    assert st == int(g["replan_status"]) and rec["n_pieces"] == int(g["record_n_pieces"])
    assert np.allclose(rec["cpts"], g["record_cpts"], rtol=0, atol=1e-12)
