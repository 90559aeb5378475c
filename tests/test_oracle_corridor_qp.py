"""CPU checks of the oracle's corridor stage (Seidel LP, FIRI/MVIE), Bezier QP and separator.

Pinned by the reference's own fixtures where they exist: the 3-cube QP problem of
traj_opt/test/test_bezier_opt.cpp:47-221 (matrix shapes, solved status, boundary conditions at
1e-3, C1/C2 continuity) and the separable 4-vs-5 point instance of
utils/separator/src/test_separator.cpp:20-32.  LP/FIRI have no reference tests: checked against
scipy's HiGHS and through geometric invariants."""
import json
import os

import numpy as np
import pytest

import pred_occ_planner_b200 as p
from tests import cases


def _cube(lo, hi):
    return [[-1, 0, 0, lo], [1, 0, 0, -hi], [0, -1, 0, lo], [0, 1, 0, -hi], [0, 0, -1, lo], [0, 0, 1, -hi]]


# the reference's own QP fixture (traj_opt/test/test_bezier_opt.cpp:59-99), parsed from its source
# into tests/golden/reference_kats.json by scripts/make_golden.py
_KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))["bezier_opt"]
FIX_POLYS = _KAT["polys"]
FIX_START = _KAT["start"]
FIX_END = _KAT["end"]
FIX_T = _KAT["t"]


def test_lp_matches_highs(oracle):
    from scipy.optimize import linprog as sp_lp
    rng = np.random.default_rng(0)
    for trial in range(120):
        d = int(rng.choice([3, 4]))
        m = int(rng.integers(4, 40))
        A = rng.normal(size=(m, d))
        x0 = rng.normal(size=d)
        b = A @ x0 + rng.uniform(0.1, 2, m)
        A = np.vstack([A, np.eye(d), -np.eye(d)])
        b = np.concatenate([b, x0 + 5, -(x0 - 5)])
        c = rng.normal(size=d)
        v, x = oracle.linprog(c, A, b, seed=trial + 1)
        r = sp_lp(c, A_ub=A, b_ub=b, bounds=[(None, None)] * d, method="highs")
        assert abs(v - r.fun) <= 1e-7 * max(1, abs(r.fun))
        assert (A @ x <= b + 1e-7).all()
    # infeasible bounded problems are reported as +-inf (callers test isinf)
    for trial in range(60):
        A = rng.normal(size=(12, 3))
        b = rng.normal(size=12) * 0.3
        A = np.vstack([A, np.eye(3), -np.eye(3)])
        b = np.concatenate([b, np.full(6, 3.0)])
        v, _ = oracle.linprog(np.zeros(3), A, b, seed=trial + 1)
        r = sp_lp(np.zeros(3), A_ub=A, b_ub=b, bounds=[(None, None)] * 3, method="highs")
        assert (r.status == 0) == (not np.isinf(v))


def test_firi_polytope_invariants(oracle):
    rng = np.random.default_rng(3)
    a, b = np.array([0.0, 0, 1.0]), np.array([0.6, 0.1, 1.0])
    lo, hi = np.array([-1.2, -1.2, 0.0]), np.array([1.8, 1.3, 2.2])
    bd = np.array([[1, 0, 0, -hi[0]], [0, 1, 0, -hi[1]], [0, 0, 1, -hi[2]],
                   [-1, 0, 0, lo[0]], [0, -1, 0, lo[1]], [0, 0, -1, lo[2]]], float)
    pc = rng.uniform(lo, hi, (400, 3))
    # keep obstacle points away from the seed segment
    seg_d = np.linalg.norm(np.cross(pc - a, pc - b), axis=1) / np.linalg.norm(b - a)
    pc = pc[seg_d > 0.35]
    h = oracle.firi(bd, pc, a, b, iterations=2)
    assert 6 <= len(h) <= 48
    ah, bh = np.append(a, 1), np.append(b, 1)
    assert (h @ ah <= 1e-9).all() and (h @ bh <= 1e-9).all()          # seed segment inside
    inside = (np.c_[pc, np.ones(len(pc))] @ h.T <= -1e-7).all(axis=1)  # no obstacle strictly inside
    assert not inside.any()
    # a or b outside the bounding box → firi returns false → no polytope
    assert len(oracle.firi(bd, pc, a, b + [5, 0, 0])) == 0
    # no obstacles → the polytope is the bounding box
    h0 = oracle.firi(bd, np.zeros((0, 3)), a, b)
    assert len(h0) == 6


def test_qp_fixture_shapes_and_solution(oracle):
    Q, A, l, u = oracle.qp_setup(FIX_START, FIX_END, FIX_T, FIX_POLYS)
    assert Q.shape == (45, 45) and A.shape[1] == 45 and len(u) == A.shape[0]  # TestOpt (:137-147)
    # single piece: Q is 15x15 (TestMinJerkCost :129-135)
    Q1, A1, _, _ = oracle.qp_setup(FIX_START, FIX_END, [2.0], [FIX_POLYS[0]])
    assert Q1.shape == (15, 15)
    assert np.allclose(Q, Q.T) and np.linalg.eigvalsh(Q).min() > -1e-9
    cfg = p.default_config()
    st, X, iters = oracle.qp_solve(cfg, FIX_START, FIX_END, FIX_T, FIX_POLYS)
    assert st in (1, 2) and X.shape == (15, 3)                                 # optimize() == true
    T = sum(FIX_T)
    s, e = np.array(FIX_START).reshape(3, 3), np.array(FIX_END).reshape(3, 3)
    for d in range(3):                                                         # TestWaypoints (:156-181)
        assert np.abs(oracle.bezier_eval(FIX_T, X, [0.0], d)[0] - s[d]).max() < 1e-3
        assert np.abs(oracle.bezier_eval(FIX_T, X, [T], d)[0] - e[d]).max() < 1e-3
    # C1 / C2 continuity at the joins (:188-199 probe it by finite differences across t = 2)
    for tj in (2.0, 6.0):
        for d in (1, 2):
            left = oracle.bezier_eval(FIX_T, X, [tj - 1e-9], d)[0]
            right = oracle.bezier_eval(FIX_T, X, [tj + 1e-9], d)[0]
            assert np.abs(left - right).max() < 5e-3
    # finite-difference acceleration consistency (TestContinuity :211-220)
    for t1, t0 in ((1.0, 0.9), (2.0, 1.9)):
        dv = np.linalg.norm(oracle.bezier_eval(FIX_T, X, [t1], 1)[0] - oracle.bezier_eval(FIX_T, X, [t0], 1)[0]) / 0.1
        am = np.linalg.norm(oracle.bezier_eval(FIX_T, X, [t1], 2)[0] + oracle.bezier_eval(FIX_T, X, [t0], 2)[0]) / 2
        assert abs(dv - am) <= 0.01
    # control points respect their corridors up to the ADMM tolerance
    for i, poly in enumerate(FIX_POLYS):
        hp = np.array(poly, float)
        c = np.c_[X[5 * i:5 * i + 5], np.ones(5)]
        assert (c @ hp.T).max() < 2e-2


def test_qp_tight_tolerance_is_the_optimum(oracle):
    """With eps 1e-7 the ADMM restatement reaches the QP optimum (KKT check with scipy's HiGHS-free
    trust-constr is slow; use the equality-constrained reduced problem on the active set)."""
    cfg = p.default_config()
    cfg.osqp_eps_abs = cfg.osqp_eps_rel = 1e-7
    cfg.osqp_max_iter = 200000
    st, X, iters = oracle.qp_solve(cfg, FIX_START, FIX_END, FIX_T, FIX_POLYS)
    assert st == 1
    Q, A, l, u = oracle.qp_setup(FIX_START, FIX_END, FIX_T, FIX_POLYS)
    x = X.reshape(-1)
    Ax = A @ x
    assert (Ax <= u + 1e-5).all() and (Ax >= l - 1e-5).all()
    act = (np.abs(Ax - u) < 1e-5) | (np.abs(Ax - l) < 1e-5)
    Aa, ba = A[act], np.where(np.abs(Ax - u) < 1e-5, u, l)[act]
    K = np.block([[Q, Aa.T], [Aa, np.zeros((len(Aa), len(Aa)))]])
    sol = np.linalg.lstsq(K, np.concatenate([np.zeros(45), ba]), rcond=None)[0]
    assert np.abs(sol[:45] - x).max() < 1e-4


def test_separator_reference_instance(oracle):
    A = [[-3.5, 21, 1.4], [-2.71, 2.13, 1.6], [0.53, 0.51, 1.4], [-3.5, 0.21, 0.3]]
    B = [[-2.3, 4.69, 6.2], [3.7, 2.13, 65.6], [6.5, 2.93, 2.8], [0.3, 4.8, 9.2], [1.5, 6.7, 2.9]]
    assert oracle.separable(A, B)
    assert not oracle.separable(A + [[1, 4, 5]], B + [[0, 3, 1.5]])   # interleaved hulls
    rng = np.random.default_rng(0)
    for _ in range(50):
        P = rng.normal(size=(10, 3))
        Qs = rng.normal(size=(10, 3))
        off = rng.normal(size=3)
        off *= 8 / np.linalg.norm(off)
        assert oracle.separable(P, Qs + off)
        assert not oracle.separable(np.vstack([P, Qs.mean(0, keepdims=True)]), Qs)


def test_corridor_stage_invariants(oracle):
    cfg = cases.cfg_c1()
    cloud, cyl = cases.world_c1()
    m = oracle.map(cfg)
    m.set_pose([-4, 0, 1.0])
    m.update_gt(cloud, cyl, 0.0)
    start = np.zeros(9)
    start[:3] = [-4, 0, 1.0]
    ret, path, _ = m.astar(start, [4, 0, 1.0], 0.02)
    n, hp, rows, goal = m.corridor(path, start, 0.02)
    assert 1 <= n <= len(path) - 1
    for i in range(n):
        h = hp[i, :rows[i]]
        a = np.append(path[i, :3], 1.0)
        b = np.append(path[i + 1, :3], 1.0)
        assert (h @ a <= 1e-6).all() and (h @ b <= 1e-6).all()
    # local goal = start point of the last kept segment, inside the last corridor
    assert np.allclose(goal, path[n - 1])
    # and the QP on those corridors is solvable
    st, X, it = oracle.qp_solve(cfg, start, np.r_[goal, 0, 0, 0], [cfg.corridor_tau] * n, hp=hp, rows=rows,
                                vmax=cfg.opt_max_vel, amax=cfg.opt_max_acc)
    assert st in (1, 2)
