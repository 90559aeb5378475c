"""GPU parity of the Bezier QP (K4) against the oracle's OSQP restatement: same status, same ADMM
iteration count, control points within 1e-4 relative (north_star tolerance for trajectory
coefficients).  Includes the reference's own 3-cube fixture (traj_opt/test/test_bezier_opt.cpp:47-106)."""
import numpy as np
import pytest

import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import POP_MAX_FACES, POP_MAX_PIECES
from tests import cases
from tests.test_oracle_corridor_qp import FIX_END, FIX_POLYS, FIX_START, FIX_T

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _close(got, ref):
    scale = max(1.0, np.abs(ref).max())
    return np.abs(got - ref).max() <= RTOL * scale


def test_reference_fixture(oracle):
    cfg = cases.cfg_c1(p.POP_CELL_U8)
    eng = p.PopEngine(cfg)
    hp, rows = oracle.pack_polys(FIX_POLYS)
    dur = np.zeros(POP_MAX_PIECES)
    dur[:3] = FIX_T
    cpts, status, iters = eng.qp_solve_batch([3], [FIX_START], [FIX_END], [dur], [hp], [rows], 3.0, 3.0)
    st, X, it = oracle.qp_solve(cfg, FIX_START, FIX_END, FIX_T, FIX_POLYS)
    assert status[0] == st and st in (1, 2)   # optimize() == true (bezier_optimizer.cpp:279)
    assert iters[0] == it
    assert _close(cpts[0, :15], X)
    # the reference's assertions on the solution (boundary conditions within 1e-3)
    got = cpts[0, :15]
    assert np.abs(oracle.bezier_eval(FIX_T, got, [0.0])[0] - FIX_START[:3]).max() < 1e-3
    assert np.abs(oracle.bezier_eval(FIX_T, got, [8.0])[0] - FIX_END[:3]).max() < 1e-3
    assert np.abs(oracle.bezier_eval(FIX_T, got, [8.0], 1)[0] - FIX_END[3:6]).max() < 1e-3


def test_pipeline_problems(oracle):
    """QPs produced by the real pipeline (A* path → corridors) for many agents, mixed piece counts."""
    n = 24
    cfg = cases.cfg_c1(p.POP_CELL_U8, n_local=n, n_total=n)
    cloud, cyl = cases.world_c1(seed=99)
    rng = np.random.default_rng(8)
    start = np.zeros((n, 9))
    start[:, :2] = rng.uniform(-4.5, 4.5, (n, 2))
    start[:, 2] = rng.uniform(0.7, 1.5, n)
    start[:, 3:5] = rng.uniform(-0.8, 0.8, (n, 2))
    goal = np.concatenate([rng.uniform(-4.5, 4.5, (n, 2)), rng.uniform(0.8, 1.5, (n, 1))], axis=1)
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    eng.set_poses(start[:, :3])
    eng.map_update_gt(2.0)
    ret, path, ln, _ = eng.astar_search_batch(start, goal, np.full(n, 0.02))
    hp, rows, npoly, lgoal = eng.corridor_batch(path, ln, start, np.full(n, 2.02))
    fin = np.zeros((n, 9))
    fin[:, :6] = lgoal
    dur = np.zeros((n, POP_MAX_PIECES))
    for a in range(n):
        dur[a, :npoly[a]] = cfg.corridor_tau
    cpts, status, iters = eng.qp_solve_batch(npoly, start, fin, dur, hp, rows, cfg.opt_max_vel, cfg.opt_max_acc)
    solved = 0
    bad = []
    for a in range(n):
        M = int(npoly[a])
        if M == 0:
            assert status[a] == 0
            continue
        st, X, it = oracle.qp_solve(cfg, start[a], fin[a], dur[a, :M], hp=hp[a], rows=rows[a],
                                    vmax=cfg.opt_max_vel, amax=cfg.opt_max_acc)
        ok = st == status[a] and it == iters[a] and (st not in (1, 2) or _close(cpts[a, :5 * M], X))
        if not ok:
            bad.append((a, M, int(st), int(status[a]), it, int(iters[a]),
                        float(np.abs(cpts[a, :5 * M] - X).max())))
        solved += st in (1, 2)
    assert not bad, bad
    assert solved >= 5


def test_infeasible_and_degenerate(oracle):
    cfg = cases.cfg_c1(p.POP_CELL_U8)
    eng = p.PopEngine(cfg)

    def cube(lo, hi):
        return [[-1, 0, 0, lo], [1, 0, 0, -hi], [0, -1, 0, lo], [0, 1, 0, -hi], [0, 0, -1, lo], [0, 0, 1, -hi]]

    problems = [
        # start outside its corridor → primal infeasible (or max-iter): must agree with the oracle
        ([cube(0, 1), cube(0.5, 2)], [3, 3, 3, 0, 0, 0, 0, 0, 0], [1.5, 1.5, 1.5, 0, 0, 0, 0, 0, 0], [0.3, 0.3]),
        # empty polytope (firi returned false in the reference) → no safety rows for that piece
        ([[], cube(-1, 2)], [0, 0, 0, 0.2, 0, 0, 0, 0, 0], [0.5, 0.1, 0.0, 0.2, 0, 0, 0, 0, 0], [0.3, 0.3]),
        # single piece
        ([cube(-1, 1)], [0, 0, 0, 0, 0, 0, 0, 0, 0], [0.3, 0.2, 0.1, 0, 0, 0, 0, 0, 0], [0.5]),
    ]
    for polys, s, e, t in problems:
        hp, rows = oracle.pack_polys(polys)
        dur = np.zeros(POP_MAX_PIECES)
        dur[:len(t)] = t
        cpts, status, iters = eng.qp_solve_batch([len(t)], [s], [e], [dur], [hp], [rows], 2.0, 6.0)
        st, X, it = oracle.qp_solve(cfg, s, e, t, polys, vmax=2.0, amax=6.0)
        assert status[0] == st, (status[0], st)
        assert iters[0] == it
        if st in (1, 2):
            assert _close(cpts[0, :5 * len(t)], X)


def _tube_problem(M, faces_per_piece, rng):
    """A straight tube of overlapping boxes along x with extra (redundant) random faces, so that the
    row count can be pushed past the register-tile kernel's limits."""
    polys = []
    step = 0.3
    for i in range(M):
        lo = np.array([i * step - 0.5, -1.0, 0.0])
        hi = np.array([(i + 1) * step + 0.5, 1.0, 2.0])
        h = [[-1, 0, 0, lo[0]], [1, 0, 0, -hi[0]], [0, -1, 0, lo[1]], [0, 1, 0, -hi[1]], [0, 0, -1, lo[2]], [0, 0, 1, -hi[2]]]
        c = (lo + hi) / 2
        while len(h) < faces_per_piece:
            nrm = rng.normal(size=3)
            nrm /= np.linalg.norm(nrm)
            d = 2.5 + rng.uniform(0, 1)            # far outside the box: never active
            h.append([nrm[0], nrm[1], nrm[2], -(nrm @ c) - d])
        polys.append(h)
    start = [0.0, 0, 1, 0.5, 0, 0, 0, 0, 0]
    end = [M * step, 0.2, 1.1, 0.3, 0, 0, 0, 0, 0]
    return polys, start, end, [0.3] * M


@pytest.mark.parametrize("M,faces", [(12, 8), (11, 20), (10, 38), (8, 48)])
def test_large_problems_take_the_shared_memory_variant(oracle, M, faces):
    """12- and 11-piece problems (n = 180, 165 > 150) and problems with more than 3*384 rows (up to 2 304) go to the
    kernel variant that streams the packed inverse factor from shared memory: same parity bar."""
    rng = np.random.default_rng(M * 100 + faces)
    polys, s, e, t = _tube_problem(M, faces, rng)
    cfg = cases.cfg_c1(p.POP_CELL_U8)
    eng = p.PopEngine(cfg)
    hp, rows = oracle.pack_polys(polys)
    dur = np.zeros(POP_MAX_PIECES)
    dur[:M] = t
    cpts, status, iters = eng.qp_solve_batch([M], [s], [e], [dur], [hp], [rows], 2.0, 6.0)
    st, X, it = oracle.qp_solve(cfg, s, e, t, polys, vmax=2.0, amax=6.0)
    assert status[0] == st, (status[0], st)
    assert iters[0] == it
    assert st in (1, 2)
    assert _close(cpts[0, :5 * M], X)


def test_row_capacity_is_reported(oracle):
    """More rows than any kernel variant holds → POP_E_CAPACITY, not a wrong answer."""
    rng = np.random.default_rng(1)
    polys, s, e, t = _tube_problem(12, 48, rng)    # 369 + 5*576 rows
    cfg = cases.cfg_c1(p.POP_CELL_U8)
    eng = p.PopEngine(cfg)
    hp, rows = oracle.pack_polys(polys)
    dur = np.zeros(POP_MAX_PIECES)
    dur[:12] = t
    with pytest.raises(p.PopError) as err:
        eng.qp_solve_batch([12], [s], [e], [dur], [hp], [rows], 2.0, 6.0)
    assert err.value.code == -4
