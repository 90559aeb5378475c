"""The device's Seidel LP (csrc/pop_lp.cuh) on its own, through `pop_lp_solve`: the warp-cooperative implementation the
planner kernels use against the single-thread one (the source the host-compiled tests share) and against the oracle,
bit for bit — value, minimiser, and the infeasible / unbounded classification — on bounded, infeasible, unbounded and
degenerate problems in dimensions 2-4, from a handful of rows to more than a warp's worth several times over."""
import numpy as np
import pytest

import pred_occ_planner_b200 as p
from tests import cases

pytestmark = pytest.mark.gpu


def _problems(rng):
    for trial in range(150):
        d = int(rng.integers(2, 5))
        kind = trial % 5
        m = int(rng.integers(d + 1, 160))
        A = rng.normal(size=(m, d))
        if kind == 0:      # a bounded polytope around a random interior point
            A /= np.linalg.norm(A, axis=1, keepdims=True)
            x0 = rng.normal(size=d)
            b = A @ x0 + rng.uniform(0.05, 2.0, m)
            A = np.vstack([A, np.eye(d), -np.eye(d)])
            b = np.concatenate([b, x0 + 5.0, -(x0 - 5.0)])
        elif kind == 1:    # infeasible: two opposite half-spaces that exclude each other, among others
            b = rng.uniform(0.5, 2.0, m)
            A[1] = -A[0]
            b[0], b[1] = -1.0, -1.0
        elif kind == 2:    # unbounded: every x far enough along -1 is feasible and the objective falls that way
            A = np.abs(A)
            b = rng.uniform(0.5, 2.0, m)
        elif kind == 3:    # degenerate: duplicated and parallel rows, many active at the optimum (a box corner)
            A = np.vstack([np.eye(d), np.eye(d), -np.eye(d), A[: m // 4] * 0.0 + np.eye(d)[rng.integers(0, d, m // 4)]])
            b = np.concatenate([np.ones(d), np.ones(d), np.ones(d), np.ones(m // 4)])
        else:              # feasibility problem (zero objective), the planner's most frequent use
            A /= np.linalg.norm(A, axis=1, keepdims=True)
            b = rng.uniform(-0.2, 1.0, len(A))
        c = np.zeros(d) if kind == 4 else rng.normal(size=d)
        if kind == 2:
            c = np.abs(c) + 0.1
        yield c, np.ascontiguousarray(A), np.ascontiguousarray(b), int(rng.integers(1, 1 << 60))


def test_warp_lp_equals_single_thread_lp_and_oracle(oracle):
    eng = p.PopEngine(cases.cfg_small_c2(p.POP_CELL_U8, n_local=1, n_total=1))
    rng = np.random.default_rng(17)
    seen = {"bounded": 0, "infeasible": 0, "unbounded": 0}
    for c, A, b, seed in _problems(rng):
        vw, xw = eng.lp_solve(c, A, b, seed, cooperative=True)
        vs, xs = eng.lp_solve(c, A, b, seed, cooperative=False)
        vo, xo = oracle.linprog(c, A, b, seed)
        for v, x in ((vs, xs), (vo, xo)):
            assert (np.isinf(vw) and np.isinf(v) and np.sign(vw) == np.sign(v)) or vw == v, (vw, v, len(b), len(c))
            assert np.array_equal(xw, x), (xw, x)
        seen["infeasible" if vw == np.inf else "unbounded" if vw == -np.inf else "bounded"] += 1
    assert min(seen.values()) >= 15, seen
    eng.close()
