"""ctypes binding of oracle/_ref/libref.so — the REFERENCE's own raycast.cpp, sdlp.hpp, lbfgs.hpp and firi.hpp,
compiled unmodified from /root/reference against a stand-in for the absent Eigen (oracle/ref_build/Eigen/Eigen,
oracle/ref_glue.cpp, oracle/Makefile).  Test infrastructure: pins the oracle's restatement of those files.
The library exists only where it was built (this container; it travels to the GPU box as a built file);
`load()` returns None elsewhere and the tests fall back to the committed vectors in tests/golden/ref_pins.npz,
which scripts/make_golden_ref.py wrote from this library.  Also holds the seeded case generators that the
generator script and the tests share."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_ref", "libref.so")
REFERENCE = "/root/reference"
DP = C.POINTER(C.c_double)
D = C.c_double


def P(a):
    return a.ctypes.data_as(DP)


class RefLib:
    def __init__(self, lib):
        self.lib = lib
        lib.ref_linprog.restype = C.c_double
        lib.ref_cost_mvie.restype = C.c_double
        lib.ref_bezier_max_rate.restype = C.c_double

    def raycaster(self, start, end, cap=8192):
        out = np.zeros((cap, 3))
        ok = C.c_int(0)
        n = self.lib.ref_raycaster(P(start), P(end), P(out), cap, C.byref(ok))
        assert n <= cap
        return out[:n].copy(), ok.value

    def linprog(self, c, A, b):
        c = np.ascontiguousarray(c, np.float64)
        A = np.ascontiguousarray(A, np.float64).reshape(-1, len(c))
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros(len(c))
        return self.lib.ref_linprog(len(c), P(c), P(A), P(b), len(b), P(x)), x

    def cost_mvie(self, A, x, eps=1e-2, wt=1e3):
        g = np.zeros(9)
        return self.lib.ref_cost_mvie(P(A), len(A), D(eps), D(wt), P(x), P(g)), g

    def lbfgs_mvie(self, A, x0, eps=1e-2, wt=1e3):
        x = x0.copy()
        f = D(0)
        ret = self.lib.ref_lbfgs_mvie(P(A), len(A), D(eps), D(wt), P(x), C.byref(f))
        return ret, x, f.value

    def mvie(self, h, R, p, r):
        R, p, r = R.copy().ravel(), p.copy(), r.copy()
        ok = self.lib.ref_mvie(P(h), len(h), P(R), P(p), P(r))
        return ok, R.reshape(3, 3), p, r

    def firi(self, bd, pc, a, b, iterations, r0=None, cap=256):
        out = np.zeros((cap, 4))
        r = np.ones(3) if r0 is None else np.array(r0, np.float64)
        pc = np.ascontiguousarray(pc, np.float64).reshape(-1, 3)
        n = self.lib.ref_firi(P(bd), len(bd), P(pc), len(pc), P(a), P(b), int(iterations), P(r), P(out), cap)
        assert n <= cap
        return out[:n].copy(), r


    def bezier_eval(self, dur, cpts, t, deriv=0):
        out = np.zeros((len(t), 3))
        self.lib.ref_bezier_eval(P(dur), len(dur), P(cpts), P(t), len(t), int(deriv), P(out))
        return out

    def bezier_max_rate(self, dur, cpts, order):
        return self.lib.ref_bezier_max_rate(P(dur), len(dur), P(cpts), int(order))

    def qp_setup(self, init, goal, dur, polys, vmax, amax):
        """BezierOpt::setup -> (Q, A, lb, ub) as the reference assembles them."""
        M = len(dur)
        rows = np.array([len(h) for h in polys], np.int32)
        cap = int(rows.max())
        hp = np.zeros((M, cap, 4))
        for i, h in enumerate(polys):
            hp[i, :len(h)] = h
        i9, g9 = np.ascontiguousarray(init, np.float64).reshape(9), np.ascontiguousarray(goal, np.float64).reshape(9)
        dur = np.ascontiguousarray(dur, np.float64)
        args = (M, P(i9), P(g9), P(dur), P(hp), cap, rows.ctypes.data_as(C.POINTER(C.c_int32)), D(vmax), D(amax))
        m = self.lib.ref_qp_setup(*args, None, None, None, None)
        n = 15 * M
        Q, A, l, u = np.zeros((n, n)), np.zeros((m, n)), np.zeros(m), np.zeros(m)
        self.lib.ref_qp_setup(*args, P(Q), P(A), P(l), P(u))
        return Q, A, l, u


    def astar(self, cfg, omap, pose, start9, goal3, t_start, cap=256, risk=False):
        """The reference's FakeRiskHybridAstar (or, risk=True, RiskHybridAstar), driven as (Fake)BaselinePlanner::replan
        drives it, on the oracle map `omap` (occupancy queries go straight to liboracle's orc_map_query_dt)."""
        from tests import oracle_lib
        prm = np.array([cfg.max_tau, cfg.max_tau, cfg.max_vel, cfg.max_acc, cfg.w_time, cfg.horizon, cfg.resolution_astar,
                        cfg.time_resolution_astar, cfg.lambda_heu, cfg.allocate_num, cfg.check_num, cfg.tolerance,
                        cfg.corridor_tau], np.float64)
        q = C.cast(oracle_lib.load().lib.orc_map_query_dt, C.c_void_p)
        pose = np.ascontiguousarray(pose, np.float32)
        start9, goal3 = np.ascontiguousarray(start9, np.float64), np.ascontiguousarray(goal3, np.float64)
        out, n, st = np.zeros((cap, 6)), C.c_int(0), np.zeros(4, np.int32)
        ret = (self.lib.ref_astar_risk if risk else self.lib.ref_astar)(P(prm), q, omap.h, pose.ctypes.data_as(C.POINTER(C.c_float)), P(start9), P(goal3),
                                 D(t_start), P(out), cap, C.byref(n), st.ctypes.data_as(C.POINTER(C.c_int32)))
        assert n.value <= cap
        return ret, out[:n.value].copy(), st


    def replan(self, cfg, omap, pose, t_map, t_plan, start9, goal3, risk=False):
        """The reference's FakeBaselinePlanner::replan (risk=True: BaselinePlanner::replan) on the oracle map `omap` (queries and obstacle points answered by
        liboracle, the OSQP solve behind IOSQP by the oracle's ADMM): (ok, pieces, control points 5M x 3, stats =
        A* iterations, A* nodes, pieces, QP iterations)."""
        from tests import oracle_lib
        ol = oracle_lib.load().lib
        fns = [C.cast(f, C.c_void_p) for f in (ol.orc_map_query_dt, ol.orc_map_obstacle_points, ol.orc_osqp_dense)]
        pose = np.ascontiguousarray(pose, np.float32)
        start9, goal3 = np.ascontiguousarray(start9, np.float64), np.ascontiguousarray(goal3, np.float64)
        M, dur, cp, st = C.c_int(0), np.zeros(16), np.zeros((80, 3)), np.zeros(4, np.int32)
        ok = (self.lib.ref_replan_risk if risk else self.lib.ref_replan)(C.byref(cfg), *fns, omap.h, pose.ctypes.data_as(C.POINTER(C.c_float)), D(t_map), None, None,
                                 None, D(t_plan), D(t_plan), P(start9), P(goal3), C.byref(M), P(dur), P(cp),
                                 st.ctypes.data_as(C.POINTER(C.c_int32)))
        return ok, M.value, cp[:5 * M.value].copy(), st

    def traj_safe(self, cfg, omap, pose, t_map, rec, tau, now, T, risk=False):
        """The reference's (Fake)BaselinePlanner::isTrajSafe(T) for a planner holding the trajectory of record `rec`
        (pieces of tau seconds) over the oracle map `omap`, asked at wall-clock times `now`."""
        from tests import oracle_lib
        ol = oracle_lib.load().lib
        q = C.cast(ol.orc_map_query_dt, C.c_void_p)
        pose = np.ascontiguousarray(pose, np.float32)
        M = int(rec["n_pieces"])
        dur = np.full(M, tau, np.float64)
        cp = np.ascontiguousarray(rec["cpts"][:5 * M], np.float64)
        now = np.ascontiguousarray(np.atleast_1d(now), np.float64)
        T = np.ascontiguousarray(np.broadcast_to(np.asarray(T, np.float64), now.shape))
        out = np.zeros(len(now), np.int8)
        (self.lib.ref_traj_safe_risk if risk else self.lib.ref_traj_safe)(
            C.byref(cfg), q, omap.h, pose.ctypes.data_as(C.POINTER(C.c_float)), D(t_map), M, P(dur), P(cp),
            D(float(rec["start_time"])), len(now), P(now), P(T), out.ctypes.data_as(C.POINTER(C.c_int8)))
        return out

    def ego_particles(self, size):
        size = np.ascontiguousarray(size, np.float64)
        out = np.zeros((64, 3))
        n = self.lib.ref_ego_particles(P(size), P(out), 64)
        return out[:n].copy()

    def particles_with_risk(self, n, ego, size, offs, counts, recs, now, q_agent, q_t0):
        """ParticleATC of drone `ego` after receiving the peers' body particles and trajectories:
        getParticlesWithRisk for every (agent, t0) query -> (points nq x cap x 3, risks, counts; -1 = false)."""
        return peer_queries(self.lib.ref_particles_with_risk, (n, ego, P(np.ascontiguousarray(size, np.float64)), C.c_float(0.0),
                            P(offs), offs.shape[1], counts.ctypes.data_as(C.POINTER(C.c_int32)),
                            C.c_void_p(recs.ctypes.data), D(now)), offs.shape[1], q_agent, q_t0)


def sep_sets(fn, pre, post, cap=4096):
    """(number of separating-plane LPs, flat record |A|, |B|, B's points per LP) of one isSafeAfterOpt."""
    out = np.zeros(cap)
    n = fn(*pre, *post, P(out), cap)
    return n, out


def sep_case(seed=13, M=5):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(rng.normal(size=(5 * M, 3))), np.full(M, 0.3), M, (9.0, 10.2, 10.95, 11.5, 12.5)


def ref_sep_sets(ref, n, ego, size, offs, counts, recs, now, cpts, dur, M):
    return sep_sets(ref.lib.ref_safe_after_opt_sets,
                    (n, ego, P(np.ascontiguousarray(size, np.float64)), P(offs), offs.shape[1],
                     counts.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(recs.ctypes.data), D(now)),
                    (P(cpts), P(dur), M))


def peer_queries(fn, pre, cap, q_agent, q_t0):
    nq = len(q_agent)
    pts, rk, on = np.zeros((nq, cap, 3)), np.zeros((nq, cap), np.float32), np.zeros(nq, np.int32)
    fn(*pre, nq, q_agent.ctypes.data_as(C.POINTER(C.c_int32)), P(q_t0), P(pts), rk.ctypes.data_as(C.POINTER(C.c_float)), cap,
       on.ctypes.data_as(C.POINTER(C.c_int32)))
    return pts, rk, on


def peer_case(seed=3, n=8, ego=2, nq=300):
    """A swarm of n drones on straight Bezier trajectories from t = 10 s: one has not published, one starts later;
    queries (agent, time) before, during and after the trajectories, the ego drone included."""
    import pred_occ_planner_b200 as p
    from pred_occ_planner_b200 import worldgen
    from tests import cases
    cfg = cases.cfg_c1(p.POP_CELL_F32, n_local=1, n_total=n)
    size = np.array(list(cfg.drone_size), np.float64)
    rng = np.random.default_rng(seed)
    trajs = worldgen.straight_trajs(rng.uniform(-4, 4, (n, 3)), rng.uniform(-4, 4, (n, 3)), 10.0, 1.5, 6, 0.3)
    trajs["n_pieces"][5] = 0
    trajs["start_time"][6] = 10.9
    offs, counts = worldgen.body_table(n, size, ego_id=ego)
    q_agent = rng.integers(0, n, nq).astype(np.int32)
    q_t0 = rng.uniform(9.5, 13.0, nq)
    return cfg, size, ego, offs, counts, np.ascontiguousarray(trajs), q_agent, q_t0


_ref = False


def load():
    """The reference library, built on demand where /root/reference exists; None where it cannot exist."""
    global _ref
    if _ref is not False:
        return _ref
    if os.environ.get("POP_NO_REF"):     # exercise the golden-only path where the library does exist
        _ref = None
        return _ref
    if os.path.isdir(REFERENCE):
        subprocess.run(["make", "-C", ORACLE_DIR, "-s", "_ref/libref.so"], check=True)
    _ref = RefLib(C.CDLL(LIB)) if os.path.exists(LIB) else None
    return _ref


# oracle-side hooks with the same signatures -----------------------------------------------------------------
class OrcHooks:
    def __init__(self, oracle):
        self.o = oracle
        self.lib = oracle.lib
        self.lib.orc_cost_mvie.restype = C.c_double

    def raycaster(self, start, end, cap=8192):
        out = np.zeros((cap, 3))
        ok = C.c_int(0)
        n = self.lib.orc_raycaster(P(start), P(end), P(out), cap, C.byref(ok))
        assert n <= cap
        return out[:n].copy(), ok.value

    def linprog(self, c, A, b, seed=1):
        return self.o.linprog(c, A, b, seed=seed)

    def cost_mvie(self, A, x, eps=1e-2, wt=1e3):
        g = np.zeros(9)
        return self.lib.orc_cost_mvie(P(A), len(A), D(eps), D(wt), P(x), P(g)), g

    def lbfgs_mvie(self, A, x0, eps=1e-2, wt=1e3):
        x = x0.copy()
        f = D(0)
        ret = self.lib.orc_lbfgs_mvie(P(A), len(A), D(eps), D(wt), P(x), C.byref(f))
        return ret, x, f.value

    def mvie(self, h, R, p, r, seed=1):
        R, p, r = R.copy().ravel(), p.copy(), r.copy()
        ok = self.lib.orc_mvie(P(h), len(h), P(R), P(p), P(r), C.c_uint64(seed))
        return ok, R.reshape(3, 3), p, r

    def firi(self, bd, pc, a, b, iterations, r0=None, seed=1, cap=256):
        out = np.zeros((cap, 4))
        r = np.ones(3) if r0 is None else np.array(r0, np.float64)
        pc = np.ascontiguousarray(pc, np.float64).reshape(-1, 3)
        n = self.lib.orc_firi_r(P(bd), P(pc), len(pc), P(a), P(b), int(iterations), C.c_uint64(seed), P(r), P(out), cap)
        assert n <= cap
        return out[:n].copy(), r


# seeded cases ----------------------------------------------------------------------------------------------------
def ray_cases(n, seed=0):
    """(start, end) in voxel units as raycastProcess passes them: generic, on voxel boundaries, axis-aligned,
    sub-voxel."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        s = rng.uniform(-60, 60, 3)
        e = rng.uniform(-60, 60, 3)
        if i % 5 == 0:
            s = np.round(s * 2) / 2
        if i % 7 == 0:
            e[rng.integers(3)] = s[rng.integers(3)]
        if i % 11 == 0:
            e = s.copy()
            e[rng.integers(3)] += rng.uniform(-0.4, 0.4)
        if i % 13 == 0:
            e = np.round(e)
        out.append((s, e))
    return out


def lp_cases(n, seed=1):
    """(c, A, b): bounded polytopes around the origin, boxes, duplicated rows, and random (mostly unbounded or
    infeasible) systems, d = 2..4."""
    rng = np.random.default_rng(seed)
    out = []
    for it in range(n):
        d = int(rng.integers(2, 5))
        m = int(rng.integers(d + 1, 60))
        A = rng.normal(size=(m, d))
        kind = it % 4
        if kind in (0, 3):
            A /= np.linalg.norm(A, axis=1, keepdims=True)
            b = rng.uniform(0.5, 2, m)
            A = np.vstack([A, np.eye(d), -np.eye(d)])
            b = np.concatenate([b, np.full(2 * d, 5.0)])
            if kind == 3:
                A = np.vstack([A, A[:3]])
                b = np.concatenate([b, b[:3]])
        elif kind == 1:
            b = rng.normal(size=m)
        else:
            A = np.vstack([np.eye(d), -np.eye(d)])
            b = np.concatenate([rng.uniform(0, 1, d), rng.uniform(0, 1, d)])
        out.append((rng.normal(size=d), np.ascontiguousarray(A), np.ascontiguousarray(b)))
    return out


def mvie_cost_cases(n, seed=3):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        M = int(rng.integers(4, 40))
        A = np.ascontiguousarray(rng.normal(size=(M, 3)) * rng.uniform(0.2, 3))
        x = np.concatenate([rng.normal(size=3) * 0.2, rng.uniform(0.3, 1.5, 3), rng.normal(size=3) * 0.3])
        out.append((A, x))
    return out


def lbfgs_cases(n, seed=4):
    """Faces a.x <= 1 around the origin (what maxVolInsEllipsoid hands to L-BFGS) and a start inside."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        M = int(rng.integers(6, 30))
        nrm = rng.normal(size=(M, 3))
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        A = np.vstack([nrm / rng.uniform(0.5, 3, (M, 1)), np.eye(3) / 3, -np.eye(3) / 3])
        x0 = np.array([0, 0, 0, 0.3, 0.3, 0.3, 0, 0, 0.0])
        x0[:3] += rng.normal(size=3) * 0.02
        out.append((np.ascontiguousarray(A), x0))
    return out


def poly_cases(n, seed=5):
    """H-polytopes (rows h0 x + h1 y + h2 z + h3 <= 0): a box around the origin cut by random planes."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        lo, hi = -rng.uniform(1, 3, 3), rng.uniform(1, 3, 3)
        rows = []
        for i in range(3):
            r = np.zeros(4); r[i] = 1; r[3] = -hi[i]; rows.append(r)
            r = np.zeros(4); r[i] = -1; r[3] = lo[i]; rows.append(r)
        for _j in range(int(rng.integers(0, 14))):
            nv = rng.normal(size=3)
            nv /= np.linalg.norm(nv)
            rows.append(np.array([*nv, -rng.uniform(0.4, 2.0)]) * rng.uniform(0.5, 2))
        out.append(np.ascontiguousarray(np.array(rows)))
    return out


def firi_cases(n, seed=2):
    """(bd, pc, a, b, r0): a line segment in a box with uniformly scattered obstacle points kept 0.25 m off the
    segment; r0 = firi's in/out initial radii (unit for two cases in three, anisotropic otherwise)."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        a = rng.uniform(-2, 2, 3)
        b = a + rng.uniform(-1.5, 1.5, 3)
        lo = np.minimum(a, b) - rng.uniform(0.5, 2, 3)
        hi = np.maximum(a, b) + rng.uniform(0.5, 2, 3)
        bd = np.zeros((6, 4))
        for i in range(3):
            bd[2 * i, i] = 1; bd[2 * i, 3] = -hi[i]
            bd[2 * i + 1, i] = -1; bd[2 * i + 1, 3] = lo[i]
        pc = rng.uniform(lo, hi, (int(rng.integers(0, 120)), 3))
        ab = b - a
        t = np.clip(((pc - a) @ ab) / max(ab @ ab, 1e-12), 0, 1)
        pc = np.ascontiguousarray(pc[np.linalg.norm(pc - (a + t[:, None] * ab), axis=1) > 0.25])
        r0 = np.ones(3) if k % 3 else rng.uniform(0.4, 2.5, 3)
        out.append((bd, pc, a, b, r0, lo, hi))
    return out


def bezier_cases(n, seed=6):
    """(durations, control points 5M x 3, sample times incl. 0, the piece joints, the end and past the end)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        M = int(rng.integers(1, 9))
        dur = rng.uniform(0.2, 2.5, M)
        cp = np.ascontiguousarray(rng.normal(size=(5 * M, 3)) * 3)
        T = dur.sum()
        t = np.concatenate([rng.uniform(0, T, 20), [0.0, T, T + 0.3], np.cumsum(dur)])
        out.append((dur, cp, t))
    return out


def qp_cases(n, seed=7, max_pieces=8):
    """(init, goal [pos; vel; acc], durations, one H-polytope per piece, vmax, amax)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        M = int(rng.integers(1, max_pieces + 1))
        polys = [np.c_[rng.normal(size=(r, 3)), -rng.uniform(0.5, 3, r)] for r in rng.integers(4, 20, M)]
        out.append((rng.normal(size=(3, 3)), rng.normal(size=(3, 3)), rng.uniform(0.3, 2.5, M), polys,
                    float(rng.uniform(1, 4)), float(rng.uniform(1, 5))))
    return out


def astar_world():
    """BASELINE config 1 grid (66 x 66 x 20 @0.15 m) with 12 static pillars and 8 moving cylinders."""
    from tests import cases
    return cases.cfg_c1(), cases.world_moving(77, (10.0, 10.0, 3.0), 12, 8)


def astar_world_risk():
    """The same world queried the RiskBase way (risk_base.cpp:228-260: ground / ceiling count as occupied, threshold
    risk_threshold_region - t * decay), thresholds chosen so that one occupied cell blocks at every slice."""
    import pred_occ_planner_b200 as p
    cfg, world = astar_world()
    cfg.query_variant = p.POP_QUERY_RISKBASE
    cfg.risk_threshold_region, cfg.risk_region_decay = 0.9, 0.05
    cfg.risk_voxel_decay = 0.01          # obstacle-point threshold risk_threshold - decay * slice stays positive
    return cfg, world


def astar_cases(n, seed=8):
    """(pose, start [pos vel acc], goal, time of the start on the map): agents scattered over the arena, goals up
    to 6 m away (some too close to search: NO_PATH after the retry)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        pose = np.array([rng.uniform(-4, 4), rng.uniform(-3, 3), 1.0], np.float32)
        start = np.zeros(9)
        start[:3] = pose.astype(np.float64) + rng.uniform(-0.3, 0.3, 3) * [1, 1, 0.2]
        start[3:6] = rng.uniform(-1, 1, 3) * [1, 1, 0.2]
        start[6:9] = rng.uniform(-1, 1, 3) * [1, 1, 0.1]
        goal = start[:3] + rng.uniform(-6, 6, 3) * [1, 1, 0.05]
        out.append((pose, start, goal, float(rng.uniform(0, 0.25))))
    return out


def ellipsoid(R, p, r, h):
    """(volume factor, Q = R diag(r^2) R', worst normalised face violation) of an ellipsoid inside polytope h."""
    E = R @ np.diag(r)
    hn = np.linalg.norm(h[:, :3], axis=1)
    viol = ((np.linalg.norm(h[:, :3] @ E, axis=1) + h[:, :3] @ p + h[:, 3]) / hn).max()
    return float(np.prod(r)), E @ E.T, float(viol)


def poly_iou(h1, h2, lo, hi, n=20000, seed=9):
    S = np.random.default_rng(seed).uniform(lo, hi, (n, 3))
    S1 = np.c_[S, np.ones(n)]
    in1, in2 = (S1 @ h1.T <= 0).all(1), (S1 @ h2.T <= 0).all(1)
    return (in1 & in2).sum() / max(1, (in1 | in2).sum())


# ragged arrays in an .npz ---------------------------------------------------------------------------------------
def pack(arrs):
    arrs = [np.asarray(a) for a in arrs]
    flat = np.concatenate([a.ravel() for a in arrs]) if arrs else np.zeros(0)
    return flat, np.cumsum([0] + [a.size for a in arrs]).astype(np.int64)


def unpack(flat, offs, cols=None):
    out = [flat[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    return [a.reshape(-1, cols) for a in out] if cols else out


def traj_safe_cases(orc, n=14, seed=67):
    """isTrajSafe problems: a trajectory planned by the oracle on the map of t = 100, then asked about on maps taken
    later from a drifted pose (so that some samples have become occupied): (cfg, map, pose, t_map, record, now[], T[])."""
    from tests import oracle_cycle
    cfg, (cloud, cyl) = astar_world()
    rng = np.random.default_rng(seed + 1)
    out = []
    for pose, start, goal, ts in astar_cases(n, seed=seed):
        m = orc.map(cfg)
        m.set_pose(pose)
        m.update_gt(cloud, cyl, 100.0)
        st, rec, _ = oracle_cycle.replan(m, None, 0, start, goal, 100.0 + ts)
        if st != 0:
            continue
        for t_map in (100.0, 100.6, 101.1):
            p2 = (pose + np.float32([0.3, 0.2, 0]) * np.float32(t_map - 100.0)).astype(np.float32)
            m2 = orc.map(cfg)
            m2.set_pose(p2)
            m2.update_gt(cloud, cyl, t_map)
            now = float(rec["start_time"]) + rng.uniform(-0.3, 2.0, 6)
            T = rng.choice([0.5, 1.0, 2.0, 5.0], 6)
            out.append((cfg, m2, p2, t_map, rec, now, T))
    return out
