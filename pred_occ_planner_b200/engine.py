"""Host side of the drop-in: Python mirror of the reference's map / planner interfaces over
the C ABI of libpop.so.

`PopEngine` owns one `pop_handle` (= N local agents on one GPU) and stands where N
`FakeBaselinePlanner` objects stood (plan_manager/src/baseline_fake.cpp:18-51).
`AgentMap` keeps the method names and return conventions of `FakeParticleRiskVoxel` /
`MapBase` (plan_env/include/plan_env/fake_particle_risk_voxel.h:49-91,
plan_env/include/plan_env/map.h:93-146) so parity tests read like reference call sites.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (POP_MAX_BODY, POP_MAX_FACES, POP_MAX_PATH, POP_MAX_PIECES, POP_STAGE_COUNT,
                   STAGE_NAMES, TRAJ_DTYPE, PopConfig, PopCylinder, PopError, PopTrajRecord, as_ptr)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class PopEngine:
    def __init__(self, cfg: PopConfig, device: int = 0, stream: int | None = None):
        self.lib = _lib.load()
        self.cfg = cfg.copy()
        self._h = C.c_void_p()
        rc = self.lib.pop_create(C.byref(self.cfg), device, C.c_void_p(stream or 0), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise PopError(rc, "pop_create failed (no CUDA device → NO_DEVICE; there is no CPU fallback)")
        self.n_local = cfg.n_agents_local
        self.n_total = cfg.n_agents_total
        self.device = int(device)

    # -- plumbing ----------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise PopError(rc, self.lib.pop_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.pop_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.pop_sync(self._h))

    # -- inputs --------------------------------------------------------------------------------
    def set_world_cloud(self, pts: np.ndarray):
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
        self._check(self.lib.pop_set_world_cloud(self._h, as_ptr(pts, C.c_float), len(pts)))

    def set_cylinders(self, cyl: np.ndarray):
        cyl = np.ascontiguousarray(cyl, dtype=_lib.CYL_DTYPE)
        self._check(self.lib.pop_set_cylinders(self._h, cyl.ctypes.data_as(C.POINTER(PopCylinder)), len(cyl)))

    def set_world_cloud_dev(self, ptr: int, n: int):
        """world cloud already in device memory (n x 3 float32 at device address ptr); async on the engine's stream"""
        self._check(self.lib.pop_set_world_cloud_dev(self._h, C.c_void_p(ptr), int(n)))

    def set_cylinders_dev(self, ptr: int, n: int, max_abs_vy: float):
        """obstacle states already in device memory (n pop_cylinder records at device address ptr); async"""
        self._check(self.lib.pop_set_cylinders_dev(self._h, C.c_void_p(ptr), int(n), float(max_abs_vy)))

    def set_resample_noise(self, z: np.ndarray):
        """standard-normal triples of the peer resampling (replan_risk_rate > 0), see pop_set_resample_noise"""
        z = np.ascontiguousarray(z, np.float32).reshape(-1, 3)
        self._check(self.lib.pop_set_resample_noise(self._h, as_ptr(z, C.c_float), len(z)))

    def get_trajectories(self) -> np.ndarray:
        """the peers' table as the next map update will read it (n_total records)"""
        recs = np.zeros(self.n_total, TRAJ_DTYPE)
        self._check(self.lib.pop_get_trajectories(self._h, recs.ctypes.data_as(C.POINTER(PopTrajRecord))))
        return recs

    def set_poses(self, poses: np.ndarray):
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(self.n_local, 3)
        self._check(self.lib.pop_set_poses(self._h, as_ptr(poses, C.c_float)))
        self._poses = poses.copy()

    def set_body_particles(self, offs: np.ndarray, counts: np.ndarray):
        offs = _f64(offs, (self.n_total, POP_MAX_BODY, 3))
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        self._check(self.lib.pop_set_body_particles(self._h, as_ptr(offs, C.c_double), as_ptr(counts, C.c_int32)))

    def set_trajectories(self, recs: np.ndarray):
        recs = np.ascontiguousarray(recs, dtype=TRAJ_DTYPE)
        assert len(recs) == self.n_total
        self._check(self.lib.pop_set_trajectories(self._h, recs.ctypes.data_as(C.POINTER(PopTrajRecord))))

    # -- map -----------------------------------------------------------------------------------
    def map_update_gt(self, t_map: float):
        """FakeParticleRiskVoxel::updateMap for every local agent (K1f + K5)."""
        self._check(self.lib.pop_map_update_gt(self._h, float(t_map)))
        self._t_map = float(t_map)

    def map_future_risk(self, risk: np.ndarray, stamp: float, poses: np.ndarray | None = None):
        """RiskBase::futureRiskCallback for every local agent: risk = n_local x V x T floats as published on
        "map/future_risk"; then the peers are projected (risk_base.cpp:60-80)."""
        c = self.cfg
        risk = np.ascontiguousarray(risk, np.float32).reshape(self.n_local, c.nx * c.ny * c.nz, c.nt)
        pp = None
        if poses is not None:
            poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(self.n_local, 3)
            self._poses = poses.copy()
            pp = as_ptr(poses, C.c_float)
        self._check(self.lib.pop_map_future_risk(self._h, as_ptr(risk, C.c_float), pp, float(stamp)))
        self._t_map = float(stamp)

    def is_traj_safe(self, t_now: float, T: float = 2.0):
        """BaselinePlanner::isTrajSafe(T) for every local agent at wall-clock time t_now:
        (safe flags int8 [n_local], relative time of the first occupied sample or -1)."""
        safe = np.zeros(self.n_local, np.int8)
        tu = np.zeros(self.n_local)
        self._check(self.lib.pop_traj_safe_batch(self._h, float(t_now), float(T), as_ptr(safe, C.c_int8), as_ptr(tu, C.c_double)))
        return safe, tu

    def agent_map(self, agent: int) -> "AgentMap":
        return AgentMap(self, agent)

    # -- planning stages -----------------------------------------------------------------------
    def astar_search_batch(self, start, goal, t_start):
        n = self.n_local
        start, goal, t_start = _f64(start, (n, 9)), _f64(goal, (n, 3)), _f64(t_start, (n,))
        ret = np.zeros(n, np.int32)
        path = np.zeros((n, POP_MAX_PATH, 6))
        ln = np.zeros(n, np.int32)
        stats = np.zeros((n, 4), np.int32)
        self._check(self.lib.pop_astar_search_batch(self._h, as_ptr(start, C.c_double), as_ptr(goal, C.c_double),
                                                    as_ptr(t_start, C.c_double), as_ptr(ret, C.c_int32),
                                                    as_ptr(path, C.c_double), as_ptr(ln, C.c_int32),
                                                    as_ptr(stats, C.c_int32)))
        return ret, path, ln, stats

    def corridor_batch(self, path, ln, start, traj_start_time):
        n = self.n_local
        path, start = _f64(path, (n, POP_MAX_PATH, 6)), _f64(start, (n, 9))
        ln = np.ascontiguousarray(ln, np.int32)
        tst = _f64(traj_start_time, (n,))
        hpoly = np.zeros((n, POP_MAX_PIECES, POP_MAX_FACES, 4))
        rows = np.zeros((n, POP_MAX_PIECES), np.int32)
        npoly = np.zeros(n, np.int32)
        goal = np.zeros((n, 6))
        self._check(self.lib.pop_corridor_batch(self._h, as_ptr(path, C.c_double), as_ptr(ln, C.c_int32),
                                                as_ptr(start, C.c_double), as_ptr(tst, C.c_double),
                                                as_ptr(hpoly, C.c_double), as_ptr(rows, C.c_int32),
                                                as_ptr(npoly, C.c_int32), as_ptr(goal, C.c_double)))
        return hpoly, rows, npoly, goal

    def qp_solve_batch(self, n_pieces, init, goal, dur, hpoly, rows, vmax, amax):
        n_pieces = np.ascontiguousarray(n_pieces, np.int32)
        n = len(n_pieces)
        init, goal = _f64(init, (n, 9)), _f64(goal, (n, 9))
        dur = _f64(dur, (n, POP_MAX_PIECES))
        hpoly = _f64(hpoly, (n, POP_MAX_PIECES, POP_MAX_FACES, 4))
        rows = np.ascontiguousarray(rows, np.int32).reshape(n, POP_MAX_PIECES)
        cpts = np.zeros((n, POP_MAX_PIECES * 5, 3))
        status = np.zeros(n, np.int32)
        iters = np.zeros(n, np.int32)
        self._check(self.lib.pop_qp_solve_batch(self._h, n, as_ptr(n_pieces, C.c_int32), as_ptr(init, C.c_double),
                                                as_ptr(goal, C.c_double), as_ptr(dur, C.c_double),
                                                as_ptr(hpoly, C.c_double), as_ptr(rows, C.c_int32),
                                                float(vmax), float(amax), as_ptr(cpts, C.c_double),
                                                as_ptr(status, C.c_int32), as_ptr(iters, C.c_int32)))
        return cpts, status, iters

    def separation_check_batch(self, n_pieces, cpts, t_now):
        n = self.n_local
        n_pieces = np.ascontiguousarray(n_pieces, np.int32)
        cpts = _f64(cpts, (n, POP_MAX_PIECES * 5, 3))
        safe = np.zeros(n, np.int8)
        self._check(self.lib.pop_separation_check_batch(self._h, as_ptr(n_pieces, C.c_int32), as_ptr(cpts, C.c_double),
                                                        float(t_now), as_ptr(safe, C.c_int8)))
        return safe

    def replan_cycle(self, start, goal, poses, t_map, t_plan):
        """One full cycle with HOST inputs/outputs (map update → A* → corridor → QP → separation)."""
        n = self.n_local
        start, goal = _f64(start, (n, 9)), _f64(goal, (n, 3))
        poses = np.ascontiguousarray(poses, np.float32).reshape(n, 3)
        status = np.zeros(n, np.int32)
        traj = np.zeros(n, TRAJ_DTYPE)
        io = _lib.PopCycleIO(as_ptr(start, C.c_double), as_ptr(goal, C.c_double), as_ptr(poses, C.c_float),
                             float(t_map), float(t_plan), as_ptr(status, C.c_int32),
                             traj.ctypes.data_as(C.POINTER(PopTrajRecord)))
        self._check(self.lib.pop_replan_cycle(self._h, C.byref(io)))
        self._poses = poses.copy()
        self._t_map = float(t_map)
        return status, traj

    def plan_cycle(self, start, goal, t_plan):
        """Planning half of the cycle (A* → corridor → QP → separation) on the SOGM of the last map update,
        e.g. the particle map's (`DspMaps.cloudCallback`): BaselinePlanner::replan (baseline.cpp:253-453)."""
        n = self.n_local
        start, goal = _f64(start, (n, 9)), _f64(goal, (n, 3))
        status = np.zeros(n, np.int32)
        traj = np.zeros(n, TRAJ_DTYPE)
        io = _lib.PopCycleIO(as_ptr(start, C.c_double), as_ptr(goal, C.c_double), None, 0.0, float(t_plan),
                             as_ptr(status, C.c_int32), traj.ctypes.data_as(C.POINTER(PopTrajRecord)))
        self._check(self.lib.pop_plan_cycle(self._h, C.byref(io)))
        return status, traj

    # -- after the optimiser ---------------------------------------------------------------------
    def traj_sample_batch(self, t_abs, all_agents: bool = False):
        """Trajectory-server sampling (bezier_traj_server.cpp:194-275): (pva [n, n_t, 9], valid [n, n_t])."""
        t_abs = _f64(t_abs).reshape(-1)
        n = self.n_total if all_agents else self.n_local
        out = np.zeros((n, len(t_abs), 9))
        valid = np.zeros((n, len(t_abs)), np.int8)
        self._check(self.lib.pop_traj_sample_batch(self._h, int(all_agents), as_ptr(t_abs, C.c_double), len(t_abs),
                                                   as_ptr(out, C.c_double), as_ptr(valid, C.c_int8)))
        return out, valid

    def swarm_advance(self, t_next: float):
        """Every local agent moves to its commanded state at t_next (perfect tracking); resident, async."""
        self._check(self.lib.pop_swarm_advance(self._h, float(t_next)))

    def fsm_update(self, t_now: float, replan_max_failures: int = 3, goal_tolerance: float = 1.0, hover_duration: float = 1.5):
        """FSM bookkeeping after a cycle (failure counters, hover trajectory, goal check): (states, failure counts)."""
        prm = _lib.PopFsmParams(int(replan_max_failures), 0, float(goal_tolerance), float(hover_duration))
        st = np.zeros(self.n_local, np.int32)
        fl = np.zeros(self.n_local, np.int32)
        self._check(self.lib.pop_fsm_update(self._h, float(t_now), C.byref(prm), as_ptr(st, C.c_int32), as_ptr(fl, C.c_int32)))
        return st, fl

    def swarm_states(self):
        out = np.zeros((self.n_local, 9))
        self._check(self.lib.pop_swarm_states(self._h, as_ptr(out, C.c_double)))
        return out

    def eval_metrics(self, t: float):
        """(min obstacle distance, min inter-agent distance) per local agent (multi_eval.hpp:240-275)."""
        mo, mu = np.zeros(self.n_local), np.zeros(self.n_local)
        self._check(self.lib.pop_eval_metrics(self._h, float(t), as_ptr(mo, C.c_double), as_ptr(mu, C.c_double)))
        return mo, mu

    def replan_cycle_resident(self, t_map, t_plan):
        self._check(self.lib.pop_replan_cycle_resident(self._h, float(t_map), float(t_plan)))

    def cycle_results(self):
        status = np.zeros(self.n_local, np.int32)
        traj = np.zeros(self.n_local, TRAJ_DTYPE)
        self._check(self.lib.pop_cycle_results(self._h, as_ptr(status, C.c_int32),
                                               traj.ctypes.data_as(C.POINTER(PopTrajRecord))))
        return status, traj

    def traj_publish_local(self):
        self._check(self.lib.pop_traj_publish_local(self._h))

    # -- multi-GPU exchange (csrc/pop_comm.cu) ------------------------------------------------
    def comm_unique_id(self) -> bytes:
        """ncclGetUniqueId: create on rank 0, hand to every rank"""
        buf = (C.c_uint8 * 128)()
        rc = self.lib.pop_comm_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise PopError(rc, "pop_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def comm_init(self, unique_id: bytes | None, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        self._check(self.lib.pop_comm_init(self._h, C.cast(buf, C.c_void_p) if buf is not None else None, int(rank), int(world)))

    def comm_info(self):
        r, w, v = C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self.lib.pop_comm_info(self._h, C.byref(r), C.byref(w), C.byref(v)))
        return r.value, w.value, v.value

    def exchange_trajs(self, pipelined: bool = False):
        """one ncclAllGather of this cycle's trajectory records (or the local publish on a single GPU); async.
        pipelined: own records at once, the other ranks' one call later (pop_exchange_trajs_pipelined)"""
        fn = self.lib.pop_exchange_trajs_pipelined if pipelined else self.lib.pop_exchange_trajs
        self._check(fn(self._h))

    def traj_buffers(self):
        s, r = C.c_void_p(), C.c_void_p()
        self._check(self.lib.pop_traj_buffers(self._h, C.byref(s), C.byref(r)))
        return s.value, r.value

    # -- measurement -----------------------------------------------------------------------------
    def profile_enable(self, on: bool | int = True):
        """0 off, 1 every stage (planning stages serialised over the batch), 2 map stages only (planning
        stages stay pipelined over agent groups)."""
        self._check(self.lib.pop_profile_enable(self._h, int(on)))

    def profile_read(self, reset: bool = True):
        ms = np.zeros(POP_STAGE_COUNT, np.float32)
        ln = np.zeros(POP_STAGE_COUNT, np.int64)
        self._check(self.lib.pop_profile_read(self._h, as_ptr(ms, C.c_float), as_ptr(ln, C.c_int64), int(reset)))
        return ({STAGE_NAMES[i]: float(ms[i]) for i in range(7)},
                {STAGE_NAMES[i]: int(ln[i]) for i in range(7)})

    def qp_profile(self, n: int) -> np.ndarray:
        out = np.zeros((n, 8), np.int64)
        self._check(self.lib.pop_qp_profile(self._h, n, as_ptr(out, C.c_int64)))
        return out

    def lp_solve(self, c, A, b, seed: int = 1, cooperative: bool = True):
        """the device LP on its own (diagnostic): min c.x s.t. A x <= b; returns (value, x)"""
        c = _f64(c)
        d = len(c)
        A = _f64(A).reshape(-1, d) if len(b) else np.zeros((0, d))
        b = _f64(b)
        x = np.zeros(d)
        v = C.c_double(0.0)
        self._check(self.lib.pop_lp_solve(self._h, d, len(b), as_ptr(c, C.c_double), as_ptr(A, C.c_double) if len(b) else None,
                                          as_ptr(b, C.c_double) if len(b) else None, int(seed), int(bool(cooperative)),
                                          as_ptr(x, C.c_double), C.byref(v)))
        return float(v.value), x

    def kernel_launches(self) -> int:
        return int(self.lib.pop_kernel_launches(self._h))


class AgentMap:
    """One agent's SOGM with the reference's method names (FakeParticleRiskVoxel / MapBase)."""

    def __init__(self, eng: PopEngine, agent: int):
        self.eng, self.agent = eng, agent

    def getMapCenter(self) -> np.ndarray:
        return self.eng._poses[self.agent].copy()

    def getMapTime(self) -> float:
        return self.eng._t_map

    def getClearOcccupancy(self, pos, t=0):
        """int t → slice index; float t → seconds after the map stamp. 0 free, 1 occupied, -1 out."""
        pos = _f64(pos).reshape(-1, 3)
        n = len(pos)
        out = np.zeros(n, np.int8)
        e = self.eng
        if np.issubdtype(np.asarray(t).dtype, np.integer):
            tt = np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.int32), (n,)))
            e._check(e.lib.pop_map_query(e._h, self.agent, as_ptr(pos, C.c_double), as_ptr(tt, C.c_int32), n,
                                         as_ptr(out, C.c_int8)))
        else:
            tt = np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.float64), (n,)))
            e._check(e.lib.pop_map_query_dt(e._h, self.agent, as_ptr(pos, C.c_double), as_ptr(tt, C.c_double), n,
                                            as_ptr(out, C.c_int8)))
        return out

    def getInflateOccupancy(self, pos, t=None):
        """MapBase::getInflateOccupancy (map.cpp:345-395): t None → (pos); int t → (pos, int t) slice index;
        float t → (pos, double t) seconds, testing the slices either side.  -1 out of range, 0, 1."""
        pos = _f64(pos).reshape(-1, 3)
        n = len(pos)
        out = np.zeros(n, np.int8)
        mode = 0 if t is None else (1 if np.issubdtype(np.asarray(t).dtype, np.integer) else 2)
        tt = np.ascontiguousarray(np.broadcast_to(np.asarray(0.0 if t is None else t, np.float64), (n,)))
        e = self.eng
        e._check(e.lib.pop_map_query_inflate(e._h, self.agent, as_ptr(pos, C.c_double), as_ptr(tt, C.c_double), mode, n,
                                             as_ptr(out, C.c_int8)))
        return out

    def getObstaclePoints(self, t_start, t_end, lower_corner, higher_corner, cap=None):
        e = self.eng
        cap = cap or e.cfg.max_obstacle_points
        lc, hc = _f64(lower_corner, (3,)), _f64(higher_corner, (3,))
        out = np.zeros((cap, 3))
        n = C.c_int(0)
        e._check(e.lib.pop_map_obstacle_points(e._h, self.agent, float(t_start), float(t_end), as_ptr(lc, C.c_double),
                                               as_ptr(hc, C.c_double), as_ptr(out, C.c_double), cap, C.byref(n)))
        return out[:n.value].copy()

    def addParticlesToRiskMap(self, pts, risks, t_index: int):
        pts = _f64(pts).reshape(-1, 3)
        risks = np.ascontiguousarray(risks, np.float32)
        e = self.eng
        e._check(e.lib.pop_map_add_particles(e._h, self.agent, as_ptr(pts, C.c_double), as_ptr(risks, C.c_float),
                                             len(pts), int(t_index)))

    def risk_maps(self) -> np.ndarray:
        """risk_maps_[VOXEL_NUM][PREDICTION_TIMES] as float32 (plan_env/include/plan_env/map.h:52)."""
        e = self.eng
        out = np.zeros((e.cfg.nx * e.cfg.ny * e.cfg.nz, e.cfg.nt), np.float32)
        e._check(e.lib.pop_map_download(e._h, self.agent, as_ptr(out, C.c_float)))
        return out
