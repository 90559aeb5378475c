"""ctypes helpers for the oracle's replan driver (test infrastructure)."""
import ctypes as C

import numpy as np

import pred_occ_planner_b200 as p
from tests import oracle_lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def replan(om, swarm, ego_id, start9, goal3, t_plan):
    """orc::replan on an OracleMap that already holds the updated grid (peers included).
    swarm = (offs, counts, recs) or None. Returns (status, record, stats)."""
    lib = om.lib
    start9 = np.ascontiguousarray(start9, np.float64).reshape(9)
    goal3 = np.ascontiguousarray(goal3, np.float64).reshape(3)
    rec = np.zeros(1, p.TRAJ_DTYPE)
    stats = np.zeros(4, np.int32)
    if swarm is None:
        st = lib.orc_replan(om.h, None, None, None, int(ego_id), _p(start9, C.c_double), _p(goal3, C.c_double),
                            C.c_double(t_plan), C.c_void_p(rec.ctypes.data), _p(stats, C.c_int32))
    else:
        offs = np.ascontiguousarray(swarm[0], np.float64)
        counts = np.ascontiguousarray(swarm[1], np.int32)
        recs = np.ascontiguousarray(swarm[2])
        st = lib.orc_replan(om.h, _p(offs, C.c_double), _p(counts, C.c_int32), C.c_void_p(recs.ctypes.data),
                            int(ego_id), _p(start9, C.c_double), _p(goal3, C.c_double), C.c_double(t_plan),
                            C.c_void_p(rec.ctypes.data), _p(stats, C.c_int32))
    return st, rec[0], stats


def safe_after_opt(cfg, swarm, ego_id, cpts, M, t0):
    lib = oracle_lib.load().lib
    offs = np.ascontiguousarray(swarm[0], np.float64)
    counts = np.ascontiguousarray(swarm[1], np.int32)
    recs = np.ascontiguousarray(swarm[2])
    cpts = np.ascontiguousarray(cpts, np.float64)
    return bool(lib.orc_safe_after_opt(C.byref(cfg), _p(offs, C.c_double), _p(counts, C.c_int32),
                                       C.c_void_p(recs.ctypes.data), int(ego_id), _p(cpts, C.c_double), int(M),
                                       C.c_double(t0)))


def cycle_batch(cfg, a0, a1, n_threads, cloud, cyl, poses, swarm, start, goal, t_map, t_plan, records=None):
    """Multi-threaded CPU cycle (map update + replan) for agents [a0, a1); returns (seconds, status).
    records (n_agents_local TRAJ_DTYPE, optional): the entries of successful replans are overwritten with the new
    trajectory records, failed ones keep what they held (as k_finalize does)."""
    lib = oracle_lib.load().lib
    lib.orc_cycle_batch.restype = C.c_double
    cloud = np.ascontiguousarray(cloud, np.float32)
    cyl = np.ascontiguousarray(cyl)
    poses = np.ascontiguousarray(poses, np.float32)
    start = np.ascontiguousarray(start, np.float64)
    goal = np.ascontiguousarray(goal, np.float64)
    status = np.full(cfg.n_agents_local, -1, np.int32)
    if swarm is None:
        offs_p = counts_p = recs_p = None
    else:
        offs = np.ascontiguousarray(swarm[0], np.float64)
        counts = np.ascontiguousarray(swarm[1], np.int32)
        recs = np.ascontiguousarray(swarm[2])
        offs_p, counts_p, recs_p = _p(offs, C.c_double), _p(counts, C.c_int32), C.c_void_p(recs.ctypes.data)
    sec = lib.orc_cycle_batch(C.byref(cfg), int(a0), int(a1), int(n_threads), _p(cloud, C.c_float), len(cloud),
                              C.c_void_p(cyl.ctypes.data), len(cyl), _p(poses, C.c_float), offs_p, counts_p, recs_p,
                              _p(start, C.c_double), _p(goal, C.c_double), C.c_double(t_map), C.c_double(t_plan),
                              _p(status, C.c_int32), C.c_void_p(records.ctypes.data) if records is not None else None)
    return sec, status
