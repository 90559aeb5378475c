"""CPU property checks of the oracle's hybrid A* (no reference tests exist for A*: parity
unpinned, so the restatement is checked through invariants of
path_searching/src/fake_risk_hybrid_a_star.cpp)."""
import numpy as np

import pred_occ_planner_b200 as p
from tests import cases


def _setup(oracle, pose, t_map=0.0):
    cfg = cases.cfg_c1()
    cloud, cyl = cases.world_c1()
    m = oracle.map(cfg)
    m.set_pose(pose)
    m.update_gt(cloud, cyl, t_map)
    return cfg, m


def test_path_is_dynamically_consistent_and_free(oracle):
    cfg, m = _setup(oracle, [-4, 0, 1.0])
    start = np.zeros(9)
    start[:3] = [-4, 0, 1.0]
    ret, path, stats = m.astar(start, [4, 0, 1.0], 0.02)
    assert ret in (3, 4, 5)
    assert len(path) >= 2 and stats[0] > 0 and stats[1] > stats[0]
    # first sample is the start state; samples are corridor_tau apart under a constant
    # acceleration bounded by max_acc: p1 = p0 + v0 t + a t^2/2, v1 = v0 + a t
    assert np.allclose(path[0, :3], start[:3])
    tau = cfg.corridor_tau
    for a, b in zip(path[:-1], path[1:]):
        acc = (b[3:] - a[3:]) / tau
        assert np.all(np.abs(acc) <= cfg.max_acc + 1e-9)
        assert np.allclose(b[:3], a[:3] + a[3:] * tau + 0.5 * acc * tau * tau, atol=1e-9)
        assert np.all(np.abs(b[3:]) <= cfg.max_vel + 1e-9)
    # every node on the path was accepted by the collision check at its own time
    t = 0.02 + tau * np.arange(1, len(path))
    assert (m.query(path[1:, :3], t) == 0).all()


def test_start_near_goal_is_no_path(oracle):
    cfg, m = _setup(oracle, [0, 0, 1.0])
    start = np.zeros(9)
    start[:3] = [0.3, 0.2, 1.0]
    ret, path, stats = m.astar(start, [0.35, 0.2, 1.0], 0.0)
    assert ret == 0 and len(path) == 0 and stats[2] == 1  # retried once with init=false


def test_deterministic(oracle):
    cfg, m = _setup(oracle, [-2, 1, 1.0])
    start = np.zeros(9)
    start[:6] = [-2, 1, 1.0, 0.5, -0.2, 0.0]
    r1 = m.astar(start, [3, -3, 1.2], 0.1)
    r2 = m.astar(start, [3, -3, 1.2], 0.1)
    assert r1[0] == r2[0] and np.array_equal(r1[1], r2[1]) and np.array_equal(r1[2], r2[2])


def test_goal_equal_to_start_terminates(oracle):
    """The first popped node is already at the goal and the optimal shot time is 0: the reference's
    computeShotTraj then loops `time += 0` forever (fake_risk_hybrid_a_star.cpp:500-515).  Treated as "no shot":
    the search ends with NO_PATH, like any near-end hit without a parent."""
    cfg, m = _setup(oracle, [0, 0, 1.0])
    start = np.zeros(9)
    start[:3] = [0.3, 0.2, 1.0]
    ret, path, stats = m.astar(start, start[:3], 0.0)
    assert ret == 0 and len(path) == 0
