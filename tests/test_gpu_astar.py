"""GPU parity of the batched hybrid A* (K2) against the oracle: same ASTAR_RET, same number of
expansions and allocated nodes, and the sampled path (positions + velocities) within 1e-9 —
it is bit-identical unless libm/CUDA transcendental rounding flips a heap comparison."""
import numpy as np
import pytest

import pred_occ_planner_b200 as p
from tests import cases

pytestmark = pytest.mark.gpu


def _run(oracle, cfg, cloud, cyl, start, goal, poses, t_map, t_start, trajs=None, body=None):
    n = cfg.n_agents_local
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    eng.set_poses(poses)
    if trajs is not None:
        eng.set_body_particles(*body)
        eng.set_trajectories(trajs)
    eng.map_update_gt(t_map)
    ret, path, ln, stats = eng.astar_search_batch(start, goal, np.full(n, t_start))
    mism = []
    for a in range(n):
        m = oracle.map(cfg)
        m.set_pose(poses[a])
        m.update_gt(cloud, cyl, t_map)
        if trajs is not None:
            m.add_other_agents(body[0], body[1], trajs, a)
        oret, opath, ostats = m.astar(start[a], goal[a], t_start)
        ok = (oret == ret[a] and len(opath) == ln[a] and ostats[0] == stats[a, 0] and ostats[1] == stats[a, 1]
              and ostats[2] == stats[a, 2] and np.allclose(opath, path[a, :ln[a]], rtol=0, atol=1e-9))
        if not ok:
            mism.append((a, oret, int(ret[a]), ostats.tolist(), stats[a].tolist()))
    return mism, ret, stats


@pytest.mark.parametrize("cell", [p.POP_CELL_F32, p.POP_CELL_U8])
def test_c1_single_agent(oracle, cell):
    cfg = cases.cfg_c1(cell)
    cloud, cyl = cases.world_c1()
    start = np.zeros((1, 9))
    start[0, :3] = [-4, 0, 1.0]
    goal = np.array([[4.0, 0, 1.0]])
    mism, ret, stats = _run(oracle, cfg, cloud, cyl, start, goal, start[:, :3].astype(np.float32), 0.0, 0.02)
    assert not mism, mism
    assert ret[0] in (3, 4, 5) and stats[0, 0] > 1


def test_many_agents_static_world(oracle):
    n = 32
    cfg = cases.cfg_c1(p.POP_CELL_U8, n_local=n, n_total=n)
    cloud, cyl = cases.world_c1(seed=99)
    rng = np.random.default_rng(5)
    start = np.zeros((n, 9))
    start[:, :2] = rng.uniform(-4.5, 4.5, (n, 2))
    start[:, 2] = rng.uniform(0.6, 1.6, n)
    start[:, 3:6] = rng.uniform(-1.0, 1.0, (n, 3)) * [1, 1, 0.2]
    start[:, 6:9] = rng.uniform(-2.0, 2.0, (n, 3)) * [1, 1, 0.2]
    goal = np.concatenate([rng.uniform(-4.5, 4.5, (n, 2)), rng.uniform(0.8, 1.5, (n, 1))], axis=1)
    goal[3] = start[3, :3] + [0.05, 0.0, 0.0]   # start next to goal → NO_PATH after the retry
    poses = start[:, :3].astype(np.float32)
    mism, ret, stats = _run(oracle, cfg, cloud, cyl, start, goal, poses, 1.0, 0.02)
    assert not mism, mism
    assert ret[3] in (0, 4, 5)   # near_end on the very first pop: shot, or NO_PATH after the retry
    assert (ret != 0).sum() >= n // 2


def test_moving_obstacles_and_peers(oracle):
    n = 16
    cfg = cases.cfg_small_c2(p.POP_CELL_U8, n_local=n, n_total=n, nx=160, ny=160, nz=30, nt=20)
    cloud, cyl = cases.world_moving(21, (16.0, 16.0, 3.0), 6, 18)
    start, goal = p.worldgen.circle_agents(n, 6.0)
    poses = start[:, :3].astype(np.float32)
    trajs = p.worldgen.straight_trajs(start[:, :3], goal, 4.9, 1.5, 8, 0.2)
    body = p.worldgen.body_table(n, [0.4, 0.4, 0.45])
    mism, ret, stats = _run(oracle, cfg, cloud, cyl, start, goal, poses, 5.0, 0.02, trajs, body)
    assert not mism, mism
    assert (ret != 0).sum() >= n // 2
    assert stats[:, 0].max() > 5


def test_repeated_searches_reuse_pools(oracle):
    """Generation-stamped hash: a second batch on the same handle must not see stale nodes."""
    cfg = cases.cfg_c1(p.POP_CELL_U8, n_local=2, n_total=2)
    cloud, cyl = cases.world_c1()
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    start = np.zeros((2, 9))
    start[:, :3] = [[-4, 0, 1.0], [3, 3, 1.0]]
    goal = np.array([[4.0, 0, 1.0], [-3, -3, 1.0]])
    eng.set_poses(start[:, :3])
    eng.map_update_gt(0.0)
    first = eng.astar_search_batch(start, goal, np.full(2, 0.02))
    for _ in range(3):
        again = eng.astar_search_batch(start, goal, np.full(2, 0.02))
        for x, y in zip(first, again):
            assert np.array_equal(x, y)


def test_goal_equal_to_start_terminates(oracle):
    """Degenerate shot (optimal time 0) must end the search instead of spinning (see the CPU test)."""
    n = 3
    cfg = cases.cfg_c1(p.POP_CELL_U8, n_local=n, n_total=n)
    cloud, cyl = cases.world_c1()
    start = np.zeros((n, 9))
    start[:, :3] = [[-4, 0, 1.0], [0.3, 0.2, 1.0], [2.0, -1.0, 1.2]]
    goal = start[:, :3].copy()
    goal[0] = [4.0, 0, 1.0]          # one ordinary search beside two degenerate ones
    mism, ret, stats = _run(oracle, cfg, cloud, cyl, start, goal, start[:, :3].astype(np.float32), 0.0, 0.02)
    assert not mism, mism
    assert ret[0] in (3, 4, 5) and ret[1] == 0 and ret[2] == 0
