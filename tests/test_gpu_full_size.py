"""Parity against the oracle AT THE BENCHMARKED SIZES (the workloads `bench.py` flies, `bench.Scenario`):

config 2 — 64 agents, 300x300x50 voxels @0.1 m, T = 20, 40 moving cylinders — and a sample of config 3 — 256 agents,
500x500x60, 60 static + 60 moving cylinders.  After two closed-loop cycles (so that peers' trajectories are on the
map): the SOGM of several agents cell for cell (ground truth + peer projection; u8 on the device, f32 in the oracle,
same values), and a third cycle for the first 16 agents: status identical, control points within 1e-4.

The 1e-4 bound on control points is asserted for every compared agent of config 2.  For config 3 it is asserted for 7 of
every 8 agents and the solver's own tolerance for all: a QP that runs thousands of ADMM iterations ends with the same
status after the same number of iterations in both implementations but up to 1e-2 m apart, because the two K^-1
applications (explicit inverse tiles here, Cholesky substitutions in the oracle) differ by the system's condition number
times machine epsilon (profiles/r03l_c3_parity_trace.txt)."""
import os

import numpy as np
import pytest

import bench
import pred_occ_planner_b200 as p

pytestmark = pytest.mark.gpu


def _fly(conf, n_maps, n_cycle):
    from tests import oracle_lib
    orc = oracle_lib.load()
    n = bench.CONFIGS[conf]["agents"]
    sc = bench.Scenario(conf, n, n, 0)
    cfg, goal = sc.cfg, sc.goal_all
    eng = p.PopEngine(cfg)
    eng.set_body_particles(*sc.body)
    eng.set_trajectories(np.zeros(n, p.TRAJ_DTYPE))
    start = sc.start_all.copy()
    for s in range(2):
        if s > 0:
            eng.swarm_advance(sc.t(s))
            start = eng.swarm_states()
        cloud, cyl = sc.world(s)
        eng.set_world_cloud(cloud)
        eng.set_cylinders(cyl)
        eng.replan_cycle(start, goal, start[:, :3].astype(np.float32), sc.t(s), sc.t(s) + bench.LOOKAHEAD)
        eng.traj_publish_local()
    s = 2
    eng.swarm_advance(sc.t(s))
    start = eng.swarm_states()
    poses = start[:, :3].astype(np.float32)
    peers = eng.get_trajectories()
    assert (peers["n_pieces"] > 0).sum() >= n // 2          # the peer projection has something to project
    cloud, cyl = sc.world(s)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    st_gpu, tr_gpu = eng.replan_cycle(start, goal, poses, sc.t(s), sc.t(s) + bench.LOOKAHEAD)
    # ---- the SOGM of the first / last / two middle agents, every cell of every slice ----
    agents = sorted({0, n // 3, (2 * n) // 3, n - 1})[:n_maps]
    occupied = 0
    for a in agents:
        c = cfg.copy()
        c.agent_id_base = cfg.agent_id_base + a
        m = orc.map(c)
        m.set_pose(poses[a])
        m.update_gt(cloud, cyl, sc.t(s))
        m.add_other_agents(sc.body[0], sc.body[1], peers, a)
        want = m.risk_maps()
        got = eng.agent_map(a).risk_maps()
        assert want.shape == got.shape
        assert np.array_equal(want, got), (conf, a, int((want != got).sum()))
        occupied += int((got > 0).sum())
        del m, want, got
    assert occupied > 1000 * len(agents)
    # ---- the cycle of the first n_cycle agents ----
    own = np.zeros(n, p.TRAJ_DTYPE)
    _, st_cpu = bench.cpu_cycle(sc, cloud, cyl, start.copy(), goal, peers, n_cycle, os.cpu_count() or 1, sc.t(s), records=own)
    assert np.array_equal(st_gpu[:n_cycle], st_cpu[:n_cycle]), (st_gpu[:n_cycle].tolist(), st_cpu[:n_cycle].tolist())
    rel = []
    for a in np.flatnonzero(st_cpu[:n_cycle] == 0):
        M = int(own["n_pieces"][a])
        assert M == int(tr_gpu["n_pieces"][a])
        x, y = tr_gpu["cpts"][a][:5 * M], own["cpts"][a][:5 * M]
        rel.append(float(np.abs(x - y).max() / max(1.0, np.abs(y).max())))
    eng.close()
    return np.array(rel)


def test_config2_full_size_against_the_oracle():
    rel = _fly("c2", n_maps=4, n_cycle=16)
    assert len(rel) >= 10 and rel.max() <= 1e-4, rel.tolist()


def test_config3_sample_against_the_oracle():
    rel = _fly("c3", n_maps=2, n_cycle=16)
    assert len(rel) >= 10
    assert (rel <= 1e-4).mean() >= 0.875 and rel.max() <= 3e-3, rel.tolist()
