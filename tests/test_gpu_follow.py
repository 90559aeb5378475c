"""GPU parity of the steps after the optimiser: trajectory-server sampling (bit-exact against the oracle's
Bezier evaluation), the closed-loop state hand-over, the evaluation metrics, and a short closed-loop flight."""
import numpy as np
import pytest

import pred_occ_planner_b200 as p
from tests import cases

pytestmark = pytest.mark.gpu


def _engine(n, with_world=True):
    cfg = cases.cfg_small_c2(p.POP_CELL_U8, n_local=n, n_total=n, nx=160, ny=160, nz=30, nt=20)
    cloud, cyl = cases.world_moving(21, (16.0, 16.0, 3.0), 6, 6)
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    eng.set_body_particles(*p.worldgen.body_table(n, [0.4, 0.4, 0.45]))
    eng.set_trajectories(np.zeros(n, p.TRAJ_DTYPE))
    return cfg, eng, cloud, cyl


def test_sampling_matches_bezier_evaluation(oracle):
    n = 5
    cfg, eng, _, _ = _engine(n)
    sp = np.random.default_rng(0).uniform(-3, 3, (n, 3))
    gp = np.random.default_rng(1).uniform(-3, 3, (n, 3))
    recs = p.worldgen.straight_trajs(sp, gp, 4.0, 1.3, 6, 0.3)
    recs["cpts"] += np.random.default_rng(2).normal(0, 0.2, recs["cpts"].shape)   # curved pieces
    recs["n_pieces"][4] = 0
    eng.set_trajectories(recs)
    t = np.array([3.9, 4.0, 4.3, 4.95, 5.8, 5.9])
    pva, valid = eng.traj_sample_batch(t, all_agents=True)
    for a in range(n):
        M = int(recs["n_pieces"][a])
        if M == 0:
            assert not valid[a].any()
            continue
        dur = recs["duration"][a][:M].astype(np.float64)
        T = dur.sum()
        for k, tk in enumerate(t):
            inside = 0.0 <= tk - 4.0 <= T
            assert bool(valid[a, k]) == inside
            if inside:
                for d in range(3):
                    ref = oracle.bezier_eval(dur, recs["cpts"][a][:5 * M], [tk - 4.0], d)[0]
                    assert np.array_equal(pva[a, k, 3 * d:3 * d + 3], ref)
    eng.close()


def test_closed_loop_flight(oracle):
    """Ten resident cycles with the agents following what they planned: states stay on the planned
    trajectories, metrics agree with a direct computation, agents make progress and keep their distance."""
    n = 8
    cfg, eng, cloud, cyl = _engine(n)
    start, goal = p.worldgen.circle_agents(n, 5.0)
    poses = start[:, :3].astype(np.float32)
    t0 = 50.0
    status, traj = eng.replan_cycle(start, goal, poses, t0, t0 + 0.02)
    eng.traj_publish_local()
    prev = traj.copy()
    dist0 = np.linalg.norm(start[:, :3] - goal, axis=1)
    for k in range(1, 11):
        t = t0 + 0.1 * k
        eng.swarm_advance(t + 0.02)
        st = eng.swarm_states()
        # the hand-over is the sampled command of the trajectory each agent last published
        pva, valid = eng.traj_sample_batch([t + 0.02])
        for a in range(n):
            if valid[a, 0]:
                assert np.array_equal(st[a], pva[a, 0])
        mo, mu = eng.eval_metrics(t + 0.02)
        pos = st[:, :3]
        ref_uav = np.array([min(np.linalg.norm(pos[b] - pos[a]) for b in range(n) if b != a) for a in range(n)])
        ref_obs = np.array([min(np.hypot(np.float32(c["x"]) - pos[a, 0], np.float32(c["y"]) - pos[a, 1]) - c["w"] / 2
                                for c in cyl) for a in range(n)])
        assert np.allclose(mu, ref_uav, rtol=0, atol=1e-12) and np.allclose(mo, ref_obs, rtol=0, atol=1e-12)
        eng.replan_cycle_resident(t, t + 0.02)
        status, traj = eng.cycle_results()
        eng.traj_publish_local()
    st = eng.swarm_states()
    dist1 = np.linalg.norm(st[:, :3] - goal, axis=1)
    assert (dist1 < dist0 - 0.3).sum() >= n // 2          # most agents moved towards their goals
    mo, mu = eng.eval_metrics(t0 + 1.02)
    assert mu.min() > 0.2 and mo.min() > 0.0              # no agent inside an obstacle or on top of a peer
    eng.close()


def test_is_traj_safe(oracle):
    """pop_traj_safe_batch = BaselinePlanner::isTrajSafe (baseline.cpp:45-68): after a cycle, the world moves on and
    every agent's own trajectory is re-checked against the new map from several wall-clock times and horizons —
    booleans and the time of the first occupied sample identical to the oracle (itself pinned against the compiled
    reference planners, tests/test_ref_pins.py)."""
    n = 10
    cfg, eng, cloud, cyl = _engine(n)
    start, goal = p.worldgen.circle_agents(n, 5.5)
    poses = start[:, :3].astype(np.float32)
    st, tr = eng.replan_cycle(start, goal, poses, 40.0, 40.02)
    assert (st == 0).sum() >= n // 2
    unsafe = 0
    for t_map, drift in ((40.0, 0.0), (40.5, 0.25), (41.0, 0.6)):
        # later map stamps seen from drifted map centres: some samples have become occupied
        poses2 = (poses + np.float32([drift, -0.5 * drift, 0.0])).astype(np.float32)
        eng.set_poses(poses2)
        eng.map_update_gt(t_map)
        oms = []
        for a in range(n):
            m = oracle.map(cfg)
            m.set_pose(poses2[a])
            m.update_gt(cloud, cyl, t_map)
            oms.append(m)
        for now, T in ((39.9, 2.0), (40.02, 2.0), (40.33, 1.0), (40.7, 5.0), (41.4, 0.5), (43.0, 2.0)):
            safe, tu = eng.is_traj_safe(now, T)
            for a in range(n):
                if st[a] != 0:
                    assert safe[a] == 1 and tu[a] == -1.0      # no trajectory: nothing to sample
                    continue
                want, wt = oms[a].traj_safe(tr[a], cfg.corridor_tau, [now], T)
                assert safe[a] == want[0] and tu[a] == wt[0], (t_map, now, T, a, safe[a], want[0], tu[a], wt[0])
                unsafe += int(want[0] == 0)
    assert unsafe > 0
    eng.close()
