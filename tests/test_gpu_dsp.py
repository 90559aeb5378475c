"""GPU parity of the DSP particle map (K1d) against the sequential oracle, through the C ABI.

Bit-exact: voxel-filter output, observation bins, C_k, cluster velocities, pyramid tables, every particle
(flag, velocity, position, weight, slot), drop counters, noise-table sequence counters, voxel occupancy.
fp32 future-risk sums: 1e-5 relative (atomic accumulation order), occupied sets equal away from the threshold."""
import numpy as np
import pytest

from tests import cases, dsp_compare

pytestmark = pytest.mark.gpu


def _setup(pop, oracle, n_agents, seed=5, sigma=0.1):
    from pred_occ_planner_b200 import dsp, worldgen
    cfg = cases.dsp_cfg_ref(n_local=n_agents, n_total=n_agents)
    d = cases.dsp_params(cfg, sigma_observation=sigma)
    eng = pop.PopEngine(cfg, device=0)
    maps = dsp.DspMaps(eng, d)
    tabs = worldgen.dsp_random_tables(d.n_randoms, seed)
    maps.set_random_tables(*tabs)
    orcs = []
    for _ in range(n_agents):
        o = oracle.dsp(cfg, d)
        o.set_random_tables(*tabs)
        orcs.append(o)
    return cfg, d, eng, maps, orcs


def _flight(a, k):
    pos, quat, st = cases.dsp_flight(k, speed=1.0 + 0.5 * a)
    pos = pos + np.array([0.0, 0.8 * a, 0.1 * a], np.float32)
    return pos, quat, st


def test_voxel_filter_bit_exact(pop, oracle):
    from pred_occ_planner_b200 import worldgen
    cfg, d, eng, maps, orcs = _setup(pop, oracle, 1)
    cyl = cases.dsp_scene()
    for k in (0, 7):
        pos, quat, _ = _flight(0, k)
        cam = worldgen.render_depth_cloud(cyl, pos, quat, t=0.1 * k, noise=0.01, seed=k)
        cam[5] = np.nan      # PCL drops non-finite points
        got = maps.filterPointCloud(0, cam)
        ref = orcs[0].filter(cam)
        assert got.shape == ref.shape and len(ref) > 100
        assert np.array_equal(got, ref)
    # empty and single-point clouds
    assert len(maps.filterPointCloud(0, np.zeros((0, 3), np.float32))) == 0
    one = np.array([[0.1, 0.2, 2.0]], np.float32)
    assert np.array_equal(maps.filterPointCloud(0, one), orcs[0].filter(one))
    eng.close()


def test_point_cap_of_filter(pop, oracle):
    """More leaves than MAX_POINT_NUM: the first max_points valid centroids are kept (risk_mapping_node.cpp:93-95)."""
    cfg = cases.dsp_cfg_ref()
    d = cases.dsp_params(cfg)
    d.max_points = 300
    from pred_occ_planner_b200 import dsp, worldgen
    eng = pop.PopEngine(cfg, device=0)
    maps = dsp.DspMaps(eng, d)
    o = oracle.dsp(cfg, d)
    pos, quat, _ = _flight(0, 3)
    cam = worldgen.render_depth_cloud(cases.dsp_scene(), pos, quat)
    got, ref = maps.filterPointCloud(0, cam), o.filter(cam)
    assert len(ref) == 300 and np.array_equal(got, ref)
    eng.close()


@pytest.mark.parametrize("n_agents,frames", [(1, 10), (3, 6)])
def test_particle_map_sequence(pop, oracle, n_agents, frames):
    from pred_occ_planner_b200 import worldgen
    cfg, d, eng, maps, orcs = _setup(pop, oracle, n_agents)
    cyl = cases.dsp_scene()
    dropped = np.zeros(2, np.int64)
    for k in range(frames):
        clouds, poses, quats = [], [], []
        for a in range(n_agents):
            pos, quat, st = _flight(a, k)
            clouds.append(worldgen.render_depth_cloud(cyl, pos, quat, t=0.1 * k))
            poses.append(pos)
            quats.append(quat)
        ok = maps.cloudCallback(clouds, np.array(poses), np.array(quats), st)
        assert ok.tolist() == [1] * n_agents
        for a in range(n_agents):
            assert orcs[a].update(clouds[a], poses[a], quats[a], st) == 1
            risk = eng.agent_map(a).risk_maps()
            bad = dsp_compare.compare_frame(maps, a, orcs[a], risk, d.occupancy_threshold)
            assert not bad, f"frame {k} agent {a}: " + "; ".join(bad)
        dropped += maps.stats(0)[:2]
    # the order-dependent paths were exercised: particles were dropped by full voxels and full pyramids
    assert dropped[0] > 0 and dropped[1] > 0
    eng.close()


def test_filtered_entry_and_rejected_odometry(pop, oracle):
    from pred_occ_planner_b200 import worldgen
    cfg, d, eng, maps, orcs = _setup(pop, oracle, 1)
    cyl = cases.dsp_scene()
    o = orcs[0]
    for k in range(4):
        pos, quat, st = _flight(0, k)
        cam = worldgen.render_depth_cloud(cyl, pos, quat, t=0.1 * k)
        pts = o.filter(cam)
        if k == 2:
            badq = np.array([1.5, 0, 0, 0], np.float32)   # |q| component > 1.001 -> update returns 0 (:183-188)
            assert maps.cloudCallback([pts], pos[None], badq[None], st, filtered=True).tolist() == [0]
            assert o.update_filtered(pts, pos, badq, st) == 0
        ok = maps.cloudCallback([pts], pos[None], quat[None], st, filtered=True)
        assert ok.tolist() == [1] and o.update_filtered(pts, pos, quat, st) == 1
        bad = dsp_compare.compare_frame(maps, 0, o, eng.agent_map(0).risk_maps(), d.occupancy_threshold)
        assert not bad, f"frame {k}: " + "; ".join(bad)
    # a jump of more than 10 m is rejected too (:196-204) and leaves the map untouched
    before = maps.particles(0)
    far = np.array([20.0, 0.0, 1.0], np.float32)
    assert maps.cloudCallback([pts], far[None], quat[None], st + 0.1, filtered=True).tolist() == [0]
    assert o.update_filtered(pts, far, quat, st + 0.1) == 0
    assert np.array_equal(maps.particles(0), before)
    eng.close()


def test_occupied_export_and_queries(pop, oracle):
    """getOccupancyMap centres and RiskBase-style region queries on the particle-fed SOGM."""
    from pred_occ_planner_b200 import worldgen
    cfg, d, eng, maps, orcs = _setup(pop, oracle, 1)
    cyl = cases.dsp_scene()
    o = orcs[0]
    for k in range(6):
        pos, quat, st = _flight(0, k)
        cam = worldgen.render_depth_cloud(cyl, pos, quat, t=0.1 * k)
        maps.cloudCallback([cam], pos[None], quat[None], st)
        o.update(cam, pos, quat, st)
        if k < 5:
            o.export()
    vox = o.voxels()
    idx = np.nonzero(vox[:, 0] > d.occupancy_threshold)[0]
    got = maps.getOccupancyMap(0)
    assert len(got) == len(idx) > 50
    res, hx, hy, hz = np.float32(cfg.resolution), *[np.float32(cfg.resolution) * np.float32(n) * np.float32(0.5)
                                                     for n in (cfg.nx, cfg.ny, cfg.nz)]
    x, y, z = idx % cfg.nx, (idx // cfg.nx) % cfg.ny, idx // (cfg.nx * cfg.ny)
    ref = np.stack([x.astype(np.float32) * res + (-hx + res * np.float32(0.5)),
                    y.astype(np.float32) * res + (-hy + res * np.float32(0.5)),
                    z.astype(np.float32) * res + (-hz + res * np.float32(0.5))], axis=1)
    assert np.array_equal(got, ref)
    # region queries read the SOGM the particle map just wrote: compare with the oracle grid loaded with
    # the device's own risk values (the query path itself is bit-exact)
    risk = eng.agent_map(0).risk_maps()
    om = oracle.map(cfg)
    om.set_pose(pos)
    om.set_risk(risk)
    rng = np.random.default_rng(3)
    q = pos[None].astype(np.float64) + rng.uniform([-4.5, -4.5, -1.0], [4.5, 4.5, 1.5], (4096, 3))
    t = rng.integers(0, cfg.nt, 4096).astype(np.int32)
    assert np.array_equal(eng.agent_map(0).getClearOcccupancy(q, t), om.query(q, t))
    assert (om.query(q, t) == 1).sum() > 20
    eng.close()


def test_plan_cycle_on_particle_map(pop, oracle):
    """The real-perception cycle: particle-map frames, then A* -> corridors -> QP on the SOGM they wrote
    (pop_plan_cycle).  The oracle plans on a grid loaded with the device's own risk values, so the planning
    half is compared at the north-star tolerance: status identical, control points within 1e-4 relative."""
    from pred_occ_planner_b200 import worldgen
    from tests import oracle_cycle
    n = 2
    cfg, d, eng, maps, orcs = _setup(pop, oracle, n)
    cyl = cases.dsp_scene()
    for k in range(6):
        clouds, poses, quats = [], [], []
        for a in range(n):
            pos, quat, st = _flight(a, k)
            clouds.append(worldgen.render_depth_cloud(cyl, pos, quat, t=0.1 * k))
            poses.append(pos)
            quats.append(quat)
        maps.cloudCallback(clouds, np.array(poses), np.array(quats), st)
    start = np.zeros((n, 9))
    goal = np.zeros((n, 3))
    for a in range(n):
        start[a, :3] = poses[a]
        start[a, 3] = 1.0
        goal[a] = [4.0, 0.5 * a, 1.0]
    status, traj = eng.plan_cycle(start, goal, st + 0.02)
    n_ok = 0
    for a in range(n):
        om = oracle.map(cfg)
        om.set_pose(poses[a])
        om.set_time(st)
        om.set_risk(eng.agent_map(a).risk_maps())
        rs, rec, stats = oracle_cycle.replan(om, None, a, start[a], goal[a], st + 0.02)
        assert rs == status[a], (a, rs, status[a], stats.tolist())
        if rs == 0:
            n_ok += 1
            M = int(rec["n_pieces"])
            assert traj["n_pieces"][a] == M
            g, r = traj["cpts"][a][:5 * M], rec["cpts"][:5 * M]
            assert np.abs(g - r).max() <= 1e-4 * max(1.0, np.abs(r).max())
    assert n_ok >= 1
    eng.close()
