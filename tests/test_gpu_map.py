"""GPU parity (through the C ABI) of the SOGM path against the oracle: ground-truth voxel update
(K1f), peer projection (K5), occupancy queries, obstacle-point extraction.  Voxel occupancy is
compared bit-exactly as the set of (voxel index, slice) plus the cell values."""
import numpy as np
import pytest

import pred_occ_planner_b200 as p
from tests import cases

pytestmark = pytest.mark.gpu


def _engine_and_oracles(oracle, cfg, cloud, cyl, poses, t_map, trajs=None, body=None):
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    eng.set_poses(poses)
    if trajs is not None:
        eng.set_body_particles(*body)
        eng.set_trajectories(trajs)
    eng.map_update_gt(t_map)
    oms = []
    for a in range(cfg.n_agents_local):
        m = oracle.map(cfg)
        m.set_pose(poses[a])
        m.update_gt(cloud, cyl, t_map)
        if trajs is not None:
            m.add_other_agents(body[0], body[1], trajs, cfg.agent_id_base + a)
        oms.append(m)
    return eng, oms


@pytest.mark.parametrize("cell", [p.POP_CELL_F32, p.POP_CELL_U8])
def test_c1_static_pillars_bit_exact(oracle, cell):
    cfg = cases.cfg_c1(cell)
    cloud, cyl = cases.world_c1()
    poses = np.array([[-4.0, 0.0, 1.0]], np.float32)
    eng, oms = _engine_and_oracles(oracle, cfg, cloud, cyl, poses, 12.5)
    got, ref = eng.agent_map(0).risk_maps(), oms[0].risk_maps()
    assert cases.occupied_set(got) == cases.occupied_set(ref)
    assert np.array_equal(got, ref)
    assert np.count_nonzero(ref) > 1000


@pytest.mark.parametrize("cell", [p.POP_CELL_F32, p.POP_CELL_U8])
def test_moving_cylinders_multi_agent_bit_exact(oracle, cell):
    cfg = cases.cfg_small_c2(cell, n_local=4, n_total=4)
    cloud, cyl = cases.world_moving(77, (12.0, 12.0, 3.0), 5, 15)
    rng = np.random.default_rng(3)
    poses = (rng.uniform(-1.5, 1.5, (4, 3)) * [1, 1, 0.2] + [0, 0, 1.0]).astype(np.float32)
    eng, oms = _engine_and_oracles(oracle, cfg, cloud, cyl, poses, 3.25)
    for a in range(4):
        got, ref = eng.agent_map(a).risk_maps(), oms[a].risk_maps()
        assert cases.occupied_set(got) == cases.occupied_set(ref), a
        assert np.array_equal(got, ref), a
        # moving obstacles really moved: slice T-1 differs from slice 0
        assert not np.array_equal(ref[:, 0], ref[:, -1])


@pytest.mark.parametrize("cell", [p.POP_CELL_F32, p.POP_CELL_U8])
def test_peer_projection_bit_exact(oracle, cell):
    n = 6
    cfg = cases.cfg_small_c2(cell, n_local=n, n_total=n)
    cloud, cyl = cases.world_moving(5, (12.0, 12.0, 3.0), 4, 4)
    start, goal = p.worldgen.circle_agents(n, 4.0)
    poses = start[:, :3].astype(np.float32)
    trajs = p.worldgen.straight_trajs(start[:, :3], goal, 9.93, 1.5, 6, 0.2)
    trajs["n_pieces"][2] = 0            # peer that never published
    trajs["start_time"][3] = 10.5       # peer whose trajectory has not started → skipped
    trajs["start_time"][4] = 9.0        # peer whose trajectory ends inside the horizon
    body = p.worldgen.body_table(n, [0.4, 0.4, 0.45])
    eng, oms = _engine_and_oracles(oracle, cfg, cloud, cyl, poses, 10.0, trajs, body)
    total_peer = 0
    for a in range(n):
        got, ref = eng.agent_map(a).risk_maps(), oms[a].risk_maps()
        assert np.array_equal(got, ref), a
        total_peer += int((ref > 1).sum()) + int(ref.sum())
    assert total_peer > 0


def test_queries_match_oracle(oracle):
    cfg = cases.cfg_c1(p.POP_CELL_U8)
    cloud, cyl = cases.world_c1()
    poses = np.array([[0.3, -0.2, 1.0]], np.float32)
    eng, oms = _engine_and_oracles(oracle, cfg, cloud, cyl, poses, 1.0)
    rng = np.random.default_rng(0)
    pos = rng.uniform([-6, -6, -0.5], [6, 6, 3.5], (20000, 3))
    t = rng.integers(0, cfg.nt, 20000).astype(np.int32)
    am = eng.agent_map(0)
    got, ref = am.getClearOcccupancy(pos, t), oms[0].query(pos, t)
    assert np.array_equal(got, ref)
    assert set(np.unique(ref).tolist()) == {-1, 0, 1}
    dt = rng.uniform(0, 2.5, 20000)
    assert np.array_equal(am.getClearOcccupancy(pos, dt), oms[0].query(pos, dt))


def test_queries_riskbase_variant(oracle):
    cfg = cases.cfg_c1(p.POP_CELL_F32)
    cfg.query_variant = p.POP_QUERY_RISKBASE
    cfg.kernel_shape = p.POP_KERNEL_CUBE
    cfg.clearance = 0.3
    cfg.risk_threshold_region = 1.2
    cfg.risk_region_decay = 0.1
    cloud, cyl = cases.world_c1()
    poses = np.array([[0.0, 0.0, 1.0]], np.float32)
    eng, oms = _engine_and_oracles(oracle, cfg, cloud, cyl, poses, 1.0)
    rng = np.random.default_rng(1)
    pos = rng.uniform([-5, -5, -0.5], [5, 5, 3.5], (20000, 3))
    t = rng.integers(0, cfg.nt, 20000).astype(np.int32)
    assert np.array_equal(eng.agent_map(0).getClearOcccupancy(pos, t), oms[0].query(pos, t))


@pytest.mark.parametrize("cell", [p.POP_CELL_F32, p.POP_CELL_U8])
def test_obstacle_points_match_oracle(oracle, cell):
    cfg = cases.cfg_small_c2(cell, n_local=2, n_total=2)
    cloud, cyl = cases.world_moving(9, (12.0, 12.0, 3.0), 10, 10)
    poses = np.array([[0.2, 0.1, 1.0], [-1.0, 2.0, 1.2]], np.float32)
    eng, oms = _engine_and_oracles(oracle, cfg, cloud, cyl, poses, 7.0)
    boxes = [([-2, -2, 0], [2, 2, 2]), ([-20, -20, -5], [20, 20, 9]), ([1.0, 1.0, 0.5], [1.05, 1.05, 0.55]),
             ([3, 3, 0], [2, 2, 1])]
    for a in range(2):
        for (lc, hc) in boxes:
            for (t0, t1) in [(7.0, 7.2), (7.35, 7.61), (6.0, 6.5), (8.5, 12.0)]:
                if (hc[0] - lc[0]) > 30 and t1 - t0 > 1:
                    continue
                got = eng.agent_map(a).getObstaclePoints(t0, t1, lc, hc, cap=400000)
                ref = oms[a].obstacle_points(t0, t1, lc, hc, cap=400000)
                assert got.shape == ref.shape, (a, lc, hc, t0, t1)
                assert np.array_equal(got, ref)


def test_add_particles_and_capacity_errors(oracle):
    cfg = cases.cfg_c1(p.POP_CELL_F32)
    eng = p.PopEngine(cfg)
    with pytest.raises(p.PopError) as e:  # query before the first update
        eng.agent_map(0).getClearOcccupancy([[0, 0, 1]], 0)
    assert e.value.code == -5
    eng.set_world_cloud(np.zeros((0, 3), np.float32))   # empty cloud is legal
    eng.set_cylinders(np.zeros(0, p.CYL_DTYPE))
    eng.set_poses(np.array([[0, 0, 1.0]], np.float32))
    eng.map_update_gt(0.0)
    assert np.count_nonzero(eng.agent_map(0).risk_maps()) == 0
    m = oracle.map(cfg)
    m.set_pose([0, 0, 1.0])
    m.update_gt(np.zeros((0, 3), np.float32), np.zeros(0, p.CYL_DTYPE), 0.0)
    rng = np.random.default_rng(2)
    pts = rng.uniform(-6, 6, (5000, 3)) * [1, 1, 0.3]
    risks = rng.uniform(0, 1, 5000).astype(np.float32)
    eng.agent_map(0).addParticlesToRiskMap(pts, risks, 3)
    m.add_particles(pts, risks, 3)
    got, ref = eng.agent_map(0).risk_maps(), m.risk_maps()
    assert cases.occupied_set(got) == cases.occupied_set(ref)
    # float atomics commute only up to rounding when several particles share a voxel
    assert np.allclose(got, ref, rtol=1e-6, atol=0)
    with pytest.raises(p.PopError) as e:
        eng.set_world_cloud(np.zeros((cfg.max_cloud_points + 1, 3), np.float32))
    assert e.value.code == -4
    pts_few = eng.agent_map(0).getObstaclePoints(0.0, 0.2, [-1, -1, 0], [1, 1, 2])
    with pytest.raises(p.PopError) as e:
        eng.agent_map(0).getObstaclePoints(0.0, 2.0, [-9, -9, -9], [9, 9, 9], cap=1)
    assert e.value.code == -4 or len(pts_few) == 0


def test_full_size_c2_properties():
    """BASELINE config 2 grid (300x300x50, T=20, u8) for 8 agents: size-independent properties —
    slice 0 equals the cropped cloud's voxel set, static world → all slices equal, update is
    idempotent."""
    cfg = cases.cfg_small_c2(p.POP_CELL_U8, n_local=8, n_total=8, nx=300, ny=300, nz=50, nt=20)
    cloud, cyl = cases.world_moving(1236, (30.0, 30.0, 5.0), 40, 0)
    poses = np.zeros((8, 3), np.float32)
    poses[:, 2] = 1.0
    poses[:, 0] = np.linspace(-3, 3, 8)
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    eng.set_poses(poses)
    eng.map_update_gt(0.0)
    r1 = eng.agent_map(3).risk_maps()
    eng.map_update_gt(0.0)
    r2 = eng.agent_map(3).risk_maps()
    assert np.array_equal(r1, r2)
    assert np.count_nonzero(r1[:, 0]) > 10000
    # static obstacles: zero velocity → slice k is the re-deposited corner set; its size never grows
    n0 = np.count_nonzero(r1[:, 0])
    for k in range(1, 20):
        assert 0 < np.count_nonzero(r1[:, k]) <= n0
    # numpy restatement of slice 0 for this agent
    res = np.float32(0.1)
    half = np.array([150, 150, 25], np.float32) * res
    q = cloud - poses[3]
    inr = np.all((q > -half) & (q < half), axis=1)
    idx3 = ((q[inr] + half) / res).astype(np.int32)
    lin = idx3[:, 2] * 300 * 300 + idx3[:, 1] * 300 + idx3[:, 0]
    assert set(np.unique(lin).tolist()) == set(np.nonzero(r1[:, 0])[0].tolist())


def test_full_size_c2_tiled_equals_scatter_path(monkeypatch):
    """BASELINE config 2 grid with MOVING cylinders: the shared-memory tile build (one bulk store per
    z-layer tile) and the clear + scatter kernels must produce byte-identical SOGMs for every agent,
    including agents whose local box sticks out of the arena."""
    n = 6
    cfg = cases.cfg_small_c2(p.POP_CELL_U8, n_local=n, n_total=n, nx=300, ny=300, nz=50, nt=20)
    cloud, cyl = cases.world_moving(77, (30.0, 30.0, 5.0), 10, 30)
    rng = np.random.default_rng(4)
    poses = np.concatenate([rng.uniform(-14, 14, (n, 2)), rng.uniform(0.5, 2.0, (n, 1))], axis=1).astype(np.float32)
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    eng.set_poses(poses)
    monkeypatch.setenv("POP_VOXEL_MODE", "tiles")
    eng.map_update_gt(1.0)
    tiled = [eng.agent_map(a).risk_maps() for a in range(n)]
    monkeypatch.setenv("POP_VOXEL_MODE", "scatter")
    eng.map_update_gt(1.0)
    for a in range(n):
        ref = eng.agent_map(a).risk_maps()
        assert np.array_equal(tiled[a], ref), f"agent {a}"
    assert sum(int(np.count_nonzero(t[:, 19])) for t in tiled) > 1000
    # empty cloud: everything is cleared by the tile path as well
    monkeypatch.setenv("POP_VOXEL_MODE", "tiles")
    eng.set_world_cloud(np.zeros((0, 3), np.float32))
    eng.map_update_gt(1.1)
    assert not eng.agent_map(0).risk_maps().any()


def test_banded_tiles_equal_scatter_path(monkeypatch):
    """A layer larger than one shared-memory tile (400x400 B → two bands of rows): fast cylinders carry
    voxels across the band boundary over the 20 slices; the banded tile build must still equal the
    scatter kernels byte for byte."""
    n = 3
    cfg = cases.cfg_small_c2(p.POP_CELL_U8, n_local=n, n_total=n, nx=400, ny=400, nz=8, nt=20)
    cloud, cyl = worldgen_fast(5, (40.0, 40.0, 0.8), 60)
    poses = np.array([[0.3, -0.2, 0.2], [7.0, 5.0, 0.3], [-15.0, 12.0, 0.1]], np.float32)
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    eng.set_poses(poses)
    monkeypatch.setenv("POP_VOXEL_MODE", "tiles")
    eng.map_update_gt(3.0)
    tiled = [eng.agent_map(a).risk_maps() for a in range(n)]
    monkeypatch.setenv("POP_VOXEL_MODE", "scatter")
    eng.map_update_gt(3.0)
    moved = 0
    for a in range(n):
        ref = eng.agent_map(a).risk_maps()
        assert np.array_equal(tiled[a], ref), f"agent {a}"
        moved += int(np.count_nonzero(ref[:, 19] != ref[:, 0]))
    assert moved > 1000   # the world really moved


def worldgen_fast(seed, arena, n_moving):
    return p.worldgen.make_world(seed, arena=arena, n_static=4, n_moving=n_moving, v_max=3.0, height=0.6)
