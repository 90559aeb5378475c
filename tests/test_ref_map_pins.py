"""Pins of the oracle's maps against the REFERENCE's own map classes.

`oracle/_ref/libref_map_<variant>.so` = plan_env/src/map.cpp + fake_particle_risk_voxel.cpp (ground-truth SOGM) or
risk_base.cpp (RiskBase), with traj_coordinator/src/particles.cpp and traj_utils/src/bernstein.cpp, compiled unmodified
from /root/reference at one grid size per library (oracle/Makefile, oracle/ref_glue_map.cpp; ROS/PCL inert stand-ins,
Eigen's fp32 vectors / quaternion / plane restated in oracle/ref_build/Eigen).  Everything here is integer or fp32
index work, so every comparison is EXACT:

  * the SOGM after updateMap (cylinders, moving cylinders, vertical rings, one peer's projected trajectory) and after
    futureRiskCallback + addOtherAgents — every cell of [V][T];
  * getClearOcccupancy(pos, int t) and (pos, double dt); getObstaclePoints(box) in emission order;
    getInflateOccupancy in its three forms; addParticlesToRiskMap / addObstaclesToRiskMap; the voxel index math;
  * always against the committed outputs of those libraries (tests/golden/ref_map_pins.npz,
    scripts/make_golden_ref_map.py), and against the libraries themselves on wider sweeps where they exist.

Not pinned, because the reference itself is undefined there: two or more peers (getParticlesWithRisk clears the shared
`risks` vector per peer while `particles` accumulates, fake_particle_risk_voxel.cpp:196-201 — the second peer's
particles read past the vector's size), and slice indices equal to PREDICTION_TIMES
(map.cpp:366,470-475 read one slice past the end, Q2)."""
import ctypes as C
import os

import numpy as np
import pytest

import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import worldgen
from tests import cases
from tests import ref_map_lib as rml

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_map_pins.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _same(got, gold, name):
    for k, v in got.items():
        w = gold[f"{name}.{k}"]
        assert v.shape == w.shape and np.array_equal(v, w), (name, k, v.shape, w.shape)


@pytest.mark.parametrize("name", rml.MAP_CASES)
def test_oracle_equals_reference_vectors(oracle, gold, name):
    """The oracle on every pinned scene: grid, queries, obstacle points, inflate queries identical to what the
    compiled reference produced."""
    case = rml.map_case(name)
    got = rml.run_oracle(oracle, case)
    _same(got, gold, name)
    assert len(got["nz"]) > 1000 and (got["q_t"] == 1).sum() > 50 and (got["q_t"] == 0).sum() > 500
    assert len(got["obs0"]) > 100 and len(got["obs1"]) > 100


@pytest.mark.parametrize("name", rml.MAP_CASES)
def test_golden_file_is_current(gold, name):
    """Where the reference libraries can be built, the committed vectors are what they produce today."""
    case = rml.map_case(name)
    got = rml.run_reference(case)
    if got is None:
        pytest.skip("oracle/_ref/libref_map_*.so not present (no /root/reference here)")
    _same(got, gold, name)


def test_reference_footprints():
    """init() of both classes: FakeParticleRiskVoxel's `for (int z = -0.3; ...)` collapses to one layer and 0.45f / 0.15f
    truncates to 2 (5 x 5 x 1); RiskBase builds the full cube."""
    lib = rml.load("fake_c1")
    if lib is None:
        pytest.skip("reference map libraries not present")
    k = rml.RefMap(lib, cases.cfg_c1(p.POP_CELL_F32)).kernel()
    assert len(k) == 25 and (k[:, 2] == 0).all() and k[:, 0].min() == -2 and k[:, 0].max() == 2
    k = rml.RefMap(rml.load("risk_ref"), cases.dsp_cfg_ref()).kernel()
    assert len(k) == 27 and np.array_equal(k[0], [-1, -1, -1]) and np.array_equal(k[1], [-1, -1, 0])   # x outer, z inner


def test_ground_truth_sweep_live(oracle):
    """Wider sweep against the live library: cylinders + rings in every id order, poses off the voxel lattice."""
    lib = rml.load("fake_c1")
    if lib is None:
        pytest.skip("reference map libraries not present")
    cfg = cases.cfg_c1(p.POP_CELL_F32)
    ring_cells = 0
    for seed in range(8):
        rng = np.random.default_rng(seed)
        cloud, cyl = cases.world_moving(seed, (10., 10., 3.), 4, 6)
        rc, rings = worldgen.make_rings(100 + seed, (10., 10., 3.), n=3)
        cloud = np.concatenate([cloud, rc])
        allc = np.concatenate([rings[:1], cyl, rings[1:]])
        pose = rng.uniform([-3, -3, 0.5], [3, 3, 1.5])
        rm = rml.RefMap(lib, cfg)
        rm.set_pose(pose)
        rm.set_cylinders(allc)
        rm.update(cloud, 5.0)
        om = oracle.map(cfg)
        om.set_pose(pose.astype(np.float32))
        om.update_gt(cloud, allc, 5.0)
        a, b = rm.risk_maps(), om.risk_maps()
        assert np.array_equal(a, b), seed
        om2 = oracle.map(cfg)
        om2.set_pose(pose.astype(np.float32))
        om2.update_gt(cloud, cyl, 5.0)
        ring_cells += np.count_nonzero(b != om2.risk_maps())
    assert ring_cells > 2000   # the rings did own (and move) voxels


def test_riskbase_and_deposits_live(oracle):
    """RiskBase: grid after futureRiskCallback, region queries with the decaying threshold, box scan with the decaying
    voxel threshold; addParticlesToRiskMap / addObstaclesToRiskMap; one peer in each of the trajectory states."""
    lib = rml.load("risk_ref")
    if lib is None:
        pytest.skip("reference map libraries not present")
    cfg = cases.dsp_cfg_ref()
    cfg.n_agents_total = 2
    V = cfg.nx * cfg.ny * cfg.nz
    for seed in range(6):
        rng = np.random.default_rng(40 + seed)
        risk = np.zeros((V, cfg.nt), np.float32)
        idx = rng.integers(0, V, 15000)
        risk[idx, :] = rng.uniform(0, 0.6, (len(idx), cfg.nt)).astype(np.float32)
        pose = rng.uniform([-1, -1, 0.8], [1, 1, 1.3])
        t_map = 30.0
        ego = seed % 2
        # peer states across the slices: not started, running, finishing inside the horizon, finished
        t_start = [t_map + d for d in ((0.4, -0.2), (-0.2, -0.9), (-1.4, -0.1), (-3.0, -4.0), (-0.5, 0.7), (0.0, 0.0))[seed]]
        recs = rml.curved_trajs(rng, 2, t_start, 1 + seed % 4, 0.3)
        offs, counts = worldgen.body_table(2, rml.BODY_SIZE, ego_id=ego)
        rm = rml.RefMap(lib, cfg)
        rm.set_pose(pose)
        rm.set_swarm(2, ego, rml.BODY_SIZE, offs, counts, recs, t_map)
        rm.future_risk(risk, t_map)
        om = oracle.map(cfg)
        om.set_pose(pose.astype(np.float32))
        om.set_risk(risk)
        om.set_time(t_map)
        om.add_other_agents(offs, counts, recs, ego)
        assert np.array_equal(rm.risk_maps(), om.risk_maps()), seed
        pos = pose + rng.uniform(-5.5, 5.5, (8000, 3)) * [1, 1, 0.35]
        tt = rng.integers(0, cfg.nt, 8000).astype(np.int32)
        assert np.array_equal(rm.query(pos, tt), om.query(pos, tt))
        dt = rng.uniform(0, 2.5, 8000)
        assert np.array_equal(rm.query(pos, dt), om.query(pos, dt))
        box = (t_map + 0.1, t_map + 1.3, pose - [2, 2, 1], pose + [2, 2, 1])
        assert np.array_equal(rm.obstacle_points(*box), om.obstacle_points(*box))
        pts = rng.uniform(-5.2, 5.2, (4000, 3)) * [1, 1, 0.31]
        rk = rng.uniform(0, 1, 4000).astype(np.float32)
        rm.add_particles(pts, rk, 2)
        om.add_particles(pts, rk, 2)
        assert np.array_equal(rm.risk_maps(), om.risk_maps())


def test_index_math_live():
    """map.h:153-215 on the reference's own grid: isInRange, getVoxelIndex (fp32 add, fp32 divide, truncation),
    getVoxelPosition — against the product's device geometry compiled for the host (tests/csrc/geom_host.cpp)."""
    lib = rml.load("fake_c1")
    if lib is None:
        pytest.skip("reference map libraries not present")
    from tests.test_device_logic_on_host import _build, HERE
    dev = _build(os.path.join(HERE, "csrc", "geom_host.cpp"), os.path.join(HERE, "_build", "libdevgeom.so"),
                 ["pop_geom.cuh", "pop_internal.h"])
    dev.dev_axis_position.restype = C.c_float
    cfg = cases.cfg_c1(p.POP_CELL_F32)
    rm = rml.RefMap(lib, cfg)
    pose = np.array([0.37, -1.21, 1.03])
    rm.set_pose(pose)
    rng = np.random.default_rng(5)
    local = rng.uniform(-5.2, 5.2, (20000, 3)).astype(np.float32) * np.array([1, 1, 0.31], np.float32)
    # points a hair inside the borders and exactly on voxel faces
    local[:2000] = (np.round(local[:2000] / np.float32(0.15)) * np.float32(0.15)).astype(np.float32)
    idx = rng.integers(0, cfg.nx * cfg.ny * cfg.nz, 5000).astype(np.int32)
    ref_idx, ref_in, ref_pos = rm.index_math(local, idx)
    posef = pose.astype(np.float32)
    for i in range(0, len(local), 7):
        if ref_in[i]:
            assert dev.dev_voxel_index(C.byref(cfg), C.c_float(local[i, 0]), C.c_float(local[i, 1]), C.c_float(local[i, 2])) == ref_idx[i]
    for j in range(0, len(idx), 5):
        v = int(idx[j])
        ijk = (v % cfg.nx, (v // cfg.nx) % cfg.ny, v // (cfg.nx * cfg.ny))
        for ax in range(3):
            assert dev.dev_axis_position(C.byref(cfg), ijk[ax], ax, C.c_float(posef[ax])) == ref_pos[j, ax]


def test_device_ownership_source_equals_reference_vectors(gold):
    """The product's obstacle-ownership rule (pop_geom.cuh: geom_cyl_prepare + geom_gt_owner, what k_vox_fill and
    k_future_slices call) compiled for the host: rebuilding the future slices of a pinned scene with cylinders AND rings
    from its slice-0 voxels gives exactly the reference's cells."""
    from tests.test_device_logic_on_host import _build, HERE
    dev = _build(os.path.join(HERE, "csrc", "geom_host.cpp"), os.path.join(HERE, "_build", "libdevgeom.so"),
                 ["pop_geom.cuh", "pop_internal.h"])
    name = "c1_pillars_rings"
    case = rml.map_case(name)
    cfg = case["cfg"]
    V, T = cfg.nx * cfg.ny * cfg.nz, cfg.nt
    nz = gold[f"{name}.nz"].astype(np.int64)
    vox0 = np.unique(nz[nz % T == 0] // T)
    f32 = np.float32
    pose = case["pose"].astype(f32)
    res, dt = f32(cfg.resolution), f32(cfg.time_resolution)
    half = np.array([cfg.nx // 2, cfg.ny // 2, cfg.nz // 2]).astype(f32) * res
    ijk = np.stack([vox0 % cfg.nx, (vox0 // cfg.nx) % cfg.ny, vox0 // (cfg.nx * cfg.ny)], axis=1)
    corners = ((ijk.astype(f32) * res - half) + pose).astype(f32)
    own = np.zeros(len(vox0), np.int32)
    cyl = np.ascontiguousarray(case["cyl"])
    dev.dev_gt_owner(C.byref(cfg), C.c_void_p(cyl.ctypes.data), len(cyl), corners.ctypes.data_as(C.POINTER(C.c_float)),
                     len(vox0), own.ctypes.data_as(C.POINTER(C.c_int32)))
    assert (cyl["type"][own[own >= 0]] == 2).sum() > 100 and (cyl["type"][own[own >= 0]] == 3).sum() > 100
    vel = np.zeros((len(vox0), 3), f32)
    vel[own >= 0, 0] = cyl["vx"][own[own >= 0]].astype(f32)
    vel[own >= 0, 1] = cyl["vy"][own[own >= 0]].astype(f32)
    cells = {int(v) * T for v in vox0}
    for k in range(1, T):
        q = ((corners + (vel * dt) * f32(k)) - pose).astype(f32)
        inr = ((q > -half) & (q < half)).all(axis=1)
        qi = ((q + half) / res).astype(np.int32)
        idx = qi[:, 2].astype(np.int64) * cfg.nx * cfg.ny + qi[:, 1] * cfg.nx + qi[:, 0]
        ok = inr & (idx >= 0) & (idx < V)
        cells.update((idx[ok] * T + k).tolist())
    assert cells == set(nz.tolist())


def test_peer_resampling_live(oracle):
    """swarm/replan_risk_rate > 0 (particles.cpp:365-408): num_resample Gaussian copies per peer particle with
    density weights normalised per particle.  The reference seeds its generator with time(NULL) in every call; the
    reference library answers time() itself, and the oracle is fed the sequence libstdc++ produces for that value
    (three distributions over one engine).  Deposited grid identical, cell for cell."""
    import ctypes as C
    from tests import oracle_lib
    lib = rml.load("fake_c1")
    if lib is None:
        pytest.skip("reference map libraries not present")
    cfg = cases.cfg_c1(p.POP_CELL_F32)
    cfg.n_agents_total, cfg.replan_risk_rate, cfg.num_resample = 2, 0.1, 10
    cells = 0
    try:
        for seed in range(6):
            clock = 1000 + 77 * seed
            lib.refmap_set_time_seed(C.c_long(clock))
            oracle_lib.set_resample_noise(oracle_lib.libstdcxx_normal_triples(clock, 64 * cfg.num_resample))
            case = rml.resample_case(seed)
            rm = rml.RefMap(lib, cfg)
            rm.set_pose(case["pose"])
            rm.set_swarm(2, case["ego"], rml.BODY_SIZE, case["offs"], case["counts"], case["recs"], case["t_map"],
                         replan_risk_rate=cfg.replan_risk_rate, num_resample=cfg.num_resample)
            rm.set_cylinders(np.zeros(0, p.CYL_DTYPE))
            rm.update(np.zeros((0, 3), np.float32), case["t_map"])
            om = oracle.map(cfg)
            om.set_pose(case["pose"].astype(np.float32))
            om.update_gt(np.zeros((0, 3), np.float32), np.zeros(0, p.CYL_DTYPE), case["t_map"])
            om.add_other_agents(case["offs"], case["counts"], case["recs"], case["ego"])
            a, b = rm.risk_maps(), om.risk_maps()
            assert np.array_equal(a, b), seed
            cells += np.count_nonzero(a)
    finally:
        oracle_lib.set_resample_noise(np.zeros((0, 3), np.float32))
    assert cells > 2000
