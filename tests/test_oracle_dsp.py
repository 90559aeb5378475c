"""CPU checks of the particle-map oracle (oracle/o_dsp.cpp): the reference holds no test for this map
(parity unpinned, see the oracle header), so the restatement is pinned by independent re-derivations
(numpy voxel grid, angular pyramid bins, scipy assignment), by the filter's invariants (weight conservation,
slot capacities) and by its behaviour on scenes with known ground truth."""
import ctypes as C

import numpy as np
import pytest

from pred_occ_planner_b200 import worldgen
from tests import cases


def _make(oracle, seed=5, **kw):
    cfg = cases.dsp_cfg_ref()
    d = cases.dsp_params(cfg, **kw)
    m = oracle.dsp(cfg, d)
    m.set_random_tables(*worldgen.dsp_random_tables(d.n_randoms, seed))
    return cfg, d, m


def test_voxel_filter_against_numpy(oracle):
    cfg, d, m = _make(oracle)
    rng = np.random.default_rng(0)
    cam = rng.uniform([-3, -1.5, 0.2], [3, 1.5, 6.0], (6000, 3)).astype(np.float32)
    cam[17] = np.inf
    got = m.filter(cam)
    # independent restatement: leaf = floor(p / leaf) relative to the cloud minimum, centroids by leaf index
    fin = np.isfinite(cam).all(axis=1)
    p = cam[fin]
    inv = np.float32(1.0) / np.float32(d.voxel_filter_resolution)
    ijk = np.floor(p * inv).astype(np.int64)
    ijk -= ijk.min(axis=0)
    div = ijk.max(axis=0) + 1
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(key, kind="stable")
    uk, first, cnt = np.unique(key[order], return_index=True, return_counts=True)
    cent = np.add.reduceat(p[order].astype(np.float64), first) / cnt[:, None]
    body = np.stack([cent[:, 2], -cent[:, 0], -cent[:, 1]], axis=1)
    half = np.array([cfg.nx, cfg.ny, cfg.nz]) / 2.0 * cfg.resolution
    keep = (np.abs(body) < half).all(axis=1)
    ref = body[keep][:d.max_points]
    assert len(got) == len(ref)
    assert np.allclose(got, ref, atol=2e-6)


def test_pyramid_bins_are_one_degree_cells(oracle):
    """With the identity attitude the plane scan of findPointPyramid*Index equals 1-degree azimuth/elevation bins."""
    cfg, d, m = _make(oracle)
    rng = np.random.default_rng(1)
    az = np.deg2rad(rng.uniform(-42.9, 42.9, 900))
    el = np.deg2rad(rng.uniform(-28.9, 28.9, 900))
    r = rng.uniform(1.0, 4.5, 900)
    # elevation planes contain the y axis: the vertical index is the angle in the x-z plane
    x, y, z = r * np.cos(az), r * np.sin(az), r * np.cos(az) * np.tan(el)
    pts = np.stack([x, y, z], axis=1).astype(np.float32)
    m.update_filtered(pts, [0, 0, 0], [1, 0, 0, 0], 1.0)
    pc, cnt, ml = m.observations()
    assert cnt.sum() == 900
    npv = 58
    hit = 0
    for i in np.nonzero(cnt)[0]:
        ih, iv = i // npv, i % npv
        for j in range(cnt[i]):
            px, py, pz = pc[i, j, :3]
            a, e = np.rad2deg(np.arctan2(py, px)), np.rad2deg(np.arctan2(pz, px))
            # h planes are scanned from -43 deg upward, v planes from -29 deg: sign conventions of :575-589
            hit += (int(np.floor(a + 43)) == ih) and (int(np.floor(29 - e)) == iv or int(np.floor(e + 29)) == iv)
            assert abs(pc[i, j, 4] - np.sqrt(px * px + py * py + pz * pz)) < 1e-5
    assert hit >= 0.98 * 900
    assert (ml[cnt > 0] > 0).all() and (ml[cnt == 0] == -1).all()


def test_assignment_is_optimal(oracle):
    scipy_opt = pytest.importorskip("scipy.optimize")
    rng = np.random.default_rng(2)
    lib = oracle.lib
    for nr, nc in [(1, 1), (3, 5), (6, 4), (12, 12), (20, 31)]:
        cost = rng.uniform(0, 1000, (nr, nc))
        out = np.zeros(nr, np.int32)
        lib.orc_dsp_assign(cost.ctypes.data_as(C.POINTER(C.c_double)), nr, nc, out.ctypes.data_as(C.POINTER(C.c_int32)))
        rows = np.nonzero(out >= 0)[0]
        assert len(rows) == min(nr, nc) and len(set(out[rows].tolist())) == len(rows)
        r, c = scipy_opt.linear_sum_assignment(cost)
        assert abs(cost[rows, out[rows]].sum() - cost[r, c].sum()) < 1e-9


def _run(m, cyl, frames, moving_sensor=True):
    for k in range(frames):
        if moving_sensor:
            pos, quat, st = cases.dsp_flight(k)
        else:
            pos, quat, st = np.array([-3.0, 0.0, 1.0], np.float32), worldgen.quat_yaw(0.0), 50.0 + 0.1 * k
        cam = worldgen.render_depth_cloud(cyl, pos, quat, t=0.1 * k)
        assert m.update(cam, pos, quat, st) == 1
        if k < frames - 1:
            m.export()
    return pos


def test_filter_invariants(oracle):
    cfg, d, m = _make(oracle)
    _run(m, cases.dsp_scene(), 8)
    parts, vox = m.particles(), m.voxels()
    live = parts[:, :, 0] > 0.1
    assert set(np.unique(parts[:, :, 0]).tolist()) <= {0.0, np.float32(0.6), 1.0}
    # every voxel's occupancy weight is the weight of its particles (resampling conserves it)
    wsum = (parts[:, :, 7] * live).sum(axis=1)
    assert np.allclose(wsum, vox[:, 0], rtol=1e-5, atol=1e-7)
    # resampled voxels hold at most MAX_PARTICLE_NUM_VOXEL particles, none lighter than the removal bound
    n = live.sum(axis=1)
    assert n.max() <= d.max_particle_num_voxel
    assert parts[:, :, 7][live].min() >= 1e-3
    assert (parts[:, :, 3][live] == 0).all()          # LIMIT_MOVEMENT_IN_XY_PLANE
    st = m.stats()
    assert st[0] > 0 and st[1] > 0 and st[3] > 0      # full voxels / pyramids dropped particles; newborns added
    # pyramid table: valid entries first, each pointing at a live particle inside the field of view
    pyr = m.pyramids()
    valid = pyr[:, :, 0] & 1
    assert (np.diff(valid, axis=1) <= 0).all() and valid.sum(axis=1).max() == d.pyramid_capacity


def test_static_scene_occupancy_sits_on_the_obstacles(oracle):
    cfg, d, m = _make(oracle)
    _, cyl = worldgen.make_world(3, arena=(8.0, 8.0, 3.0), n_static=5, n_moving=0, height=3.0)
    pos = _run(m, cyl, 10, moving_sensor=False)
    vox = m.voxels()
    idx = np.nonzero(vox[:, 0] > d.occupancy_threshold)[0]
    assert len(idx) > 100
    res = cfg.resolution
    c = np.stack([idx % cfg.nx, (idx // cfg.nx) % cfg.ny, idx // (cfg.nx * cfg.ny)], axis=1) * res \
        - np.array([cfg.nx, cfg.ny, cfg.nz]) * res / 2 + res / 2 + pos
    dist = np.full(len(c), np.inf)
    for cy in cyl:
        dist = np.minimum(dist, np.abs(np.hypot(c[:, 0] - cy["x"], c[:, 1] - cy["y"]) - cy["w"] / 2))
    dist = np.minimum(dist, np.abs(c[:, 2]))          # ground plane z = 0
    assert np.mean(dist < 0.35) > 0.95
    # static world: the predicted slices keep the mass where it is
    risk, nocc = m.export()
    assert nocc == len(idx)
    occ0 = set(np.nonzero(risk[:, 0] > d.occupancy_threshold)[0].tolist())
    occ5 = set(np.nonzero(risk[:, 5] > d.occupancy_threshold)[0].tolist())
    assert len(occ0) > 30 and len(occ0 & occ5) > 0.5 * len(occ0)


def test_moving_obstacle_velocity_and_prediction(oracle):
    cfg, d, m = _make(oracle)
    cyl = np.zeros(1, worldgen.CYL_DTYPE)
    cyl[0] = (3, 0, 0.0, -1.0, 0.0, 0.6, 1.2, 0.0, 1.0, 1, 0, 0, 0)   # short pillar moving along +y at 1 m/s
    vels = []
    for k in range(8):
        pos, quat, st = np.array([-3.0, 0.0, 0.8], np.float32), worldgen.quat_yaw(0.0), 10.0 + 0.1 * k
        cam = worldgen.render_depth_cloud(cyl, pos, quat, t=0.1 * k, ground_z=-5.0)
        assert m.update(cam, pos, quat, st) == 1
        pv = m.points_with_velocity()
        dyn = pv[(pv[:, 3] > 0.01) & (pv[:, 4] > -100)]
        if len(dyn):
            vels.append(dyn[0, 4:7])
        risk, _ = m.export()
    vels = np.array(vels)
    assert len(vels) >= 5
    # centroid differences of voxel-filtered clusters over 0.1 s are quantised (0.15 m leaves): judge the median
    assert abs(np.median(vels[:, 1]) - 1.0) < 0.4 and abs(np.median(vels[:, 0])) < 0.4
    # the predicted mass moves along +y with the slices
    V = cfg.nx * cfg.ny * cfg.nz
    yy = ((np.arange(V) // cfg.nx) % cfg.ny) * cfg.resolution
    com = [(risk[:, k] * yy).sum() / risk[:, k].sum() for k in range(cfg.nt)]
    assert com[5] - com[0] > 0.15   # only the dynamic share of the particle mixture moves


def test_noise_tables_wrap_and_determinism(oracle):
    cfg = cases.dsp_cfg_ref()
    d = cases.dsp_params(cfg, n_randoms=4096)       # far fewer entries than one frame draws
    tabs = worldgen.dsp_random_tables(4096, 9)
    outs = []
    for _ in range(2):
        m = oracle.dsp(cfg, d)
        m.set_random_tables(*tabs)
        _run(m, cases.dsp_scene(), 3)
        outs.append((m.particles(), m.stats()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert 0 <= outs[0][1][4] < 4096


def test_oracle_reproduces_committed_vectors(oracle):
    """tests/golden/oracle_dsp.npz pins the oracle itself (sensor + particle map) against accidental change."""
    import os
    import sys
    from tests import oracle_lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "scripts"))
    from make_golden_dsp import scenario
    g = np.load(os.path.join(root, "tests", "golden", "oracle_dsp.npz"))
    cfg, d, cam, cloud, tabs, frames = scenario()
    m = oracle.dsp(cfg, d)
    m.set_random_tables(*tabs)
    for k, (pos, quat, st) in enumerate(frames):
        depth, pc, _ = oracle_lib.render(cam, pos, quat, cloud)
        if k in (0, 5):
            assert np.array_equal(depth, g[f"depth_{k}"]) and np.array_equal(m.filter(pc), g[f"filtered_{k}"])
        assert m.update(pc, pos.astype(np.float32), quat.astype(np.float32), st) == 1
        risk, nocc = m.export()
        assert np.array_equal(m.stats(), g[f"stats_{k}"]) and nocc == int(g[f"nocc_{k}"])
    parts = m.particles()
    live = g["live_slots"]
    assert np.array_equal(np.argwhere(parts[:, :, 0] > 0.1), live)
    assert np.array_equal(parts[live[:, 0], live[:, 1]], g["live_state"])
    nz = g["risk_idx"]
    assert np.array_equal(np.argwhere(risk != 0), nz) and np.array_equal(risk[nz[:, 0], nz[:, 1]], g["risk_val"])


def test_non_finite_odometry_is_rejected(oracle):
    cfg, d, m = _make(oracle)
    pts = np.array([[2.0, 0.0, 0.5]], np.float32)
    assert m.update_filtered(pts, [0, 0, 1.0], [1, 0, 0, 0], 1.0) == 1
    before = m.particles()
    assert m.update_filtered(pts, [np.nan, 0, 1.0], [1, 0, 0, 0], 1.1) == 0
    assert m.update_filtered(pts, [0, 0, 1.0], [np.nan, 0, 0, 0], 1.1) == 0
    assert m.update_filtered(pts, [0, 0, 1.0], [1, 0, 0, 0], np.inf) == 0
    assert np.array_equal(m.particles(), before)
