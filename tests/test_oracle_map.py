"""CPU checks of the oracle's SOGM restatement on hand-computable cases
(plan_env/src/fake_particle_risk_voxel.cpp:80-170,309-348; plan_env/include/plan_env/map.h:153-215)."""
import numpy as np

import pred_occ_planner_b200 as p
from tests import cases


def test_kernel_sizes_follow_fp32_truncation(oracle):
    # clearance 0.45 / 0.15 in fp32 = 2.9999998 -> inf_step 2 -> 5x5x1 (DISC) / 5x5x5 (CUBE)
    c = cases.cfg_c1()
    assert oracle.map(c).kernel_size() == 25
    c.kernel_shape = p.POP_KERNEL_CUBE
    assert oracle.map(c).kernel_size() == 125
    c.clearance = 0.3
    assert oracle.map(c).kernel_size() == 125  # 0.3f/0.15f = 2.0
    c.clearance = 0.15
    assert oracle.map(c).kernel_size() == 27


def test_single_point_static(oracle):
    c = cases.cfg_c1()
    m = oracle.map(c)
    m.set_pose([0, 0, 1.0])
    pt = np.array([[1.0, 2.0, 1.2]], np.float32)
    m.update_gt(pt, np.zeros(0, p.CYL_DTYPE), 10.0)
    r = m.risk_maps()
    # voxel = int((p - pose + N/2*res)/res) per axis
    res = np.float32(0.15)
    ix = int((np.float32(1.0) + np.float32(33) * res) / res)
    iy = int((np.float32(2.0) + np.float32(33) * res) / res)
    iz = int((np.float32(0.2) + np.float32(10) * res) / res)
    idx = iz * 66 * 66 + iy * 66 + ix
    assert r[idx, 0] == 1.0
    # no cylinder owns it -> zero velocity -> the voxel CORNER is re-deposited in every slice;
    # every slice holds exactly one cell
    assert (np.count_nonzero(r, axis=0) == 1).all()
    # queries: occupied inside the 5x5 disc around it, free 3 voxels away, -1 outside the map / height band
    assert m.query([[1.0, 2.0, 1.2]], 0)[0] == 1
    assert m.query([[1.0 + 0.3, 2.0, 1.2]], 0)[0] == 1
    assert m.query([[1.0 + 0.6, 2.0, 1.2]], 0)[0] == 0
    assert m.query([[1.0, 2.0, 1.2 + 0.15]], 0)[0] == 0  # single z layer
    assert m.query([[20.0, 2.0, 1.2]], 0)[0] == -1
    assert m.query([[1.0, 2.0, 3.5]], 0)[0] == -1
    assert m.query([[1.0, 2.0, -0.5]], 0)[0] == -1
    # double-time overload: floor(dt/0.2) clamped to T-1
    assert m.query([[1.0, 2.0, 1.2]], 0.39)[0] == m.query([[1.0, 2.0, 1.2]], 1)[0]
    assert m.query([[1.0, 2.0, 1.2]], 99.0)[0] == m.query([[1.0, 2.0, 1.2]], 8)[0]


def test_moving_cylinder_prediction(oracle):
    c = cases.cfg_c1()
    m = oracle.map(c)
    m.set_pose([0, 0, 1.0])
    cyl = np.zeros(1, p.CYL_DTYPE)
    cyl[0] = (3, 0, 1.0, 1.0, 0, 0.6, 4.0, 1.0, -0.5, 1, 0, 0, 0)
    pt = np.array([[1.05, 1.05, 1.0]], np.float32)
    m.update_gt(pt, cyl, 0.0)
    r = m.risk_maps()
    v0 = np.nonzero(r[:, 0])[0]
    assert len(v0) == 1
    x0, y0 = v0[0] % 66, (v0[0] // 66) % 66
    v8 = np.nonzero(r[:, 8])[0]
    assert len(v8) == 1
    x8, y8 = v8[0] % 66, (v8[0] // 66) % 66
    # 8 slices * 0.2 s * (1.0, -0.5) m/s = (+1.6, -0.8) m = about (+10.7, -5.3) voxels
    assert x8 - x0 in (10, 11) and y0 - y8 in (5, 6)


def test_obstacle_points_order_and_box(oracle):
    c = cases.cfg_c1()
    m = oracle.map(c)
    m.set_pose([0.5, -0.25, 1.0])
    cloud, cyl = cases.world_c1()
    m.update_gt(cloud, cyl, 5.0)
    pts = m.obstacle_points(5.0, 5.3, [-2, -2, 0.0], [2, 2, 2.0])
    assert len(pts) > 0
    assert (pts[:, 0] >= -2 - 0.151).all() and (pts[:, 0] <= 2 + 1e-6).all()
    # emission order: z outer, then y, then x (slices innermost -> duplicates adjacent)
    key = np.round(pts[:, 2] / 0.15) * 1e6 + np.round(pts[:, 1] / 0.15) * 1e3 + np.round(pts[:, 0] / 0.15)
    assert (np.diff(key) >= 0).all()


def test_peer_projection_counts(oracle):
    c = cases.cfg_c1(n_total=3)
    m = oracle.map(c)
    m.set_pose([0, 0, 1.0])
    m.update_gt(np.zeros((0, 3), np.float32), np.zeros(0, p.CYL_DTYPE), 100.0)
    offs, counts = p.worldgen.body_table(3, [0.4, 0.4, 0.45], ego_id=0)
    start = np.array([[0, 0, 1.0], [1.0, 1.0, 1.0], [-2.0, 0.5, 1.2]])
    goal = np.array([[3, 0, 1.0], [-1.0, 1.0, 1.0], [2.0, 0.5, 1.2]])
    recs = p.worldgen.straight_trajs(start, goal, 99.95, 1.0, 4, 0.3)
    m.add_other_agents(offs, counts, recs, 0)
    r = m.risk_maps()
    nb = counts[0]
    # both peers are active (t_start < t < t_end) for slices 0..5 (t_end = 99.95+1.2 > 100+0.2k for k<=5)
    for k in range(9):
        expect = 2 * nb if 100.0 + np.float32(0.2) * k < 99.95 + 4 * np.float32(0.3) else 0
        assert r[:, k].sum() == expect, k
