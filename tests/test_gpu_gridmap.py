"""GPU parity of the raycast map (log-odds GridMap) against the sequential oracle: occupancy_buffer_ (doubles),
occupancy_buffer_inflate_ and the local bounds bit-exact, frame after frame, including the frames where the
reference's char frame counter wraps (128) and meets stale per-voxel stamps (255 on)."""
import numpy as np
import pytest

import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import gridmap, worldgen
from tests import oracle_lib

pytestmark = pytest.mark.gpu


def _cfg():
    c = gridmap.default_gridmap_config()
    c.map_size[0], c.map_size[1], c.map_size[2] = 20.0, 20.0, 4.0
    c.local_map_margin = 5
    c.max_points = 20000
    return c


def _scene():
    cloud, cyl = worldgen.make_world(11, arena=(10.0, 10.0, 3.0), n_static=8, n_moving=0, height=3.0)
    gx, gy = np.meshgrid(np.arange(-5, 5, 0.1), np.arange(-5, 5, 0.1), indexing="ij")
    ground = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1).astype(np.float32)
    return np.concatenate([cloud, ground]).astype(np.float32)


def _frame(cam, cloud, a, k):
    pos = np.array([-4.0 + 0.15 * k, -1.0 + 1.0 * a, 1.0 + 0.1 * a])
    yaw = 0.1 * a + 0.05 * k
    quat = np.array([np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)])
    _, _, pw = oracle_lib.render(cam, pos, quat, cloud)
    return pos, pw.astype(np.float64)


def _same(maps, m, orc):
    occ_g, inf_g, b_g = maps.buffers(m)
    occ_o, inf_o, b_o = orc.buffers()
    bad = []
    if not np.array_equal(b_g, b_o):
        bad.append(f"bounds {b_g.tolist()} vs {b_o.tolist()}")
    if not np.array_equal(occ_g, occ_o):
        bad.append(f"log-odds differ in {np.count_nonzero(occ_g != occ_o)} voxels")
    if not np.array_equal(inf_g, inf_o):
        bad.append(f"inflated map differs in {np.count_nonzero(inf_g != inf_o)} voxels")
    return bad


def test_depth_camera_frames_bit_exact():
    n = 2
    c = _cfg()
    cam = p._lib.PopCamera()
    p.load().pop_camera_default(p._lib.C.byref(cam), 160, 120)
    cloud = _scene()
    maps = gridmap.GridMaps(c, n_maps=n)
    orcs = [oracle_lib.OracleGridMap(c) for _ in range(n)]
    for k in range(6):
        clouds, cams = [], []
        for a in range(n):
            pos, pw = _frame(cam, cloud, a, k)
            if k == 3 and a == 1:
                pw = np.concatenate([pw, [[30.0, 0.0, 1.0], [0.0, -40.0, 0.5], [9.0, 9.0, 3.5]]])   # outside / far points
            clouds.append(pw)
            cams.append(pos)
            orcs[a].update(pw, pos)
        maps.updateOccupancy(clouds, np.array(cams))
        for a in range(n):
            bad = _same(maps, a, orcs[a])
            assert not bad, f"frame {k} map {a}: " + "; ".join(bad)
    occ, inf, _ = maps.buffers(0)
    assert (occ > np.log(c.p_occ / (1 - c.p_occ))).sum() > 200 and inf.sum() > 1000
    # queries
    rng = np.random.default_rng(0)
    q = rng.uniform([-11, -11, -0.5], [11, 11, 4.5], (5000, 3))
    o, i = maps.getOccupancy(0, q), maps.getInflateOccupancy(0, q)
    org = np.array([-10.0, -10.0, c.ground_height])
    inside = np.all((q >= org + 1e-4) & (q <= org + np.array([20.0, 20.0, 4.0]) - 1e-4), axis=1)
    idx = np.floor((q[inside] - org) / c.resolution).astype(int)
    assert (o[~inside] == -1).all() and (i[~inside] == -1).all()
    assert np.array_equal(o[inside], (occ[idx[:, 0], idx[:, 1], idx[:, 2]] > np.log(c.p_occ / (1 - c.p_occ))).astype(np.int8))
    assert np.array_equal(i[inside], inf[idx[:, 0], idx[:, 1], idx[:, 2]])
    maps.close()


def test_parity_across_the_counter_wrap():
    """300 small frames.  The reference's frame counter is a `char` (grid_map.h:140) like the per-voxel stamps it is
    compared with: it wraps 127 -> -128 at frame 128, equals the stamps' initial -1 at frame 255 (every never-touched
    voxel then reads as already traversed / already a ray end) and meets 256-frame-old stamps from frame 257 on.  The
    oracle is pinned to the compiled reference across all of that (tests/test_ref_gridmap_pins.py); here the device
    follows the oracle.  The second map receives no points on every third frame, so its counter (which only advances
    on frames with points, grid_map.cpp:316-320) reaches those frames later than the first map's."""
    c = _cfg()
    maps = gridmap.GridMaps(c, n_maps=2)
    orcs = [oracle_lib.OracleGridMap(c), oracle_lib.OracleGridMap(c)]
    rng = np.random.default_rng(4)
    base = rng.uniform([1.0, -2.0, 0.2], [4.0, 2.0, 2.2], (300, 3))
    check = {0, 60, 126, 127, 128, 134, 189, 190, 191, 192, 199, 253, 254, 255, 256, 257, 258, 270, 299}
    for k in range(300):
        pos = np.array([-1.0 + 0.005 * k, 0.0, 1.0])
        pts = np.concatenate([base + rng.normal(0, 0.02, base.shape), base[:40]])   # repeated end voxels
        second = pts[::2] if k % 3 else np.zeros((0, 3))
        maps.updateOccupancy([pts, second], np.stack([pos, pos]))
        orcs[0].update(pts, pos)
        orcs[1].update(second, pos)
        if k in check:   # map 0: frame = k + 1; map 1 reaches its frame 128 at k = 191
            for m in range(2):
                bad = _same(maps, m, orcs[m])
                assert not bad, f"frame {k} map {m}: " + "; ".join(bad)
    maps.close()


def test_empty_frames_and_errors():
    c = _cfg()
    maps = gridmap.GridMaps(c, n_maps=2)
    maps.updateOccupancy([np.zeros((0, 3)), np.zeros((0, 3))], np.zeros((2, 3)) + [0, 0, 1.0])
    occ, inf, _ = maps.buffers(1)
    assert (occ < np.log(c.p_min / (1 - c.p_min))).all() and not inf.any()    # all unknown
    with pytest.raises(p.PopError):
        maps.updateOccupancy([np.zeros((1, 3)), np.zeros((1, 3))], np.array([[0, 0, 1.0], [50.0, 0, 1.0]]))   # camera outside
    maps.close()


def test_depth_image_pipeline_bit_exact():
    """projectDepthImage on the device (ordered compaction, the reference's pointer quirks) feeding the raycast:
    projected points and all map buffers equal the oracle's, with and without the depth filter."""
    cam = p._lib.PopCamera()
    p.load().pop_camera_default(p._lib.C.byref(cam), 160, 120)
    cloud = _scene()
    n = 2
    for use_filter in (True, False):
        c = _cfg()
        maps = gridmap.GridMaps(c, n_maps=n)
        P = gridmap.depth_params(cam.height, cam.width, cam.fx, cam.fy, cam.cx, cam.cy, use_filter=use_filter,
                                 skip_pixel=2, margin=2, mindist=0.2, maxdist=5.0)
        maps.setDepthParams(P)
        orcs = [oracle_lib.OracleGridMap(c) for _ in range(n)]
        has_first = [False] * n
        # cam02body of the simulator: camera z forward = body x, camera x = -body y, camera y = -body z
        c2b = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
        for k in range(4):
            imgs, Rs, cams = [], [], []
            for a in range(n):
                pos = np.array([-4.0 + 0.15 * k, -1.0 + 1.0 * a, 1.0 + 0.1 * a])
                yaw = 0.1 * a + 0.05 * k
                quat = np.array([np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)])
                depth, _, _ = oracle_lib.render(cam, pos, quat, cloud)
                img = np.where(depth >= 65535, 0, depth).astype(np.uint16)     # empty pixels -> 0, millimetres
                Rb = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
                imgs.append(img); Rs.append(Rb @ c2b); cams.append(pos)
            maps.updateOccupancyFromDepth(np.stack(imgs), np.stack(Rs), np.stack(cams))
            for a in range(n):
                pts, has_first[a] = oracle_lib.gridmap_project(P, has_first[a], imgs[a], Rs[a], cams[a], c.max_ray_length)
                got = maps.projectedPoints(a)
                assert got.shape == pts.shape and np.array_equal(got, pts), (use_filter, k, a)
                orcs[a].update(pts, cams[a])
                bad = _same(maps, a, orcs[a])
                assert not bad, f"filter {use_filter} frame {k} map {a}: " + "; ".join(bad)
        occ, inf, _ = maps.buffers(0)
        assert (occ > np.log(c.p_min / (1 - c.p_min)) + 0.5).sum() > 50     # voxels that collected hits
        maps.close()
