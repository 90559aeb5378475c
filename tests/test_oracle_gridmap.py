"""CPU checks of the raycast-map oracle (oracle/o_gridmap.cpp) on scenes with known geometry."""
import numpy as np

import pred_occ_planner_b200 as p
from tests import oracle_lib


def _cfg(**kw):
    c = p.PopGridmapConfig()
    p.load().pop_gridmap_default_config(p._lib.C.byref(c))
    c.map_size[0], c.map_size[1], c.map_size[2] = 20.0, 20.0, 4.0
    c.local_map_margin = 5
    c.max_points = 8192
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _wall(x=3.0, n=40):
    ys, zs = np.meshgrid(np.linspace(-1.5, 1.5, n), np.linspace(0.3, 2.0, n), indexing="ij")
    return np.stack([np.full(ys.size, x), ys.ravel(), zs.ravel()], axis=1)


def test_wall_becomes_occupied_and_the_space_in_front_free():
    c = _cfg()
    m = oracle_lib.OracleGridMap(c)
    cam = np.array([0.0, 0.0, 1.0])
    lo = lambda q: np.log(q / (1 - q))
    for _ in range(6):
        m.update(_wall(), cam)
    occ, inf, b = m.buffers()
    res, org = c.resolution, np.array([-10.0, -10.0, c.ground_height])
    idx = lambda q: tuple(np.floor((np.asarray(q) - org) / res).astype(int))
    assert occ[idx([3.0, 0.0, 1.0])] > lo(c.p_occ)                 # the wall
    assert abs(occ[idx([1.5, 0.0, 1.0])] - lo(c.p_min)) < 1e-12     # carved free space, clamped
    assert occ[idx([5.0, 0.0, 1.0])] < lo(c.p_min) - 1e-3           # behind the wall: still unknown
    assert inf[idx([3.0, 0.0, 1.0])] == 1 and inf[idx([2.9, 0.0, 1.0])] == 1 and inf[idx([1.5, 0.0, 1.0])] == 0
    # six hits from "unknown" (clamp_min - 0.01): log-odds accumulate and clamp
    assert abs(occ[idx([3.0, 0.0, 1.0])] - min(lo(c.p_min) - 0.01 + 6 * lo(c.p_hit), lo(c.p_max))) < 1e-9
    assert (b[:3] <= idx(cam)).all() and (b[3:] >= idx([3.0, 1.4, 1.9])).all()
    # virtual ceiling inside the local bounds
    ceil_id = int(np.floor((c.virtual_ceil_height - c.ground_height) / res)) - 1
    assert inf[idx(cam)[0], idx(cam)[1], ceil_id] == 1


def test_far_and_outside_points_only_carve():
    c = _cfg()
    m = oracle_lib.OracleGridMap(c)
    cam = np.array([0.0, 0.0, 1.0])
    pts = np.array([[8.0, 0.0, 1.0],      # beyond max_ray_length: clipped to 4.5 m, counted as a miss
                    [30.0, 0.0, 1.0]])    # outside the map: clipped to the boundary, then to 4.5 m
    m.update(pts, cam)
    occ, inf, _ = m.buffers()
    assert not (occ > np.log(c.p_occ / (1 - c.p_occ))).any() and not inf[:, :, :20].any()
    free = np.isclose(occ, np.log(c.p_min / (1 - c.p_min)))
    assert 40 <= free.sum() <= 60                                   # one 4.5 m ray of 0.1 m voxels


def test_rays_of_a_frame_stop_on_earlier_rays():
    """Two points behind each other: the second ray stops where it meets the first one's voxels (counting one
    more miss there).  The run continues past frame 127, where the reference's vector<char> flags stop matching
    the int frame counter and every point casts a full ray (the GPU parity test compares across that wrap)."""
    c = _cfg()
    cam = np.array([0.0, 0.05, 1.05])
    pts = np.array([[3.05, 0.05, 1.05], [2.05, 0.05, 1.05]])
    m = oracle_lib.OracleGridMap(c)
    m.update(pts, cam)
    occ1 = m.buffers()[0].copy()
    org = np.array([-10.0, -10.0, c.ground_height])
    idx = lambda q: tuple(np.floor((np.asarray(q) - org) / c.resolution).astype(int))
    lo = lambda q: np.log(q / (1 - q))
    unk = lo(c.p_min) - 0.01
    # far voxel: 1 hit + 1 miss (its own ray starts there) -> tie -> hit.  Near voxel: 1 hit, missed by the first
    # ray and by its own ray's first step (which then stops: already traversed) -> miss.  A voxel between the
    # camera and the near point is traversed by the first ray only: one miss.
    assert abs(occ1[idx(pts[0])] - (unk + lo(c.p_hit))) < 1e-12
    assert abs(occ1[idx(pts[1])] - lo(c.p_min)) < 1e-12
    assert abs(occ1[idx([1.05, 0.05, 1.05])] - lo(c.p_min)) < 1e-12
    for _ in range(140):
        m.update(pts, cam)
    occ2 = m.buffers()[0]
    assert occ2[idx(pts[0])] == lo(c.p_max) and abs(occ2[idx(pts[1])] - lo(c.p_min)) < 1e-12


def test_depth_projection():
    """projectDepthImage: pinhole back-projection in (v, u) order; the first filtered frame only arms the filter;
    too-near pixels are skipped, too-far and (next-sample) zero pixels become max-range free rays."""
    from pred_occ_planner_b200 import gridmap
    rows, cols = 24, 32
    P = gridmap.depth_params(rows, cols, 30.0, 30.0, 16.0, 12.0, use_filter=True, skip_pixel=2, margin=2,
                             mindist=0.2, maxdist=5.0)
    img = np.full((rows, cols), 3000, np.uint16)
    img[6, 10] = 100        # nearer than mindist: skipped
    img[8, 12] = 9000       # farther than maxdist: max_ray + 0.1
    img[10, 16] = 0         # zero at a sampled pixel: the PREVIOUS sample (u = 14) is the one turned into a free ray
    R = np.eye(3)
    cam = np.array([0.0, 0.0, 1.0])
    pts, hf = oracle_lib.gridmap_project(P, False, img, R, cam, 4.5)
    assert len(pts) == 0 and hf                               # first frame
    pts, hf = oracle_lib.gridmap_project(P, True, img, R, cam, 4.5)
    nu, nv = len(range(2, cols - 2, 2)), len(range(2, rows - 2, 2))
    assert len(pts) == nu * nv - 2                            # the too-near pixel and the zero pixel itself (depth 0)
    z = pts[:, 2] - 1.0
    assert np.isclose(z, 3.0).sum() == nu * nv - 4 and np.isclose(z, 4.6).sum() == 2   # 9000 and the sample before the zero
    # first sample: u = v = 2
    assert np.allclose(pts[0], [(2 - 16.0) * 3.0 / 30.0, (2 - 12.0) * 3.0 / 30.0, 4.0])
    # without the filter the row pointer advances by one per sample (skip_pixel = 2 reads columns 0, 1, 2, ...)
    P2 = gridmap.depth_params(rows, cols, 30.0, 30.0, 16.0, 12.0, use_filter=False, skip_pixel=2)
    img2 = np.tile(np.arange(cols, dtype=np.uint16) * 100 + 1000, (rows, 1))
    pts2, _ = oracle_lib.gridmap_project(P2, False, img2, R, cam, 4.5)
    assert len(pts2) == (cols // 2) * (rows // 2)
    assert np.allclose(pts2[:4, 2] - 1.0, [1.0, 1.1, 1.2, 1.3])
