"""The raycast-map oracle (oracle/o_gridmap.cpp: projectDepthImage, raycastProcess, clearAndInflateLocalMap) against the
reference's own plan_env/src/grid_map.cpp, compiled unmodified into oracle/_ref/libref_gridmap.so: projected points,
log-odds doubles, inflated map and local bounds bit for bit, frame after frame, with and without the depth filter.
Committed vectors (tests/golden/ref_gridmap_pins.npz, scripts/make_golden_ref_gridmap.py) make the same comparison
where the reference tree is absent; the live test also crosses frames 128 and 255-257, where the reference's `char`
frame counter wraps and then meets the initial value (-1) and the 256-frame-old values of its per-voxel flags."""
import os

import numpy as np
import pytest

from pred_occ_planner_b200 import gridmap
from tests import oracle_lib
from tests import ref_gridmap_lib as rg
from tests.conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "ref_gridmap_pins.npz")


def _dp(cam, use_filter):
    return gridmap.depth_params(cam.height, cam.width, cam.fx, cam.fy, cam.cx, cam.cy, use_filter=bool(use_filter),
                                skip_pixel=2, margin=2, mindist=0.2, maxdist=5.0)


@pytest.mark.parametrize("use_filter", [1, 0])
def test_oracle_equals_recorded_reference_frames(use_filter):
    g = np.load(GOLD)
    cam, cloud, cfg = rg.camera(), rg.scene(), rg.config()
    dp = _dp(cam, use_filter)
    orc = oracle_lib.OracleGridMap(cfg)
    has_first = False
    sample = g[f"f{use_filter}_sample"]
    for k in range(rg.N_FRAMES):
        img, R, pos = rg.frame(cam, cloud, k)
        pts, has_first = oracle_lib.gridmap_project(dp, has_first, img, R, pos, cfg.max_ray_length)
        want = g[f"f{use_filter}_k{k}_pts"]
        assert pts.shape == want.shape and np.array_equal(pts, want), (use_filter, k)
        orc.update(pts, pos)
        occ, inf, b = orc.buffers()
        assert np.array_equal(b, g[f"f{use_filter}_k{k}_bounds"])
        assert np.array_equal(occ.ravel()[sample], g[f"f{use_filter}_k{k}_occ_sample"])
        assert rg.digest(occ) == bytes(g[f"f{use_filter}_k{k}_occ_sha"]).hex(), (use_filter, k)
        assert rg.digest(inf) == bytes(g[f"f{use_filter}_k{k}_inf_sha"]).hex(), (use_filter, k)


@pytest.mark.skipif(not rg.available(), reason="oracle/_ref/libref_gridmap.so not built (needs /root/reference)")
def test_oracle_equals_reference_live_across_the_counter_wrap():
    cam, cloud, cfg = rg.camera(), rg.scene(), rg.config()
    dp = _dp(cam, 1)
    ref = rg.RefGridMap(cfg, dp)
    orc = oracle_lib.OracleGridMap(cfg)
    assert ref.dims == orc.dims
    has_first = False
    rng = np.random.default_rng(3)
    for k in range(8):
        img, R, pos = rg.frame(cam, cloud, k)
        want = ref.project(img, R, pos)
        pts, has_first = oracle_lib.gridmap_project(dp, has_first, img, R, pos, cfg.max_ray_length)
        assert np.array_equal(pts, want), k
        ref.raycast()
        orc.update(pts, pos)
    # 300 more frames of explicit points (cheaper than rendering): far, outside-the-map and duplicate end points.
    # `char raycast_num_` (grid_map.h:140) wraps 127 -> -128 at frame 128; at frame 255 it equals the flags' initial
    # -1, so every never-touched voxel reads as already traversed; from frame 257 on it meets 256-frame-old flags.
    for k in range(300):
        pos = np.array([-3.0 + 0.02 * k, 0.3, 1.0])
        pts = pos + rng.normal(size=(300, 3)) * np.array([3.0, 3.0, 0.6])
        pts[::17] = pos + np.array([9.0, 0.0, 0.0])            # beyond max_ray_length
        pts[::23] = np.array([40.0, 1.0, 1.0])                 # outside the map
        pts[1::2] = pts[::2][: len(pts[1::2])]                 # duplicates: one ray per end voxel while flags match
        ref.raycast(pts, pos)
        orc.update(pts, pos)
        if k % 16 == 0 or 115 < k < 125 or 243 < k < 254 or k > 295:
            (o1, i1, b1), (o2, i2, b2) = ref.buffers(), orc.buffers()
            assert np.array_equal(b1, b2) and np.array_equal(o1, o2) and np.array_equal(i1, i2), k
    (o1, i1, b1), (o2, i2, b2) = ref.buffers(), orc.buffers()
    assert np.array_equal(b1, b2) and np.array_equal(o1, o2) and np.array_equal(i1, i2)
    assert (o1 > np.log(cfg.p_occ / (1 - cfg.p_occ))).sum() > 100
    # the two queries of grid_map.h:333-349 on the final map
    occ_thr = np.log(cfg.p_occ / (1 - cfg.p_occ))
    org = np.array([-cfg.map_size[0] / 2, -cfg.map_size[1] / 2, cfg.ground_height])
    for q in rng.uniform([-9, -9, 0.0], [9, 9, 2.9], (500, 3)):
        idx = tuple(np.floor((q - org) / cfg.resolution).astype(int))
        assert ref.query(q) == int(o2[idx] > occ_thr) and ref.query(q, inflate=True) == int(i2[idx])
