"""CPU checks of the depth-sensor oracle (oracle/o_render.cpp) on scenes with known geometry."""
import numpy as np

import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import worldgen
from tests import oracle_lib


def _camera(w=160, h=120):
    cam = p._lib.PopCamera()
    p.load().pop_camera_default(p._lib.C.byref(cam), w, h)
    return cam


def test_wall_depth_and_frames():
    cam = _camera()
    # a dense wall 3 m ahead of a sensor at the origin looking along +x
    ys, zs = np.meshgrid(np.arange(-3, 3, 0.05), np.arange(-2, 2, 0.05), indexing="ij")
    wall = np.stack([np.full(ys.size, 3.0), ys.ravel(), zs.ravel()], axis=1).astype(np.float32)
    depth, pc, pw = oracle_lib.render(cam, [0, 0, 0], [1, 0, 0, 0], wall)
    assert depth.shape == (120, 160)
    assert (depth == 3000).mean() > 0.99          # millimetres along the optical axis, everywhere
    assert len(pc) == depth.size and np.allclose(pc[:, 2], 3.0)
    # camera frame: x right, y down, z forward; world frame: x forward, y left, z up (cam02body)
    assert np.allclose(pw[:, 0], 3.0, atol=1e-6)
    assert np.allclose(pw[:, 1], -pc[:, 0], atol=1e-6) and np.allclose(pw[:, 2], -pc[:, 1], atol=1e-6)
    # pixel order of render_pcl_world: u outer, v inner -> camera x non-decreasing
    assert (np.diff(pc[:, 0]) >= -1e-6).sum() >= len(pc) - 160


def test_occlusion_horizon_and_pose():
    cam = _camera()
    near = np.array([[2.0, 0.0, 0.0]], np.float32)
    far = np.array([[5.0, 0.0, 0.0]], np.float32)
    beyond = np.array([[0.0, 7.5, 0.0]], np.float32)
    depth, pc, pw = oracle_lib.render(cam, [0, 0, 0], [1, 0, 0, 0], np.concatenate([far, near]))
    u, v = int(cam.cx + 0.5), int(cam.cy + 0.5)
    assert depth[v, u] == 2000                                     # the nearer splat wins the z-buffer
    r_near, r_far = int(0.0573 * cam.fx / 2.0 + 0.5), int(0.0573 * cam.fx / 5.0 + 0.5)
    assert (depth == 2000).sum() == (2 * r_near + 1) ** 2 and (depth == 5000).sum() == 0 or r_far > r_near
    # yawed 90 deg to the left the sensor looks along +y: the 7.5 m point is seen but outside the 6 m horizon
    q = worldgen.quat_yaw(np.pi / 2).astype(np.float64)
    depth, pc, pw = oracle_lib.render(cam, [0, 0, 0], q, beyond)
    assert (depth == 7500).sum() > 0 and len(pc) == 0
    # translated sensor: same scene seen from 2 m closer
    depth, pc, pw = oracle_lib.render(cam, [0, 2.0, 0], q, beyond)
    assert (depth == 5500).sum() > 0 and len(pc) == (depth == 5500).sum() and np.allclose(pw[:, 1], 7.5, atol=1e-5)
    # nothing behind the camera, nothing from an empty cloud
    depth, pc, _ = oracle_lib.render(cam, [0, 0, 0], [1, 0, 0, 0], np.array([[-2.0, 0, 0]], np.float32))
    assert (depth == 999999).all() and len(pc) == 0
    depth, pc, _ = oracle_lib.render(cam, [0, 0, 0], [1, 0, 0, 0], np.zeros((0, 3), np.float32))
    assert (depth == 999999).all() and len(pc) == 0


def test_point_on_the_lens_fills_the_image_and_terminates():
    """A point millimetres in front of the camera has a footprint radius of ~1e4 pixels: it covers the whole
    image (the reference would loop over 1e8 pixel offsets and overflow its int conversion closer still)."""
    cam = _camera()
    depth, pc, _ = oracle_lib.render(cam, [0, 0, 0], [1, 0, 0, 0], np.array([[0.002, 0.0, 0.0], [1e-7, 0.0, 0.0]], np.float32))
    assert (depth == 0).all() or (depth == 2).all()       # the nearer one rounds to 0 mm and wins everywhere
    assert len(pc) == 0 or np.allclose(pc[:, 2], 0.002)
