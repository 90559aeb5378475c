"""The simulator's depth sensor against the reference's OWN kernel: uav_simulator/local_sensing/src/depth_render.cu,
compiled unmodified by nvcc for sm_100a into oracle/_ref/libref_depth.so (oracle/ref_glue_depth.cu), run on the B200
beside the product's kernels.  For identical inputs (cloud, intrinsics, camera pose handed over as the 4x4 the node
gives `render_pose`) the depth images are integer millimetres produced by atomicMin, so the comparison is exact:
reference kernel == oracle restatement (o_render.cpp: render_depth) == product (k_depth_splat through pop_render_depth)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import cases, oracle_lib
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_depth.so")


def _quat_to_rot(q):
    w, x, y, z = (float(v) for v in q)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return [1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)]


def _cam_pose(cam, pos, quat):
    """cam2world = body pose * cam02body, then its rigid inverse, with the oracle's (= the node's) order of double
    operations (o_render.cpp: render_cam2world, render_depth) — plain Python floats, no BLAS."""
    R = _quat_to_rot(quat)
    B = [R[0], R[1], R[2], float(pos[0]), R[3], R[4], R[5], float(pos[1]), R[6], R[7], R[8], float(pos[2]), 0.0, 0.0, 0.0, 1.0]
    c2b = [float(v) for v in cam.cam2body]
    c2w = [0.0] * 16
    for i in range(4):
        for j in range(4):
            s = 0.0
            for k in range(4):
                s += B[4 * i + k] * c2b[4 * k + j]
            c2w[4 * i + j] = s
    T = [0.0] * 16
    for i in range(3):
        s = 0.0
        for j in range(3):
            T[4 * i + j] = c2w[4 * j + i]
            s += c2w[4 * j + i] * c2w[4 * j + 3]
        T[4 * i + 3] = -s
    T[15] = 1.0
    return np.array(T, np.float64)


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_depth.so not built (needs /root/reference and nvcc)")
def test_depth_images_equal_the_reference_kernel(pop):
    from pred_occ_planner_b200 import dsp, worldgen
    ref = C.CDLL(LIB)
    n = 3
    cfg = cases.dsp_cfg_ref(n_local=n, n_total=n)
    eng = pop.PopEngine(cfg, device=0)
    maps = dsp.DspMaps(eng, cases.dsp_params(cfg))
    cam = maps.set_camera(160, 120)
    cloud, _ = worldgen.make_world(11, arena=(10.0, 10.0, 3.0), n_static=8, n_moving=0, height=3.0)
    gx, gy = np.meshgrid(np.arange(-5, 5, 0.1), np.arange(-5, 5, 0.1), indexing="ij")
    cloud = np.concatenate([cloud, np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1)]).astype(np.float32)
    eng.set_world_cloud(cloud)
    seen = 0
    for k in range(4):
        pos = np.array([[-4.0 + 0.4 * k, -1.0 + 1.0 * a, 1.0 + 0.1 * a] for a in range(n)])
        quat = np.array([[np.cos(0.1 * a + 0.15 * k), 0, 0, np.sin(0.1 * a + 0.15 * k)] for a in range(n)])
        _, depth = maps.render(pos, quat, want_depth=True)
        for a in range(n):
            T = _cam_pose(cam, pos[a], quat[a])
            want = np.zeros((cam.height, cam.width), np.int32)
            rc = ref.refdepth_render(cloud.ctypes.data_as(C.POINTER(C.c_float)), len(cloud), C.c_float(cam.fx), C.c_float(cam.fy),
                                     C.c_float(cam.cx), C.c_float(cam.cy), cam.width, cam.height,
                                     T.ctypes.data_as(C.POINTER(C.c_double)), want.ctypes.data_as(C.POINTER(C.c_int)))
            assert rc == 0
            orc_depth, _, _ = oracle_lib.render(cam, pos[a], quat[a], cloud)
            assert np.array_equal(orc_depth, want), (k, a, int((orc_depth != want).sum()))     # oracle == reference kernel
            assert np.array_equal(depth[a], want), (k, a, int((depth[a] != want).sum()))       # product == reference kernel
            seen += int((want < 999999).sum())
    assert seen > 12 * 5000
    eng.close()
