"""Pins of the oracle's particle map against the REFERENCE's own node.

`oracle/_ref/libref_dsp_ref.so` = plan_env/src/risk_mapping_node.cpp with plan_env/include/plan_env/dsp_dynamic.h,
compiled unmodified from /root/reference (oracle/Makefile, oracle/ref_glue_dsp.cpp): the node's filterPointCloud,
odomCallback, cloudCallback and publishMap drive dsp_map::DSPMap exactly as in the reference.  The reference seeds its
noise with time(NULL); here the two Gaussian tables and the rand() stream are the same inputs on both sides (the
library answers the reference's rand() calls itself), which makes every frame deterministic, and EVERYTHING is compared
bit for bit: the observation bins with their C_k, every particle of voxels_with_particle (flag, slot, velocity,
position, weight), pyramids_in_fov, the per-point cluster velocities, voxel occupancy and mean velocity, the three
sequence counters, and the future-risk grid the node publishes.

Third-party arithmetic inside that path is NOT pinned (absent libraries, restated as stand-ins under
oracle/ref_build_map): PCL's VoxelGrid and EuclideanClusterExtraction (FLANN radius search), munkres-cpp.  Eigen's
Quaternionf product is restated as Eigen 3.3's x86-64 SSE kernel evaluates it (oracle/ref_build/Eigen/Geometry)."""
import os

import numpy as np
import pytest

from pred_occ_planner_b200 import worldgen
from tests import cases
from tests import ref_dsp_lib as rdl

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_dsp_pins.npz")
EXACT = ("n_obs", "n_live", "n_pyr", "n_pv", "obs", "flags", "particles", "pyramids", "velocity", "voxels", "occupancy_sum")


def _oracle_map(oracle, seed=5):
    dcfg = cases.dsp_cfg_ref()
    dpar = cases.dsp_params(dcfg)
    od = oracle.dsp(dcfg, dpar)
    od.set_random_tables(*worldgen.dsp_random_tables(dpar.n_randoms, seed))
    return od, dpar


def test_oracle_flight_equals_reference_vectors(oracle):
    gold = np.load(GOLD)
    od, _ = _oracle_map(oracle)
    for k, (cam, pos, quat, st) in enumerate(rdl.flight_frames()):
        assert od.update(cam, pos, quat, st) == gold[f"f{k}.ok"] == 1
        got = rdl.frame_state(od)
        for name in EXACT:
            assert np.array_equal(got[name], gold[f"f{k}.{name}"]), (k, name)
        assert np.array_equal(od.stats()[4:8], gold[f"f{k}.stats"]), k
        risk, _ = od.export()
        if k in rdl.RISK_FRAMES:
            nz, val = rdl.sparse(risk)
            assert np.array_equal(nz, gold[f"f{k}.risk_nz"]) and np.array_equal(val, gold[f"f{k}.risk_val"]), k
    assert gold["f9.n_live"] > 5000 and gold["f9.n_obs"] > 500


def test_golden_file_is_current():
    gold = np.load(GOLD)
    dpar = cases.dsp_params(cases.dsp_cfg_ref())
    rd = rdl.fresh(sigma_observation=dpar.sigma_observation, filter_res=dpar.voxel_filter_resolution,
                   risk_threshold=dpar.occupancy_threshold)
    if rd is None:
        pytest.skip("oracle/_ref/libref_dsp_ref.so not present (no /root/reference here)")
    rd.set_random_tables(*worldgen.dsp_random_tables(dpar.n_randoms, 5))
    for k, (cam, pos, quat, st) in enumerate(rdl.flight_frames(4)):
        assert rd.update(cam, pos, quat, st) == 1
        got = rdl.frame_state(rd)
        for name in EXACT:
            assert np.array_equal(got[name], gold[f"f{k}.{name}"]), (k, name)
        rd.export()


@pytest.mark.parametrize("seed", [11, 12])
def test_other_flights_live(oracle, seed):
    """Other noise tables, a faster and turning flight, a frame with rejected odometry and an empty cloud: the oracle
    against the live node, array by array."""
    dpar = cases.dsp_params(cases.dsp_cfg_ref())
    rd = rdl.fresh(sigma_observation=dpar.sigma_observation, filter_res=dpar.voxel_filter_resolution,
                   risk_threshold=dpar.occupancy_threshold)
    if rd is None:
        pytest.skip("oracle/_ref/libref_dsp_ref.so not present (no /root/reference here)")
    od, _ = _oracle_map(oracle, seed)
    rd.set_random_tables(*worldgen.dsp_random_tables(dpar.n_randoms, seed))
    scene = cases.dsp_scene(seed=seed, n_static=4, n_moving=6)
    rng = np.random.default_rng(seed)
    for k in range(8):
        pos, quat, st = cases.dsp_flight(k, speed=1.8)
        quat = worldgen.quat_yaw(0.15 * k - 0.4)
        cam = worldgen.render_depth_cloud(scene, pos, quat, t=0.1 * k)
        if k == 3:
            quat = quat.copy()
            quat[0] = 1.5                       # invalid quaternion: both reject the frame (dsp_dynamic.h:183-188)
        if k == 5:
            cam = cam[:0]                       # nothing in view
        a, b = rd.update(cam, pos, quat, st), od.update(cam, pos, quat, st)
        assert a == b == (0 if k == 3 else 1), k
        if not a:
            continue
        pc_r, cnt_r, ml_r = rd.observations()
        pc_o, cnt_o, ml_o = od.observations()
        assert np.array_equal(cnt_r, cnt_o) and np.array_equal(ml_r, ml_o)
        m = np.arange(pc_o.shape[1])[None, :] < cnt_o[:, None]
        assert np.array_equal(pc_r[m][:, :4], pc_o[m][:, :4])
        assert np.array_equal(rd.points_with_velocity(), od.points_with_velocity())
        pa_r, pa_o = rd.particles(), od.particles()
        assert np.array_equal(pa_r[:, :, 0], pa_o[:, :, 0])
        live = pa_o[:, :, 0] > 0.1
        assert np.array_equal(pa_r[live], pa_o[live])
        py_r, py_o = rd.pyramids(), od.pyramids()
        assert np.array_equal(py_r[:, :, 0] != 0, py_o[:, :, 0] != 0)
        assert np.array_equal(py_r[py_r[:, :, 0] != 0], py_o[py_o[:, :, 0] != 0])
        assert np.array_equal(rd.voxels(), od.voxels())
        assert np.array_equal(rd.stats(), od.stats()[4:8])
        assert np.array_equal(rd.export()[0], od.export()[0])
    assert live.sum() > 3000


def test_filter_point_cloud_live(oracle):
    """filterPointCloud of the node (risk_mapping_node.cpp:74-99) against the oracle's: axis swap, range test against
    the double LOCAL_RANGE constants, the 5 000-point cap (the voxel-grid centroids come from the stand-in)."""
    dpar = cases.dsp_params(cases.dsp_cfg_ref())
    rd = rdl.fresh(sigma_observation=dpar.sigma_observation)
    if rd is None:
        pytest.skip("oracle/_ref/libref_dsp_ref.so not present (no /root/reference here)")
    od, _ = _oracle_map(oracle)
    rng = np.random.default_rng(3)
    for n, scale in ((3000, 3.0), (60000, 6.5), (0, 1.0)):
        cam = (rng.uniform(-1, 1, (n, 3)) * [scale, 1.6, scale]).astype(np.float32)
        a, b = rd.filter(cam), od.filter(cam)
        assert a.shape == b.shape and np.array_equal(a, b), n
    assert len(b) == 0 and len(rd.filter((rng.uniform(-1, 1, (60000, 3)) * [6.5, 1.4, 6.5]).astype(np.float32))) == 5000
