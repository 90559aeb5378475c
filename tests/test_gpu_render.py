"""GPU parity of the depth sensor (K0) and of the device-resident perception loop it feeds:
world cloud -> depth images -> camera-frame clouds -> voxel filter -> particle map, nothing crossing PCIe.
Depth images and clouds are bit-exact against the oracle (integer z-buffer, fixed IEEE sequences)."""
import numpy as np
import pytest

from tests import cases, dsp_compare, oracle_lib

pytestmark = pytest.mark.gpu


def _setup(pop, n, w=160, h=120):
    from pred_occ_planner_b200 import dsp, worldgen
    cfg = cases.dsp_cfg_ref(n_local=n, n_total=n)
    d = cases.dsp_params(cfg)
    eng = pop.PopEngine(cfg, device=0)
    maps = dsp.DspMaps(eng, d)
    cam = maps.set_camera(w, h)
    cloud, cyl = worldgen.make_world(11, arena=(10.0, 10.0, 3.0), n_static=8, n_moving=0, height=3.0)
    # a ground sheet so that the lower half of the image is filled too
    gx, gy = np.meshgrid(np.arange(-5, 5, 0.1), np.arange(-5, 5, 0.1), indexing="ij")
    ground = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1).astype(np.float32)
    cloud = np.concatenate([cloud, ground]).astype(np.float32)
    eng.set_world_cloud(cloud)
    return cfg, d, eng, maps, cam, cloud


def _poses(n, k):
    pos = np.array([[-4.0 + 0.1 * k, -1.0 + 1.0 * a, 1.0 + 0.1 * a] for a in range(n)])
    quat = np.array([[np.cos(0.1 * a + 0.02 * k), 0, 0, np.sin(0.1 * a + 0.02 * k)] for a in range(n)])
    return pos, quat


def test_depth_images_and_clouds_bit_exact(pop):
    n = 3
    cfg, d, eng, maps, cam, cloud = _setup(pop, n)
    pos, quat = _poses(n, 0)
    counts, depth = maps.render(pos, quat, want_depth=True)
    for a in range(n):
        ref_depth, ref_cam, _ = oracle_lib.render(cam, pos[a], quat[a], cloud)
        assert np.array_equal(depth[a], ref_depth)
        got = maps.rendered_cloud(a)
        assert counts[a] == len(ref_cam) > 1000
        assert np.array_equal(got, ref_cam)
    # an empty world renders empty images
    eng.set_world_cloud(np.zeros((0, 3), np.float32))
    counts, depth = maps.render(pos, quat, want_depth=True)
    assert (depth == 999999).all() and (counts == 0).all()
    eng.close()


def test_device_resident_perception_loop(pop, oracle):
    """render -> cloudCallbackRendered equals the oracle fed with the oracle-rendered clouds, frame by frame."""
    from pred_occ_planner_b200 import worldgen
    n = 2
    cfg, d, eng, maps, cam, cloud = _setup(pop, n)
    tabs = worldgen.dsp_random_tables(d.n_randoms, 5)
    maps.set_random_tables(*tabs)
    orcs = [oracle.dsp(cfg, d) for _ in range(n)]
    for o in orcs:
        o.set_random_tables(*tabs)
    for k in range(5):
        pos, quat = _poses(n, k)
        st = 30.0 + 0.1 * k
        counts, _ = maps.render(pos, quat)
        ok = maps.cloudCallbackRendered(pos.astype(np.float32), quat.astype(np.float32), st)
        assert ok.tolist() == [1] * n and (counts > 1000).all()
        for a in range(n):
            _, ref_cam, _ = oracle_lib.render(cam, pos[a], quat[a], cloud)
            assert orcs[a].update(ref_cam, pos[a].astype(np.float32), quat[a].astype(np.float32), st) == 1
            bad = dsp_compare.compare_frame(maps, a, orcs[a], eng.agent_map(a).risk_maps(), d.occupancy_threshold)
            assert not bad, f"frame {k} agent {a}: " + "; ".join(bad)
    # a host-cloud update in between invalidates the rendered input
    with pytest.raises(pop.PopError):
        maps.cloudCallback([np.zeros((0, 3), np.float32)] * n, pos.astype(np.float32), quat.astype(np.float32), st + 0.1)
        maps.cloudCallbackRendered(pos.astype(np.float32), quat.astype(np.float32), st + 0.2)
    eng.close()


def test_point_on_the_lens(pop):
    """Footprints far larger than the image are clipped to it (and capped before the int conversion)."""
    cfg, d, eng, maps, cam, cloud = _setup(pop, 1)
    near = np.array([[-3.998, -1.0, 1.0], [-4.0 + 1e-7, -1.0, 1.0], [-1.0, -1.0, 1.0]], np.float32)
    eng.set_world_cloud(near)
    pos, quat = _poses(1, 0)
    counts, depth = maps.render(pos, quat, want_depth=True)
    ref_depth, ref_cam, _ = oracle_lib.render(cam, pos[0], quat[0], near)
    assert np.array_equal(depth[0], ref_depth) and counts[0] == len(ref_cam)
    eng.close()
