"""libpop (CUDA, through the C ABI) against the COMMITTED golden vectors of BASELINE config 1
(tests/golden/oracle_c1.npz, made by scripts/make_golden.py): occupied (voxel, slice) set and
occupancy queries bit-exact, A* path identical, QP control points and the published trajectory
within 1e-4 relative.  Nothing here needs the oracle at run time."""
import os

import numpy as np
import pytest

import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import POP_MAX_PIECES
from tests import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-4   # north_star tolerance on trajectory coefficients


@pytest.mark.parametrize("cell", ["u8", "f32"])
def test_config1_against_committed_vectors(cell):
    g = np.load(os.path.join(GOLD, "oracle_c1.npz"))
    cfg = cases.cfg_c1(p.POP_CELL_U8 if cell == "u8" else p.POP_CELL_F32)
    cloud, cyl = cases.world_c1()
    start = g["start"].reshape(1, 9)
    goal = g["goal"].reshape(1, 3)
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    eng.set_poses(start[:, :3].astype(np.float32))
    eng.map_update_gt(float(g["t_map"]))
    amap = eng.agent_map(0)
    occ = np.argwhere(amap.risk_maps() > 0).astype(np.int32)
    assert np.array_equal(occ, g["occupied"]), "occupied (voxel, slice) set differs"
    assert np.array_equal(amap.getClearOcccupancy(g["query_pos"], g["query_t"]), g["query_out"])
    ret, path, ln, _ = eng.astar_search_batch(start, goal, [float(g["t_plan"] - g["t_map"])])
    assert ret[0] == int(g["astar_ret"]) and ln[0] == len(g["astar_path"])
    assert np.array_equal(path[0, :ln[0]], g["astar_path"])
    hp, rows, npoly, lgoal = eng.corridor_batch(path, ln, start, [float(g["t_plan"])])
    M = int(g["n_poly"])
    assert npoly[0] == M and np.array_equal(rows[0, :M], g["rows"])
    fin = np.zeros((1, 9))
    fin[0, :6] = lgoal[0]
    dur = np.zeros((1, POP_MAX_PIECES))
    dur[0, :M] = cfg.corridor_tau
    cpts, status, iters = eng.qp_solve_batch(npoly, start, fin, dur, hp, rows, cfg.opt_max_vel, cfg.opt_max_acc)
    assert status[0] == int(g["qp_status"]) and iters[0] == int(g["qp_iters"])
    ref = g["qp_cpts"]
    assert np.abs(cpts[0, :5 * M] - ref).max() <= RTOL * max(1.0, np.abs(ref).max())
    st, tr = eng.replan_cycle(start, goal, start[:, :3].astype(np.float32), float(g["t_map"]), float(g["t_plan"]))
    assert st[0] == int(g["replan_status"]) and tr["n_pieces"][0] == int(g["record_n_pieces"])
    rc = g["record_cpts"]
    assert np.abs(tr["cpts"][0] - rc).max() <= RTOL * max(1.0, np.abs(rc).max())


def test_sensor_and_particle_map_against_committed_vectors():
    """Depth sensor + particle map on the fixed 6-frame flight of scripts/make_golden_dsp.py against
    tests/golden/oracle_dsp.npz: depth images, voxel-filter output, every live particle, drop counters and
    voxel occupancy bit-exact; future risk within 1e-5 relative (atomic fp32 deposits)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), os.pardir, "scripts"))
    from make_golden_dsp import scenario
    from pred_occ_planner_b200 import dsp
    g = np.load(os.path.join(GOLD, "oracle_dsp.npz"))
    cfg, d, cam, cloud, tabs, frames = scenario()
    eng = p.PopEngine(cfg)
    maps = dsp.DspMaps(eng, d)
    maps.set_camera(cam=cam)
    maps.set_random_tables(*tabs)
    eng.set_world_cloud(cloud)
    for k, (pos, quat, st) in enumerate(frames):
        counts, depth = maps.render(pos[None], quat[None], want_depth=True)
        if k in (0, 5):
            assert np.array_equal(depth[0], g[f"depth_{k}"]) and counts[0] == int(g[f"ncloud_{k}"])
            assert np.array_equal(maps.filterPointCloud(0, maps.rendered_cloud(0)), g[f"filtered_{k}"])
            maps.render(pos[None], quat[None])      # the filter call above consumed the rendered input
        assert maps.cloudCallbackRendered(pos[None].astype(np.float32), quat[None].astype(np.float32), st).tolist() == [1]
        assert np.array_equal(maps.stats(0)[:7], g[f"stats_{k}"][:7]), k
        assert len(maps.getOccupancyMap(0)) == int(g[f"nocc_{k}"])
    parts = maps.particles(0)
    live = np.argwhere(parts[:, :, 0] > 0.1).astype(np.int32)
    assert np.array_equal(live, g["live_slots"])
    assert np.array_equal(parts[live[:, 0], live[:, 1]], g["live_state"])
    vox = maps.voxels(0)
    occ = np.nonzero(vox[:, 0] > 0)[0].astype(np.int32)
    assert np.array_equal(occ, g["vox_idx"]) and np.array_equal(vox[occ], g["vox_val"])
    risk = eng.agent_map(0).risk_maps()
    nz = np.argwhere(risk != 0).astype(np.int32)
    assert np.array_equal(nz, g["risk_idx"])
    assert np.allclose(risk[nz[:, 0], nz[:, 1]], g["risk_val"], rtol=1e-5, atol=1e-6)
    eng.close()
