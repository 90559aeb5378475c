"""Shared seeded test cases: BASELINE configs restated as pop_config + synthetic inputs."""
import numpy as np

import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import worldgen


def cfg_c1(cell=p.POP_CELL_F32, n_local=1, n_total=1):
    """BASELINE config 1: 10x10x3 m @0.15 m (66x66x20), reference simulator parameters
    (plan_manager/config/sim_fake.yaml), T = 9 slices of 0.2 s."""
    c = p.default_config()
    c.cell_dtype = cell
    c.n_agents_local, c.n_agents_total = n_local, n_total
    return c


def cfg_small_c2(cell=p.POP_CELL_U8, n_local=4, n_total=4, nx=120, ny=120, nz=30, nt=10):
    """BASELINE config 2 geometry (0.1 m voxels, 0.1 s slices, moving cylinders) on a smaller grid."""
    c = p.default_config()
    c.nx, c.ny, c.nz, c.nt = nx, ny, nz, nt
    c.resolution = 0.1
    c.time_resolution = 0.1
    c.resolution_astar = 0.1
    c.time_resolution_astar = 0.2
    c.corridor_tau = 0.2
    c.cell_dtype = cell
    c.n_agents_local, c.n_agents_total = n_local, n_total
    return c


def world_c1(seed=1235):
    return worldgen.make_world(seed, arena=(10.0, 10.0, 3.0), n_static=20, n_moving=0)


def world_moving(seed, arena, n_static, n_moving):
    return worldgen.make_world(seed, arena=arena, n_static=n_static, n_moving=n_moving)


def occupied_set(risk_maps: np.ndarray):
    """(voxel index, slice) pairs with non-zero risk — the bit-exact parity object."""
    v, t = np.nonzero(risk_maps)
    return set(zip(v.tolist(), t.tolist()))


# the particle-map configuration and scene live in the package (bench.py and scripts use them too)
from pred_occ_planner_b200.scenes import dsp_cfg_ref, dsp_params, dsp_scene, dsp_flight   # noqa: E402,F401
