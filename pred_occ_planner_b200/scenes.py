"""The reference's particle-map configuration and a small depth-camera scene (what `bench.py`'s particle-map leg,
`scripts/dsp_bench.py` and the tests fly): grid and filter parameters of plan_env/include/plan_env/map_parameters.h and
risk_mapping_node.cpp, a few static and moving cylinders, a camera flying a straight line."""
import numpy as np

import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import worldgen


def dsp_cfg_ref(n_local=1, n_total=1):
    """The reference's particle-map configuration: 66x66x20 @0.15 m, 6 future slices
    (plan_env/include/plan_env/map_parameters.h:5-18), RiskBase queries (plan_manager/config/real_cfg.yaml)."""
    c = p.default_config()
    c.nx, c.ny, c.nz, c.nt = 66, 66, 20, 6
    c.resolution = 0.15
    c.time_resolution = 0.3
    c.risk_threshold = 0.2
    c.risk_threshold_region = 0.5
    c.risk_region_decay = 0.01
    c.risk_voxel_decay = 0.02
    c.clearance = 0.15
    c.kernel_shape = p.POP_KERNEL_CUBE
    c.query_variant = p.POP_QUERY_RISKBASE
    c.cell_dtype = p.POP_CELL_F32
    c.n_agents_local, c.n_agents_total = n_local, n_total
    return c


def dsp_params(cfg, n_randoms=1 << 20, sigma_observation=0.1):
    """pop_dsp_config with the reference's values (map_parameters.h; risk_mapping_node.cpp:232-241),
    written out here so that the CPU-only tests do not need libpop's default filler."""
    d = p.PopDspConfig()
    d.half_fov_h, d.half_fov_v, d.angle_resolution = 43, 29, 1
    d.max_particle_num_voxel = 7
    V = cfg.nx * cfg.ny * cfg.nz
    d.pyramid_capacity = int(V * 7 + 1e5) // (360 * 180) * 2
    d.max_points_per_pyramid = 100
    d.max_points = 5000
    d.max_input_points = 32768
    d.new_born_particle_number_each_point = 20
    d.dynamic_cluster_max_point_num = 200
    d.dynamic_cluster_max_center_height = 1.5
    d.new_born_particle_weight = 0.0001
    d.sigma_observation = sigma_observation
    d.kappa = 0.01
    d.p_detection = 0.95
    d.voxel_filter_resolution = 0.15
    d.obstacle_thickness_for_occlusion = 0.3
    d.occupancy_threshold = 0.2
    d.position_prediction_stddev = 0.01
    d.velocity_prediction_stddev = 0.01
    for k in range(cfg.nt):
        d.prediction_future_time[k] = 0.3 * (k + 1)
    d.n_randoms = n_randoms
    d.seed = 20221023
    return d


def dsp_scene(seed=7, n_static=6, n_moving=4, arena=(10.0, 10.0, 3.0)):
    """Cylinders around a forward-flying sensor (the reference's simulator world at config-1 scale)."""
    _, cyl = worldgen.make_world(seed, arena=arena, n_static=n_static, n_moving=n_moving, height=3.0)
    return cyl


def dsp_flight(k, speed=1.0, dt=0.1):
    """Sensor pose of frame k: straight flight along +x from (-4, 0, 1) with a slow yaw sweep."""
    import math
    pos = np.array([-4.0 + speed * dt * k, 0.2 * math.sin(0.3 * k), 1.0], np.float32)
    quat = worldgen.quat_yaw(0.25 * math.sin(0.2 * k))
    return pos, quat, 100.0 + dt * k
