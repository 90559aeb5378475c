"""ctypes binding of libpop.so (include/pop.h).

The library is the product; this module only declares its C ABI.  There is no CPU
fallback: if the shared object is missing, or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

POP_ABI_VERSION = 1
POP_MAX_PIECES = 12
POP_MAX_BODY = 64
POP_MAX_PATH = 16
POP_MAX_FACES = 48
POP_STAGE_COUNT = 8

POP_CELL_F32, POP_CELL_U8 = 0, 1
POP_KERNEL_DISC_Z0, POP_KERNEL_CUBE = 0, 1
POP_QUERY_FAKE, POP_QUERY_RISKBASE = 0, 1

STAGE_NAMES = ["voxel_update", "peers", "astar", "corridor", "qp", "separation", "dsp", "r7"]
ERRORS = {0: "OK", -1: "INVALID", -2: "NO_DEVICE", -3: "CUDA", -4: "CAPACITY", -5: "STATE",
          -6: "UNSUPPORTED"}


class PopCylinder(C.Structure):
    _fields_ = [("type", C.c_int32), ("_pad", C.c_int32),
                ("x", C.c_double), ("y", C.c_double), ("z", C.c_double),
                ("w", C.c_double), ("h", C.c_double),
                ("vx", C.c_double), ("vy", C.c_double),
                ("qw", C.c_double), ("qx", C.c_double), ("qy", C.c_double), ("qz", C.c_double)]


class PopTrajRecord(C.Structure):
    _fields_ = [("drone_id", C.c_int32), ("traj_id", C.c_int32), ("n_pieces", C.c_int32),
                ("valid", C.c_int32), ("start_time", C.c_double),
                ("duration", C.c_float * POP_MAX_PIECES),
                ("cpts", C.c_double * (POP_MAX_PIECES * 15)),
                ("_pad", C.c_double * 3)]


assert C.sizeof(PopTrajRecord) == 1536

TRAJ_DTYPE = np.dtype([("drone_id", "<i4"), ("traj_id", "<i4"), ("n_pieces", "<i4"), ("valid", "<i4"),
                       ("start_time", "<f8"), ("duration", "<f4", (POP_MAX_PIECES,)),
                       ("cpts", "<f8", (POP_MAX_PIECES * 5, 3)), ("_pad", "<f8", (3,))])
assert TRAJ_DTYPE.itemsize == 1536

CYL_DTYPE = np.dtype([("type", "<i4"), ("_pad", "<i4"), ("x", "<f8"), ("y", "<f8"), ("z", "<f8"),
                      ("w", "<f8"), ("h", "<f8"), ("vx", "<f8"), ("vy", "<f8"),
                      ("qw", "<f8"), ("qx", "<f8"), ("qy", "<f8"), ("qz", "<f8")])
assert CYL_DTYPE.itemsize == C.sizeof(PopCylinder)


# the planning stages are pipelined over up to 32 side streams: ask for that many hardware work
# queues before any CUDA context exists (a value chosen by the application wins)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


class PopFsmParams(C.Structure):
    _fields_ = [("replan_max_failures", C.c_int32), ("_pad", C.c_int32), ("goal_tolerance", C.c_double),
                ("hover_duration", C.c_double)]


class PopConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("nt", C.c_int32),
        ("resolution", C.c_float), ("time_resolution", C.c_float), ("risk_threshold", C.c_float),
        ("clearance", C.c_float), ("ground_height", C.c_float), ("ceiling_height", C.c_float),
        ("risk_threshold_region", C.c_float), ("risk_region_decay", C.c_float),
        ("risk_voxel_decay", C.c_float),
        ("kernel_shape", C.c_int32), ("query_variant", C.c_int32), ("cell_dtype", C.c_int32),
        ("n_agents_local", C.c_int32), ("n_agents_total", C.c_int32), ("agent_id_base", C.c_int32),
        ("drone_size", C.c_double * 3),
        ("replan_risk_rate", C.c_float), ("num_resample", C.c_int32),
        ("max_tau", C.c_double), ("max_vel", C.c_double), ("max_acc", C.c_double),
        ("w_time", C.c_double), ("horizon", C.c_double), ("lambda_heu", C.c_double),
        ("resolution_astar", C.c_double), ("time_resolution_astar", C.c_double),
        ("allocate_num", C.c_int32), ("check_num", C.c_int32), ("tolerance", C.c_int32),
        ("_pad0", C.c_int32),
        ("corridor_tau", C.c_double), ("init_range", C.c_double), ("shrink_size", C.c_double),
        ("opt_max_vel", C.c_double), ("opt_max_acc", C.c_double),
        ("max_cloud_points", C.c_int32), ("max_cylinders", C.c_int32),
        ("max_obstacle_points", C.c_int32),
        ("osqp_max_iter", C.c_int32), ("osqp_adaptive_rho_interval", C.c_int32),
        ("osqp_check_termination", C.c_int32),
        ("osqp_eps_abs", C.c_double), ("osqp_eps_rel", C.c_double),
    ]

    def copy(self) -> "PopConfig":
        c = PopConfig()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(PopConfig))
        return c


POP_DSP_MAX_T = 40


class PopDspConfig(C.Structure):
    """pop_dsp_config (include/pop.h): parameters of one dsp_map::DSPMap."""
    _fields_ = [
        ("half_fov_h", C.c_int32), ("half_fov_v", C.c_int32), ("angle_resolution", C.c_int32),
        ("max_particle_num_voxel", C.c_int32), ("pyramid_capacity", C.c_int32),
        ("max_points_per_pyramid", C.c_int32), ("max_points", C.c_int32), ("max_input_points", C.c_int32),
        ("new_born_particle_number_each_point", C.c_int32), ("dynamic_cluster_max_point_num", C.c_int32),
        ("dynamic_cluster_max_center_height", C.c_float), ("new_born_particle_weight", C.c_float),
        ("sigma_observation", C.c_float), ("kappa", C.c_float), ("p_detection", C.c_float),
        ("voxel_filter_resolution", C.c_float), ("obstacle_thickness_for_occlusion", C.c_float),
        ("occupancy_threshold", C.c_float), ("position_prediction_stddev", C.c_float),
        ("velocity_prediction_stddev", C.c_float),
        ("prediction_future_time", C.c_float * POP_DSP_MAX_T),
        ("n_randoms", C.c_int32), ("_pad", C.c_int32), ("seed", C.c_uint64),
    ]

    def copy(self) -> "PopDspConfig":
        c = PopDspConfig()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(PopDspConfig))
        return c


class PopCamera(C.Structure):
    """pop_camera (include/pop.h): pinhole depth camera of uav_simulator/local_sensing."""
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32), ("sensing_horizon", C.c_float), ("_pad", C.c_int32),
                ("cam2body", C.c_double * 16)]


class PopGridmapConfig(C.Structure):
    """pop_gridmap_config (include/pop.h): parameters of GridMap (plan_env/src/grid_map.cpp:15-131)."""
    _fields_ = [("resolution", C.c_double), ("map_size", C.c_double * 3), ("local_update_range", C.c_double * 3),
                ("obstacles_inflation", C.c_double),
                ("p_hit", C.c_double), ("p_miss", C.c_double), ("p_min", C.c_double), ("p_max", C.c_double),
                ("p_occ", C.c_double), ("max_ray_length", C.c_double), ("virtual_ceil_height", C.c_double),
                ("ground_height", C.c_double), ("local_map_margin", C.c_int32), ("max_points", C.c_int32)]


class PopDepthParams(C.Structure):
    """pop_depth_params (include/pop.h): GridMap::projectDepthImage parameters."""
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("k_depth_scaling_factor", C.c_double), ("depth_filter_mindist", C.c_double),
                ("depth_filter_maxdist", C.c_double), ("use_depth_filter", C.c_int32),
                ("depth_filter_margin", C.c_int32), ("skip_pixel", C.c_int32), ("rows", C.c_int32),
                ("cols", C.c_int32), ("_pad", C.c_int32)]


class PopCycleIO(C.Structure):
    _fields_ = [("start", C.POINTER(C.c_double)), ("goal", C.POINTER(C.c_double)),
                ("poses", C.POINTER(C.c_float)), ("t_map", C.c_double), ("t_plan", C.c_double),
                ("status", C.POINTER(C.c_int32)), ("traj", C.POINTER(PopTrajRecord))]


class PopError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpop error {code} ({ERRORS.get(code, '?')}): {msg}")
        self.code = code


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpop.so")

# every symbol include/pop.h declares: (name, restype, argtypes)
_VP, _I, _D = C.c_void_p, C.c_int, C.c_double
_PD, _PF, _PI32, _PI8 = (C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32),
                         C.POINTER(C.c_int8))
SYMBOLS = [
    ("pop_create", _I, [C.POINTER(PopConfig), _I, _VP, C.POINTER(_VP)]),
    ("pop_destroy", _I, [_VP]),
    ("pop_last_error", C.c_char_p, [_VP]),
    ("pop_abi_version", _I, []),
    ("pop_sync", _I, [_VP]),
    ("pop_default_config", None, [C.POINTER(PopConfig)]),
    ("pop_set_world_cloud", _I, [_VP, _PF, _I]),
    ("pop_set_cylinders", _I, [_VP, C.POINTER(PopCylinder), _I]),
    ("pop_set_world_cloud_dev", _I, [_VP, _VP, _I]),
    ("pop_set_cylinders_dev", _I, [_VP, _VP, _I, _D]),
    ("pop_get_trajectories", _I, [_VP, C.POINTER(PopTrajRecord)]),
    ("pop_fsm_update", _I, [_VP, _D, C.POINTER(PopFsmParams), _PI32, _PI32]),
    ("pop_set_resample_noise", _I, [_VP, _PF, _I]),
    ("pop_comm_unique_id", _I, [_VP]),
    ("pop_comm_init", _I, [_VP, _VP, _I, _I]),
    ("pop_comm_info", _I, [_VP, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    ("pop_exchange_trajs", _I, [_VP]),
    ("pop_exchange_trajs_pipelined", _I, [_VP]),
    ("pop_set_poses", _I, [_VP, _PF]),
    ("pop_set_body_particles", _I, [_VP, _PD, _PI32]),
    ("pop_set_trajectories", _I, [_VP, C.POINTER(PopTrajRecord)]),
    ("pop_map_update_gt", _I, [_VP, _D]),
    ("pop_traj_safe_batch", _I, [_VP, _D, _D, _PI8, _PD]),
    ("pop_map_future_risk", _I, [_VP, _PF, _PF, _D]),
    ("pop_map_query", _I, [_VP, _I, _PD, _PI32, _I, _PI8]),
    ("pop_map_query_dt", _I, [_VP, _I, _PD, _PD, _I, _PI8]),
    ("pop_map_query_inflate", _I, [_VP, _I, _PD, _PD, _I, _I, _PI8]),
    ("pop_map_obstacle_points", _I, [_VP, _I, _D, _D, _PD, _PD, _PD, _I, C.POINTER(_I)]),
    ("pop_map_add_particles", _I, [_VP, _I, _PD, _PF, _I, _I]),
    ("pop_map_download", _I, [_VP, _I, _PF]),
    ("pop_astar_search_batch", _I, [_VP, _PD, _PD, _PD, _PI32, _PD, _PI32, _PI32]),
    ("pop_corridor_batch", _I, [_VP, _PD, _PI32, _PD, _PD, _PD, _PI32, _PI32, _PD]),
    ("pop_qp_solve_batch", _I, [_VP, _I, _PI32, _PD, _PD, _PD, _PD, _PI32, _D, _D, _PD, _PI32, _PI32]),
    ("pop_separation_check_batch", _I, [_VP, _PI32, _PD, _D, _PI8]),
    ("pop_replan_cycle", _I, [_VP, C.POINTER(PopCycleIO)]),
    ("pop_plan_cycle", _I, [_VP, C.POINTER(PopCycleIO)]),
    ("pop_replan_cycle_resident", _I, [_VP, _D, _D]),
    ("pop_cycle_results", _I, [_VP, _PI32, C.POINTER(PopTrajRecord)]),
    ("pop_traj_publish_local", _I, [_VP]),
    ("pop_traj_buffers", _I, [_VP, C.POINTER(_VP), C.POINTER(_VP)]),
    ("pop_profile_enable", _I, [_VP, _I]),
    ("pop_profile_read", _I, [_VP, _PF, C.POINTER(C.c_int64), _I]),
    ("pop_qp_profile", _I, [_VP, _I, C.POINTER(C.c_int64)]),
    ("pop_lp_solve", _I, [_VP, _I, _I, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint64, _I,
                      C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("pop_kernel_launches", C.c_int64, [_VP]),
    ("pop_dsp_default_config", None, [C.POINTER(PopConfig), C.POINTER(PopDspConfig)]),
    ("pop_dsp_attach", _I, [_VP, C.POINTER(PopDspConfig)]),
    ("pop_dsp_set_random_tables", _I, [_VP, _PF, _PF, _PI32, _I]),
    ("pop_dsp_update", _I, [_VP, _PF, _PI32, _PF, _PF, _D, _PI32]),
    ("pop_dsp_update_filtered", _I, [_VP, _PF, _PI32, _PF, _PF, _D, _PI32]),
    ("pop_dsp_filter", _I, [_VP, _I, _PF, _I, _PF, _PI32]),
    ("pop_dsp_download_particles", _I, [_VP, _I, _PF]),
    ("pop_dsp_download_voxels", _I, [_VP, _I, _PF]),
    ("pop_dsp_download_observations", _I, [_VP, _I, _PF, _PI32, _PF]),
    ("pop_dsp_download_pyramids", _I, [_VP, _I, _PI32]),
    ("pop_dsp_download_points_with_velocity", _I, [_VP, _I, _PF, _I, _PI32]),
    ("pop_dsp_stats", _I, [_VP, _I, _PI32]),
    ("pop_dsp_occupied", _I, [_VP, _I, _PF, _I, _PI32]),
    ("pop_camera_default", None, [C.POINTER(PopCamera), _I, _I]),
    ("pop_render_set_camera", _I, [_VP, C.POINTER(PopCamera)]),
    ("pop_render_depth", _I, [_VP, _PD, _PD, _PI32, _PI32]),
    ("pop_render_download_cloud", _I, [_VP, _I, _PF, _I, _PI32]),
    ("pop_dsp_update_rendered", _I, [_VP, _PF, _PF, _D, _PI32]),
    ("pop_traj_sample_batch", _I, [_VP, _I, _PD, _I, _PD, _PI8]),
    ("pop_swarm_advance", _I, [_VP, _D]),
    ("pop_swarm_states", _I, [_VP, _PD]),
    ("pop_eval_metrics", _I, [_VP, _D, _PD, _PD]),
    ("pop_traj_to_wire", _I, [C.POINTER(PopTrajRecord), _D, C.POINTER(C.c_uint8), _I]),
    ("pop_traj_from_wire", _I, [C.POINTER(C.c_uint8), _I, C.POINTER(PopTrajRecord), _PD]),
    ("pop_gridmap_default_config", None, [C.POINTER(PopGridmapConfig)]),
    ("pop_gridmap_create", _I, [C.POINTER(PopGridmapConfig), _I, _I, _VP, C.POINTER(_VP)]),
    ("pop_gridmap_destroy", _I, [_VP]),
    ("pop_gridmap_last_error", C.c_char_p, [_VP]),
    ("pop_gridmap_update", _I, [_VP, _PD, _PI32, _PD]),
    ("pop_gridmap_query", _I, [_VP, _I, _PD, _I, _PI8, _PI8]),
    ("pop_gridmap_download", _I, [_VP, _I, _PD, _PI8, _PI32]),
    ("pop_gridmap_dims", _I, [_VP, _PI32]),
    ("pop_gridmap_set_depth_params", _I, [_VP, C.POINTER(PopDepthParams)]),
    ("pop_gridmap_update_depth", _I, [_VP, C.POINTER(C.c_uint16), _PD, _PD]),
    ("pop_gridmap_projected_points", _I, [_VP, _I, _PD, _I, _PI32]),
]

_lib = None


def load() -> C.CDLL:
    """dlopen libpop.so and type every export; raises if the CUDA library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if lib.pop_abi_version() != POP_ABI_VERSION:
        raise ImportError("libpop.so ABI version mismatch")
    _lib = lib
    return lib


def default_config() -> PopConfig:
    cfg = PopConfig()
    load().pop_default_config(C.byref(cfg))
    return cfg


def as_ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))
