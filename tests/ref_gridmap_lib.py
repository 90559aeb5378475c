"""ctypes access to oracle/_ref/libref_gridmap.so — the reference's plan_env/src/grid_map.cpp (+ raycast.cpp) compiled
unmodified against stand-ins (oracle/ref_glue_gridmap.cpp) — and the scene the raycast-map pins fly.  Test infrastructure."""
import ctypes as C
import hashlib
import os

import numpy as np

from tests.conftest import ROOT

LIB = os.path.join(ROOT, "oracle", "_ref", "libref_gridmap.so")


class RefGmParams(C.Structure):
    _fields_ = [("resolution", C.c_double), ("map_size", C.c_double * 3), ("local_update_range", C.c_double * 3),
                ("obstacles_inflation", C.c_double),
                ("p_hit", C.c_double), ("p_miss", C.c_double), ("p_min", C.c_double), ("p_max", C.c_double), ("p_occ", C.c_double),
                ("min_ray_length", C.c_double), ("max_ray_length", C.c_double), ("virtual_ceil_height", C.c_double),
                ("ground_height", C.c_double),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("k_depth_scaling_factor", C.c_double), ("depth_filter_mindist", C.c_double),
                ("depth_filter_maxdist", C.c_double), ("depth_filter_tolerance", C.c_double),
                ("local_map_margin", C.c_int), ("use_depth_filter", C.c_int), ("depth_filter_margin", C.c_int),
                ("skip_pixel", C.c_int)]


def available() -> bool:
    return os.path.exists(LIB)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefGridMap:
    """One GridMap of the reference, configured like a pop_gridmap_config + pop_depth_params pair."""

    def __init__(self, cfg, dp):
        self.lib = C.CDLL(LIB)
        self.lib.refgm_create.restype = C.c_void_p
        q = RefGmParams()
        q.resolution = cfg.resolution
        for k in range(3):
            q.map_size[k] = cfg.map_size[k]
            q.local_update_range[k] = cfg.local_update_range[k]
        q.obstacles_inflation = cfg.obstacles_inflation
        q.p_hit, q.p_miss, q.p_min, q.p_max, q.p_occ = cfg.p_hit, cfg.p_miss, cfg.p_min, cfg.p_max, cfg.p_occ
        q.min_ray_length, q.max_ray_length = 0.1, cfg.max_ray_length
        q.virtual_ceil_height, q.ground_height = cfg.virtual_ceil_height, cfg.ground_height
        q.fx, q.fy, q.cx, q.cy = dp.fx, dp.fy, dp.cx, dp.cy
        q.k_depth_scaling_factor = dp.k_depth_scaling_factor
        q.depth_filter_mindist, q.depth_filter_maxdist, q.depth_filter_tolerance = dp.depth_filter_mindist, dp.depth_filter_maxdist, 0.15
        q.local_map_margin, q.use_depth_filter = cfg.local_map_margin, dp.use_depth_filter
        q.depth_filter_margin, q.skip_pixel = dp.depth_filter_margin, dp.skip_pixel
        self.h = C.c_void_p(self.lib.refgm_create(C.byref(q)))
        d = np.zeros(3, np.int32)
        self.lib.refgm_dims(self.h, _p(d, C.c_int32))
        self.dims = tuple(int(v) for v in d)

    def __del__(self):
        try:
            self.lib.refgm_destroy(self.h)
        except Exception:
            pass

    def project(self, img, R, cam, cap=1 << 18):
        img = np.ascontiguousarray(img, np.uint16)
        R = np.ascontiguousarray(R, np.float64).reshape(9)
        cam = np.ascontiguousarray(cam, np.float64).reshape(3)
        out = np.zeros((cap, 3))
        n = self.lib.refgm_project(self.h, _p(img, C.c_uint16), img.shape[0], img.shape[1], _p(R, C.c_double), _p(cam, C.c_double),
                                   _p(out, C.c_double), cap)
        assert n <= cap
        return out[:n].copy()

    def raycast(self, pts=None, cam=None):
        """raycastProcess + clearAndInflateLocalMap on the points the last project() left, or on pts / cam"""
        if pts is None:
            self.lib.refgm_raycast(self.h, None, 0, None)
        else:
            pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
            cam = np.ascontiguousarray(cam, np.float64).reshape(3)
            self.lib.refgm_raycast(self.h, _p(pts, C.c_double), len(pts), _p(cam, C.c_double))

    def buffers(self):
        n = self.dims[0] * self.dims[1] * self.dims[2]
        occ, inf, b = np.zeros(n), np.zeros(n, np.int8), np.zeros(6, np.int32)
        self.lib.refgm_download(self.h, _p(occ, C.c_double), _p(inf, C.c_int8), _p(b, C.c_int32))
        return occ.reshape(self.dims), inf.reshape(self.dims), b

    def query(self, pos, inflate=False):
        pos = np.ascontiguousarray(pos, np.float64).reshape(3)
        return int(self.lib.refgm_query(self.h, _p(pos, C.c_double), int(inflate)))


# ---- the scene of the pins: a depth camera flying past cylinders over a ground plane ----
C2B = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])   # camera z forward = body x, x = -body y, y = -body z
N_FRAMES = 5


def config():
    from pred_occ_planner_b200 import gridmap
    c = gridmap.default_gridmap_config()
    c.map_size[0], c.map_size[1], c.map_size[2] = 20.0, 20.0, 4.0
    c.local_map_margin = 5
    c.max_points = 20000
    return c


def camera():
    import pred_occ_planner_b200 as p
    cam = p._lib.PopCamera()
    p.load().pop_camera_default(C.byref(cam), 160, 120)
    return cam


def scene():
    from pred_occ_planner_b200 import worldgen
    cloud, _ = worldgen.make_world(11, arena=(10.0, 10.0, 3.0), n_static=8, n_moving=0, height=3.0)
    gx, gy = np.meshgrid(np.arange(-5, 5, 0.1), np.arange(-5, 5, 0.1), indexing="ij")
    ground = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1).astype(np.float32)
    return np.concatenate([cloud, ground]).astype(np.float32)


def frame(cam, cloud, k):
    """(uint16 depth image in millimetres, camera rotation, camera position) of frame k"""
    from tests import oracle_lib
    pos = np.array([-4.0 + 0.15 * k, -0.5, 1.05])
    yaw = 0.07 * k
    quat = np.array([np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)])
    depth, _, _ = oracle_lib.render(cam, pos, quat, cloud)
    img = np.where(depth >= 65535, 0, depth).astype(np.uint16)
    Rb = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
    return img, Rb @ C2B, pos


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
