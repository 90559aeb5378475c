"""Host mirror of `GridMap` (plan_env/include/plan_env/grid_map.h): the log-odds raycast map, batched over
drones, over the C ABI (pop_gridmap_*)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PopError, PopGridmapConfig, as_ptr


def default_gridmap_config() -> PopGridmapConfig:
    c = PopGridmapConfig()
    _lib.load().pop_gridmap_default_config(C.byref(c))
    return c


class GridMaps:
    def __init__(self, cfg: PopGridmapConfig | None = None, n_maps: int = 1, device: int = 0, stream: int | None = None):
        self.lib = _lib.load()
        self.cfg = cfg or default_gridmap_config()
        self.n_maps = n_maps
        self._h = C.c_void_p()
        rc = self.lib.pop_gridmap_create(C.byref(self.cfg), n_maps, device, C.c_void_p(stream or 0), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise PopError(rc, "pop_gridmap_create failed (no CUDA device -> NO_DEVICE; there is no CPU fallback)")
        dims = np.zeros(3, np.int32)
        self.lib.pop_gridmap_dims(self._h, as_ptr(dims, C.c_int32))
        self.dims = tuple(int(v) for v in dims)

    def _check(self, rc):
        if rc != 0:
            raise PopError(rc, self.lib.pop_gridmap_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.pop_gridmap_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def updateOccupancy(self, clouds, cam_pos):
        """raycastProcess + clearAndInflateLocalMap for every map: clouds = list of [n,3] world-frame points
        (proj_points_), cam_pos [n_maps,3]."""
        cap = self.cfg.max_points
        buf = np.zeros((self.n_maps, cap, 3))
        counts = np.zeros(self.n_maps, np.int32)
        for m, c in enumerate(clouds):
            c = np.asarray(c, np.float64).reshape(-1, 3)
            if len(c) > cap:
                raise PopError(-4, f"map {m}: {len(c)} points exceed max_points={cap}")
            buf[m, :len(c)] = c
            counts[m] = len(c)
        cam = np.ascontiguousarray(cam_pos, np.float64).reshape(self.n_maps, 3)
        self._check(self.lib.pop_gridmap_update(self._h, as_ptr(buf, C.c_double), as_ptr(counts, C.c_int32),
                                                as_ptr(cam, C.c_double)))

    def getOccupancy(self, m: int, pos):
        return self._query(m, pos)[0]

    def getInflateOccupancy(self, m: int, pos):
        return self._query(m, pos)[1]

    def _query(self, m, pos):
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        o, i = np.zeros(len(pos), np.int8), np.zeros(len(pos), np.int8)
        self._check(self.lib.pop_gridmap_query(self._h, m, as_ptr(pos, C.c_double), len(pos), as_ptr(o, C.c_int8),
                                               as_ptr(i, C.c_int8)))
        return o, i

    def buffers(self, m: int):
        """(occupancy_buffer_ log-odds [nx,ny,nz], occupancy_buffer_inflate_ [nx,ny,nz], local bounds (6,))."""
        n = self.dims[0] * self.dims[1] * self.dims[2]
        occ, inf, b = np.zeros(n), np.zeros(n, np.int8), np.zeros(6, np.int32)
        self._check(self.lib.pop_gridmap_download(self._h, m, as_ptr(occ, C.c_double), as_ptr(inf, C.c_int8),
                                                  as_ptr(b, C.c_int32)))
        return occ.reshape(self.dims), inf.reshape(self.dims), b

    def setDepthParams(self, params):
        self.depth = params
        self._check(self.lib.pop_gridmap_set_depth_params(self._h, C.byref(params)))

    def updateOccupancyFromDepth(self, depth, camera_r, cam_pos):
        """projectDepthImage + raycastProcess + clearAndInflateLocalMap: depth [n_maps, rows, cols] uint16,
        camera_r [n_maps, 3, 3] (camera to world), cam_pos [n_maps, 3]."""
        depth = np.ascontiguousarray(depth, np.uint16).reshape(self.n_maps, self.depth.rows, self.depth.cols)
        R = np.ascontiguousarray(camera_r, np.float64).reshape(self.n_maps, 9)
        cam = np.ascontiguousarray(cam_pos, np.float64).reshape(self.n_maps, 3)
        self._check(self.lib.pop_gridmap_update_depth(self._h, depth.ctypes.data_as(C.POINTER(C.c_uint16)),
                                                      as_ptr(R, C.c_double), as_ptr(cam, C.c_double)))

    def projectedPoints(self, m: int):
        cap = self.cfg.max_points
        out = np.zeros((cap, 3))
        n = C.c_int32(0)
        self._check(self.lib.pop_gridmap_projected_points(self._h, m, as_ptr(out, C.c_double), cap, C.byref(n)))
        return out[:n.value].copy()


def depth_params(rows: int, cols: int, fx: float, fy: float, cx: float, cy: float, use_filter: bool = True,
                 skip_pixel: int = 2, margin: int = 2, mindist: float = 0.2, maxdist: float = 5.0,
                 scaling: float = 1000.0) -> "_lib.PopDepthParams":
    """"grid_map/..." depth parameters (plan_env/src/grid_map.cpp:29-42) for a rows x cols uint16 depth image."""
    p = _lib.PopDepthParams()
    p.fx, p.fy, p.cx, p.cy = fx, fy, cx, cy
    p.k_depth_scaling_factor = scaling
    p.depth_filter_mindist, p.depth_filter_maxdist = mindist, maxdist
    p.use_depth_filter = int(use_filter)
    p.depth_filter_margin, p.skip_pixel, p.rows, p.cols = margin, skip_pixel, rows, cols
    return p

