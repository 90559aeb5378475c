"""Host mirror of the particle-map side of plan_env: `risk_mapping_node` + `dsp_map::DSPMap`
(plan_env/src/risk_mapping_node.cpp:155-188, plan_env/include/plan_env/dsp_dynamic.h:165-364) over the C ABI.

`DspMaps` attaches one particle map per local agent to a `PopEngine`; `cloudCallback` is the per-frame entry
point (camera cloud + odometry in, SOGM updated on the device).  The getters keep the reference's names.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PopDspConfig, as_ptr


def default_dsp_config(cfg) -> PopDspConfig:
    d = PopDspConfig()
    _lib.load().pop_dsp_default_config(C.byref(cfg), C.byref(d))
    return d


class DspMaps:
    def __init__(self, eng, dcfg: PopDspConfig | None = None):
        self.eng = eng
        self.d = (dcfg or default_dsp_config(eng.cfg)).copy()
        eng._check(eng.lib.pop_dsp_attach(eng._h, C.byref(self.d)))
        c = eng.cfg
        self.V = c.nx * c.ny * c.nz
        self.S = 2 * self.d.max_particle_num_voxel
        self.NP = (self.d.half_fov_h * 2 // self.d.angle_resolution) * (self.d.half_fov_v * 2 // self.d.angle_resolution)
        self.n_local = eng.n_local

    def set_random_tables(self, pg, vg, ur):
        pg = np.ascontiguousarray(pg, np.float32)
        vg = np.ascontiguousarray(vg, np.float32)
        ur = np.ascontiguousarray(ur, np.int32)
        e = self.eng
        e._check(e.lib.pop_dsp_set_random_tables(e._h, as_ptr(pg, C.c_float), as_ptr(vg, C.c_float),
                                                 as_ptr(ur, C.c_int32), len(pg)))

    def _pack(self, clouds, cap):
        buf = np.zeros((self.n_local, cap, 3), np.float32)
        counts = np.zeros(self.n_local, np.int32)
        for a, c in enumerate(clouds):
            c = np.asarray(c, np.float32).reshape(-1, 3)
            buf[a, :len(c)] = c
            counts[a] = len(c)
        return buf, counts

    def cloudCallback(self, clouds, poses, quats, stamp: float, filtered: bool = False):
        """One frame for every local agent: clouds = list of [n,3] camera-frame arrays (or body-frame
        `valid_clouds_` when filtered=True), poses [n_local,3], quats [n_local,4] (w,x,y,z).  Returns the
        per-agent accepted flags (DSPMap::update's return value)."""
        cap = self.d.max_points if filtered else self.d.max_input_points
        buf, counts = self._pack(clouds, cap)
        return self.cloudCallbackPacked(buf, counts, poses, quats, stamp, filtered)

    def pack(self, clouds, filtered: bool = False):
        """(buffer [n_local, capacity, 3] float32, counts) in the layout pop_dsp_update takes."""
        return self._pack(clouds, self.d.max_points if filtered else self.d.max_input_points)

    def cloudCallbackPacked(self, buf, counts, poses, quats, stamp: float, filtered: bool = False):
        e = self.eng
        poses = np.ascontiguousarray(poses, np.float32).reshape(self.n_local, 3)
        quats = np.ascontiguousarray(quats, np.float32).reshape(self.n_local, 4)
        ok = np.zeros(self.n_local, np.int32)
        fn = e.lib.pop_dsp_update_filtered if filtered else e.lib.pop_dsp_update
        e._check(fn(e._h, as_ptr(buf, C.c_float), as_ptr(counts, C.c_int32), as_ptr(poses, C.c_float),
                    as_ptr(quats, C.c_float), float(stamp), as_ptr(ok, C.c_int32)))
        e._poses = poses.copy()
        e._t_map = float(stamp)
        return ok

    def filterPointCloud(self, agent: int, cam):
        cam = np.ascontiguousarray(cam, np.float32).reshape(-1, 3)
        out = np.zeros((self.d.max_points, 3), np.float32)
        n = C.c_int32(0)
        e = self.eng
        e._check(e.lib.pop_dsp_filter(e._h, agent, as_ptr(cam, C.c_float), len(cam), as_ptr(out, C.c_float), C.byref(n)))
        return out[:n.value].copy()

    # -- depth sensor (uav_simulator/local_sensing) ----------------------------------------------
    def set_camera(self, width: int = 160, height: int = 120, cam: "_lib.PopCamera | None" = None):
        """DepthRender::set_para (depth_render.cu:83-91): the reference camera scaled to width x height."""
        e = self.eng
        if cam is None:
            cam = _lib.PopCamera()
            e.lib.pop_camera_default(C.byref(cam), int(width), int(height))
        e._check(e.lib.pop_render_set_camera(e._h, C.byref(cam)))
        self.camera = cam
        return cam

    def render(self, pos, quat, want_depth: bool = False):
        """render_currentpose + render_pcl_world for every local agent from the engine's world cloud; the
        camera-frame clouds stay on the device for `cloudCallbackRendered`.  Returns (counts, depth_mm|None)."""
        e = self.eng
        pos = np.ascontiguousarray(pos, np.float64).reshape(self.n_local, 3)
        quat = np.ascontiguousarray(quat, np.float64).reshape(self.n_local, 4)
        counts = np.zeros(self.n_local, np.int32)
        depth = np.zeros((self.n_local, self.camera.height, self.camera.width), np.int32) if want_depth else None
        e._check(e.lib.pop_render_depth(e._h, as_ptr(pos, C.c_double), as_ptr(quat, C.c_double),
                                        as_ptr(depth, C.c_int32) if want_depth else None, as_ptr(counts, C.c_int32)))
        return counts, depth

    def rendered_cloud(self, agent: int):
        cap = self.d.max_input_points
        out = np.zeros((cap, 3), np.float32)
        n = C.c_int32(0)
        e = self.eng
        e._check(e.lib.pop_render_download_cloud(e._h, agent, as_ptr(out, C.c_float), cap, C.byref(n)))
        return out[:n.value].copy()

    def cloudCallbackRendered(self, poses, quats, stamp: float):
        """cloudCallback on the clouds the last `render` left on the device."""
        e = self.eng
        poses = np.ascontiguousarray(poses, np.float32).reshape(self.n_local, 3)
        quats = np.ascontiguousarray(quats, np.float32).reshape(self.n_local, 4)
        ok = np.zeros(self.n_local, np.int32)
        e._check(e.lib.pop_dsp_update_rendered(e._h, as_ptr(poses, C.c_float), as_ptr(quats, C.c_float), float(stamp),
                                               as_ptr(ok, C.c_int32)))
        e._poses = poses.copy()
        e._t_map = float(stamp)
        return ok

    # -- parity views ---------------------------------------------------------------------------
    def particles(self, agent: int):
        out = np.zeros((self.V, self.S, 8), np.float32)
        e = self.eng
        e._check(e.lib.pop_dsp_download_particles(e._h, agent, as_ptr(out, C.c_float)))
        return out

    def voxels(self, agent: int):
        out = np.zeros((self.V, 4), np.float32)
        e = self.eng
        e._check(e.lib.pop_dsp_download_voxels(e._h, agent, as_ptr(out, C.c_float)))
        return out

    def observations(self, agent: int):
        pc = np.zeros((self.NP, self.d.max_points_per_pyramid, 5), np.float32)
        cnt = np.zeros(self.NP, np.int32)
        ml = np.zeros(self.NP, np.float32)
        e = self.eng
        e._check(e.lib.pop_dsp_download_observations(e._h, agent, as_ptr(pc, C.c_float), as_ptr(cnt, C.c_int32),
                                                     as_ptr(ml, C.c_float)))
        return pc, cnt, ml

    def pyramids(self, agent: int):
        out = np.zeros((self.NP, self.d.pyramid_capacity, 3), np.int32)
        e = self.eng
        e._check(e.lib.pop_dsp_download_pyramids(e._h, agent, as_ptr(out, C.c_int32)))
        return out

    def points_with_velocity(self, agent: int, cap: int = 8192):
        out = np.zeros((cap, 7), np.float32)
        n = C.c_int32(0)
        e = self.eng
        e._check(e.lib.pop_dsp_download_points_with_velocity(e._h, agent, as_ptr(out, C.c_float), cap, C.byref(n)))
        return out[:n.value].copy()

    def stats(self, agent: int):
        out = np.zeros(8, np.int32)
        e = self.eng
        e._check(e.lib.pop_dsp_stats(e._h, agent, as_ptr(out, C.c_int32)))
        return out

    def getOccupancyMap(self, agent: int, cap: int | None = None):
        """Centres of the occupied voxels (dsp_dynamic.h:402-420), map frame."""
        cap = cap or self.V
        out = np.zeros((cap, 3), np.float32)
        n = C.c_int32(0)
        e = self.eng
        e._check(e.lib.pop_dsp_occupied(e._h, agent, as_ptr(out, C.c_float), cap, C.byref(n)))
        return out[:n.value].copy()
