"""ctypes binding of oracle/liboracle.so — the CPU restatement used as the parity checker.
Only tests, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR)
            if f.endswith((".cpp", ".hpp")) and f.startswith(("o_", "oracle"))]   # ref_glue.cpp belongs to _ref/libref.so
    srcs.append(os.path.join(ROOT, "include", "pop.h"))
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        build()
    lib = C.CDLL(LIB)
    lib.orc_map_create.restype = C.c_void_p
    lib.orc_bezier_max_rate.restype = C.c_double
    _lib = OracleLib(lib)
    return _lib


def default_config():
    """pop_config with the reference's simulator defaults, from the oracle (no libpop involved)."""
    from pred_occ_planner_b200._lib import PopConfig
    cfg = PopConfig()
    load().lib.orc_default_config(C.byref(cfg))
    return cfg


def set_resample_noise(z):
    """standard-normal triples consumed by the oracle's peer resampling (replan_risk_rate > 0)"""
    z = np.ascontiguousarray(z, np.float32).reshape(-1, 3)
    load().lib.orc_set_resample_noise(z.ctypes.data_as(C.POINTER(C.c_float)), len(z))


def libstdcxx_normal_triples(seed, n):
    """the (nx, ny, nz) / stddev sequence libstdc++ produces for the reference's three distributions over one engine"""
    out = np.zeros((n, 3), np.float32)
    load().lib.orc_libstdcxx_normal_triples(C.c_ulong(int(seed)), int(n), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleMap:
    """orc::Grid — restatement of FakeParticleRiskVoxel / RiskBase map state and queries."""

    def __init__(self, lib, cfg):
        self.lib, self.cfg = lib, cfg
        self.h = C.c_void_p(lib.orc_map_create(C.byref(cfg)))

    def __del__(self):
        try:
            self.lib.orc_map_destroy(self.h)
        except Exception:
            pass

    def set_pose(self, p):
        p = np.ascontiguousarray(p, np.float32)
        self.lib.orc_map_set_pose(self.h, _p(p, C.c_float))

    def set_risk(self, risk):
        risk = np.ascontiguousarray(risk, np.float32)
        self.lib.orc_map_set_risk(self.h, _p(risk, C.c_float))

    def set_time(self, t_map):
        self.lib.orc_map_set_time(self.h, C.c_double(t_map))

    def kernel_size(self):
        return self.lib.orc_map_kernel_size(self.h)

    def update_gt(self, cloud, cyl, t_map):
        cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 3)
        cyl = np.ascontiguousarray(cyl)
        self.lib.orc_map_update_gt(self.h, _p(cloud, C.c_float), len(cloud), C.c_void_p(cyl.ctypes.data), len(cyl),
                                   C.c_double(t_map))

    def add_other_agents(self, offs, counts, recs, ego_id):
        offs = np.ascontiguousarray(offs, np.float64)
        counts = np.ascontiguousarray(counts, np.int32)
        recs = np.ascontiguousarray(recs)
        self.lib.orc_map_add_other_agents(self.h, _p(offs, C.c_double), _p(counts, C.c_int32),
                                          C.c_void_p(recs.ctypes.data), int(ego_id))

    def add_particles(self, pts, risks, t):
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        risks = np.ascontiguousarray(risks, np.float32)
        self.lib.orc_map_add_particles(self.h, _p(pts, C.c_double), _p(risks, C.c_float), len(pts), int(t))

    def query(self, pos, t):
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        n = len(pos)
        out = np.zeros(n, np.int8)
        if np.issubdtype(np.asarray(t).dtype, np.integer):
            tt = np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.int32), (n,)))
            self.lib.orc_map_query(self.h, _p(pos, C.c_double), _p(tt, C.c_int32), n, _p(out, C.c_int8))
        else:
            tt = np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.float64), (n,)))
            self.lib.orc_map_query_dt(self.h, _p(pos, C.c_double), _p(tt, C.c_double), n, _p(out, C.c_int8))
        return out

    def inflate_query(self, pos, t=None):
        """getInflateOccupancy(pos) / (pos, int t) / (pos, double t)"""
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        n = len(pos)
        out = np.zeros(n, np.int8)
        mode = 0 if t is None else (1 if np.issubdtype(np.asarray(t).dtype, np.integer) else 2)
        tt = np.ascontiguousarray(np.broadcast_to(np.asarray(0.0 if t is None else t, np.float64), (n,)))
        self.lib.orc_map_inflate_query(self.h, mode, _p(pos, C.c_double), _p(tt, C.c_double), n, _p(out, C.c_int8))
        return out

    def traj_safe(self, rec, tau, now, T):
        """isTrajSafe(T) at wall-clock times `now` for the trajectory record `rec` with piece duration tau:
        (safe flags, relative time of the first unsafe sample or -1)."""
        rec = np.ascontiguousarray(rec)
        now = np.ascontiguousarray(np.atleast_1d(now), np.float64)
        T = np.ascontiguousarray(np.broadcast_to(np.asarray(T, np.float64), now.shape))
        out = np.zeros(len(now), np.int8)
        tu = np.zeros(len(now))
        self.lib.orc_traj_safe(self.h, C.c_void_p(rec.ctypes.data), C.c_double(tau), len(now), _p(now, C.c_double),
                               _p(T, C.c_double), _p(out, C.c_int8), _p(tu, C.c_double))
        return out, tu

    def obstacle_points(self, t0, t1, lc, hc, cap=200000):
        lc = np.ascontiguousarray(lc, np.float64)
        hc = np.ascontiguousarray(hc, np.float64)
        out = np.zeros((cap, 3))
        n = self.lib.orc_map_obstacle_points(self.h, C.c_double(t0), C.c_double(t1), _p(lc, C.c_double),
                                             _p(hc, C.c_double), _p(out, C.c_double), cap)
        assert n <= cap
        return out[:n].copy()

    def astar(self, start9, goal3, t_start, cap=64):
        start9 = np.ascontiguousarray(start9, np.float64).reshape(9)
        goal3 = np.ascontiguousarray(goal3, np.float64).reshape(3)
        path = np.zeros((cap, 6))
        ln = C.c_int(0)
        stats = np.zeros(4, np.int32)
        ret = self.lib.orc_astar(self.h, _p(start9, C.c_double), _p(goal3, C.c_double), C.c_double(t_start),
                                 _p(path, C.c_double), cap, C.byref(ln), _p(stats, C.c_int32))
        return ret, path[:ln.value].copy(), stats

    def corridor(self, path, start9, traj_start_time):
        from pred_occ_planner_b200 import POP_MAX_FACES, POP_MAX_PIECES
        path = np.ascontiguousarray(path, np.float64).reshape(-1, 6)
        start9 = np.ascontiguousarray(start9, np.float64).reshape(9)
        hp = np.zeros((POP_MAX_PIECES, POP_MAX_FACES, 4))
        rows = np.zeros(POP_MAX_PIECES, np.int32)
        goal = np.zeros(6)
        n = self.lib.orc_corridor(self.h, _p(path, C.c_double), len(path), _p(start9, C.c_double),
                                  C.c_double(traj_start_time), _p(hp, C.c_double), _p(rows, C.c_int32),
                                  _p(goal, C.c_double))
        return n, hp, rows, goal

    def risk_maps(self):
        c = self.cfg
        out = np.zeros((c.nx * c.ny * c.nz, c.nt), np.float32)
        self.lib.orc_map_download(self.h, _p(out, C.c_float))
        return out


class OracleLib:
    def __init__(self, lib):
        self.lib = lib

    def map(self, cfg) -> OracleMap:
        return OracleMap(self.lib, cfg)

    def dsp(self, cfg, dcfg) -> "OracleDsp":
        return OracleDsp(self.lib, cfg, dcfg)

    def ego_particles(self, size):
        size = np.ascontiguousarray(size, np.float64)
        out = np.zeros((256, 3))
        n = self.lib.orc_ego_particles(_p(size, C.c_double), _p(out, C.c_double), 256)
        return out[:n].copy()

    def linprog(self, c, A, b, seed=1):
        c = np.ascontiguousarray(c, np.float64)
        A = np.ascontiguousarray(A, np.float64).reshape(-1, len(c))
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros(len(c))
        self.lib.orc_linprog.restype = C.c_double
        v = self.lib.orc_linprog(len(c), _p(c, C.c_double), _p(A, C.c_double), _p(b, C.c_double), len(b),
                                 _p(x, C.c_double), C.c_uint64(seed))
        return v, x

    def firi(self, bd, pc, a, b, iterations=2, seed=1, cap=256):
        bd = np.ascontiguousarray(bd, np.float64).reshape(6, 4)
        pc = np.ascontiguousarray(pc, np.float64).reshape(-1, 3)
        a = np.ascontiguousarray(a, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        out = np.zeros((cap, 4))
        rows = self.lib.orc_firi(_p(bd, C.c_double), _p(pc, C.c_double), len(pc), _p(a, C.c_double),
                                 _p(b, C.c_double), int(iterations), C.c_uint64(seed), _p(out, C.c_double), cap)
        return out[:rows].copy()

    @staticmethod
    def pack_polys(polys):
        from pred_occ_planner_b200 import POP_MAX_FACES, POP_MAX_PIECES
        hp = np.zeros((POP_MAX_PIECES, POP_MAX_FACES, 4))
        rows = np.zeros(POP_MAX_PIECES, np.int32)
        for i, h in enumerate(polys):
            h = np.asarray(h, np.float64).reshape(-1, 4)
            hp[i, :len(h)] = h
            rows[i] = len(h)
        return hp, rows

    def qp_setup(self, init, goal, dur, polys, vmax=3.0, amax=3.0):
        M = len(dur)
        hp, rows = self.pack_polys(polys)
        init = np.ascontiguousarray(init, np.float64).reshape(9)
        goal = np.ascontiguousarray(goal, np.float64).reshape(9)
        dur = np.ascontiguousarray(dur, np.float64)
        n = 15 * M
        args = (M, _p(init, C.c_double), _p(goal, C.c_double), _p(dur, C.c_double), _p(hp, C.c_double),
                _p(rows, C.c_int32), C.c_double(vmax), C.c_double(amax))
        m = self.lib.orc_qp_setup(*args, None, None, None, None)
        Q, A, l, u = np.zeros((n, n)), np.zeros((m, n)), np.zeros(m), np.zeros(m)
        self.lib.orc_qp_setup(*args, _p(Q, C.c_double), _p(A, C.c_double), _p(l, C.c_double), _p(u, C.c_double))
        return Q, A, l, u

    def qp_solve(self, cfg, init, goal, dur, polys=None, vmax=3.0, amax=3.0, hp=None, rows=None):
        M = len(dur)
        if hp is None:
            hp, rows = self.pack_polys(polys)
        hp = np.ascontiguousarray(hp, np.float64)
        rows = np.ascontiguousarray(rows, np.int32)
        init = np.ascontiguousarray(init, np.float64).reshape(9)
        goal = np.ascontiguousarray(goal, np.float64).reshape(9)
        dur = np.ascontiguousarray(dur, np.float64)
        cpts = np.zeros((5 * M, 3))
        iters = C.c_int(0)
        st = self.lib.orc_qp_solve(C.byref(cfg), M, _p(init, C.c_double), _p(goal, C.c_double), _p(dur, C.c_double),
                                   _p(hp, C.c_double), _p(rows, C.c_int32), C.c_double(vmax), C.c_double(amax),
                                   _p(cpts, C.c_double), C.byref(iters))
        return st, cpts, iters.value

    def separable(self, A, B):
        A = np.ascontiguousarray(A, np.float64).reshape(-1, 3)
        B = np.ascontiguousarray(B, np.float64).reshape(-1, 3)
        return bool(self.lib.orc_separable(_p(A, C.c_double), len(A), _p(B, C.c_double), len(B)))

    def bezier_eval(self, dur, cpts, t, deriv=0):
        dur = np.ascontiguousarray(dur, np.float64)
        cpts = np.ascontiguousarray(cpts, np.float64).reshape(-1, 3)
        t = np.ascontiguousarray(np.atleast_1d(t), np.float64)
        out = np.zeros((len(t), 3))
        self.lib.orc_bezier_eval(_p(dur, C.c_double), len(dur), _p(cpts, C.c_double), _p(t, C.c_double), len(t),
                                 int(deriv), _p(out, C.c_double))
        return out

    def piece_eval(self, cpts, t0, tf, t, deriv=0):
        cpts = np.ascontiguousarray(cpts, np.float64).reshape(5, 3)
        out = np.zeros(3)
        self.lib.orc_piece_eval(_p(cpts, C.c_double), C.c_double(t0), C.c_double(tf), C.c_double(t), int(deriv),
                                _p(out, C.c_double))
        return out

    def bezier_max_rate(self, dur, cpts, order):
        dur = np.ascontiguousarray(dur, np.float64)
        cpts = np.ascontiguousarray(cpts, np.float64).reshape(-1, 3)
        return self.lib.orc_bezier_max_rate(_p(dur, C.c_double), len(dur), _p(cpts, C.c_double), int(order))

    def bernstein_coeff(self):
        out = np.zeros(25)
        self.lib.orc_bernstein_coeff(_p(out, C.c_double))
        return out.reshape(5, 5)


class OracleDsp:
    """orc::DspMap — sequential restatement of dsp_map::DSPMap + filterPointCloud (oracle/o_dsp.cpp)."""

    def __init__(self, lib, cfg, dcfg):
        self.lib, self.cfg, self.d = lib, cfg, dcfg
        lib.orc_dsp_create.restype = C.c_void_p
        self.h = C.c_void_p(lib.orc_dsp_create(C.byref(cfg), C.byref(dcfg)))
        self.V = cfg.nx * cfg.ny * cfg.nz
        self.S = 2 * dcfg.max_particle_num_voxel
        self.NP = (dcfg.half_fov_h * 2 // dcfg.angle_resolution) * (dcfg.half_fov_v * 2 // dcfg.angle_resolution)

    def __del__(self):
        try:
            self.lib.orc_dsp_destroy(self.h)
        except Exception:
            pass

    def set_random_tables(self, pg, vg, ur):
        pg = np.ascontiguousarray(pg, np.float32)
        vg = np.ascontiguousarray(vg, np.float32)
        ur = np.ascontiguousarray(ur, np.int32)
        self.lib.orc_dsp_set_random_tables(self.h, _p(pg, C.c_float), _p(vg, C.c_float), _p(ur, C.c_int32), len(pg))

    def filter(self, cam):
        cam = np.ascontiguousarray(cam, np.float32).reshape(-1, 3)
        out = np.zeros((self.d.max_points, 3), np.float32)
        n = self.lib.orc_dsp_filter(self.h, _p(cam, C.c_float), len(cam), _p(out, C.c_float))
        return out[:n].copy()

    def update_filtered(self, pts, pos, quat, stamp):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        pos = np.ascontiguousarray(pos, np.float32)
        quat = np.ascontiguousarray(quat, np.float32)
        return self.lib.orc_dsp_update_filtered(self.h, _p(pts, C.c_float), len(pts), _p(pos, C.c_float),
                                                _p(quat, C.c_float), C.c_double(stamp))

    def update(self, cam, pos, quat, stamp):
        cam = np.ascontiguousarray(cam, np.float32).reshape(-1, 3)
        pos = np.ascontiguousarray(pos, np.float32)
        quat = np.ascontiguousarray(quat, np.float32)
        return self.lib.orc_dsp_update(self.h, _p(cam, C.c_float), len(cam), _p(pos, C.c_float),
                                       _p(quat, C.c_float), C.c_double(stamp))

    def export(self, thr=None):
        """[V][T] future risk (clears the accumulators) and the number of occupied voxels."""
        out = np.zeros((self.V, self.cfg.nt), np.float32)
        n = self.lib.orc_dsp_export(self.h, _p(out, C.c_float),
                                    C.c_float(self.d.occupancy_threshold if thr is None else thr))
        return out, n

    def particles(self):
        out = np.zeros((self.V, self.S, 8), np.float32)
        self.lib.orc_dsp_particles(self.h, _p(out, C.c_float))
        return out

    def voxels(self):
        out = np.zeros((self.V, 4), np.float32)
        self.lib.orc_dsp_voxels(self.h, _p(out, C.c_float))
        return out

    def observations(self):
        pc = np.zeros((self.NP, self.d.max_points_per_pyramid, 5), np.float32)
        cnt = np.zeros(self.NP, np.int32)
        ml = np.zeros(self.NP, np.float32)
        self.lib.orc_dsp_observations(self.h, _p(pc, C.c_float), _p(cnt, C.c_int32), _p(ml, C.c_float))
        return pc, cnt, ml

    def pyramids(self):
        out = np.zeros((self.NP, self.d.pyramid_capacity, 3), np.int32)
        self.lib.orc_dsp_pyramids(self.h, _p(out, C.c_int32))
        return out

    def points_with_velocity(self, cap=8192):
        out = np.zeros((cap, 7), np.float32)
        n = self.lib.orc_dsp_points_with_velocity(self.h, _p(out, C.c_float), cap)
        return out[:min(n, cap)].copy()

    def stats(self):
        out = np.zeros(8, np.int32)
        self.lib.orc_dsp_stats(self.h, _p(out, C.c_int32))
        return out


def render(cam, pos, quat, cloud, cap=1 << 20):
    """orc::render_* — the simulator's depth sensor for one agent: (depth_mm [H,W], camera-frame cloud,
    world-frame cloud) (oracle/o_render.cpp)."""
    lib = load().lib
    pos = np.ascontiguousarray(pos, np.float64).reshape(3)
    quat = np.ascontiguousarray(quat, np.float64).reshape(4)
    cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 3)
    depth = np.zeros((cam.height, cam.width), np.int32)
    oc = np.zeros((cap, 3), np.float32)
    ow = np.zeros((cap, 3), np.float32)
    n = lib.orc_render(C.byref(cam), _p(pos, C.c_double), _p(quat, C.c_double), _p(cloud, C.c_float), len(cloud),
                       _p(depth, C.c_int32), _p(oc, C.c_float), _p(ow, C.c_float), cap)
    assert n <= cap
    return depth, oc[:n].copy(), ow[:n].copy()


class OracleGridMap:
    """orc::GridMapO — sequential restatement of GridMap::raycastProcess + clearAndInflateLocalMap
    (oracle/o_gridmap.cpp)."""

    def __init__(self, cfg):
        self.lib = load().lib
        self.lib.orc_gridmap_create.restype = C.c_void_p
        self.h = C.c_void_p(self.lib.orc_gridmap_create(C.byref(cfg)))
        d = np.zeros(3, np.int32)
        self.lib.orc_gridmap_dims(self.h, _p(d, C.c_int32))
        self.dims = tuple(int(v) for v in d)

    def __del__(self):
        try:
            self.lib.orc_gridmap_destroy(self.h)
        except Exception:
            pass

    def update(self, pts, cam):
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        cam = np.ascontiguousarray(cam, np.float64).reshape(3)
        self.lib.orc_gridmap_update(self.h, _p(pts, C.c_double), len(pts), _p(cam, C.c_double))

    def buffers(self):
        n = self.dims[0] * self.dims[1] * self.dims[2]
        occ, inf, b = np.zeros(n), np.zeros(n, np.int8), np.zeros(6, np.int32)
        self.lib.orc_gridmap_download(self.h, _p(occ, C.c_double), _p(inf, C.c_int8), _p(b, C.c_int32))
        return occ.reshape(self.dims), inf.reshape(self.dims), b


def gridmap_project(params, has_first, img, R, cam, max_ray, cap=1 << 20):
    """orc::project_depth — GridMap::projectDepthImage for one map; returns (points, has_first')."""
    lib = load().lib
    img = np.ascontiguousarray(img, np.uint16)
    R = np.ascontiguousarray(R, np.float64).reshape(9)
    cam = np.ascontiguousarray(cam, np.float64).reshape(3)
    out = np.zeros((cap, 3))
    hf = C.c_int(int(has_first))
    n = lib.orc_gridmap_project(C.byref(params), C.byref(hf), img.ctypes.data_as(C.POINTER(C.c_uint16)),
                                _p(R, C.c_double), _p(cam, C.c_double), C.c_double(max_ray), _p(out, C.c_double), cap)
    assert n <= cap
    return out[:n].copy(), bool(hf.value)
