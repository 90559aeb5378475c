"""ctypes binding of oracle/_ref/libref_dsp_*.so — the REFERENCE's own particle-map node: plan_env/src/
risk_mapping_node.cpp with plan_env/include/plan_env/dsp_dynamic.h, compiled unmodified from /root/reference
(oracle/Makefile, oracle/ref_glue_dsp.cpp).  Test infrastructure: pins the oracle's A2-A10 restatement (oracle/o_dsp.cpp).

The reference keeps the map in file-static arrays and function-local statics (the first odometry is remembered in
`static float sensor_px_last`), so one loaded library = one map for the life of the process; `fresh()` loads a private
copy of the .so per map.  Returns None where the library was not built."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefDsp:
    def __init__(self, lib, filter_res=0.15, risk_threshold=0.2, sigma_observation=0.1, sigma_localization=0.0):
        self.lib = lib
        d = np.zeros(8, np.int32)
        res = C.c_double(0)
        lib.refdsp_dims(_p(d, C.c_int32), C.byref(res))
        self.nx, self.ny, self.nz, self.nt, self.n_randoms, self.S, self.PCAP, self.NP = (int(v) for v in d)
        self.V = self.nx * self.ny * self.nz
        self.res = res.value
        lib.refdsp_start(C.c_float(filter_res), C.c_float(risk_threshold), C.c_float(sigma_observation),
                         C.c_float(sigma_localization))

    def set_random_tables(self, pg, vg, ur):
        pg = np.ascontiguousarray(pg, np.float32)
        vg = np.ascontiguousarray(vg, np.float32)
        ur = np.ascontiguousarray(ur, np.int32)
        assert len(pg) == self.n_randoms == len(vg), "the library's GAUSSIAN_RANDOMS_NUM is a compile-time constant"
        self.lib.refdsp_set_random_tables(_p(pg, C.c_float), _p(vg, C.c_float), _p(ur, C.c_int32), len(ur))

    def odom(self, pos, quat):
        pos = np.ascontiguousarray(pos, np.float64)
        quat = np.ascontiguousarray(quat, np.float64)
        self.lib.refdsp_odom(_p(pos, C.c_double), _p(quat, C.c_double))

    def cloud(self, cam, stamp):
        cam = np.ascontiguousarray(cam, np.float32).reshape(-1, 3)
        nv = C.c_int32(0)
        return self.lib.refdsp_cloud(_p(cam, C.c_float), len(cam), C.c_double(stamp), C.byref(nv))

    def update(self, cam, pos, quat, stamp):
        """odometry then cloud, as the node receives them; 1 = map updated"""
        self.odom(pos, quat)
        return self.cloud(cam, stamp)

    def filter(self, cam):
        cam = np.ascontiguousarray(cam, np.float32).reshape(-1, 3)
        out = np.zeros((5000, 3), np.float32)
        n = self.lib.refdsp_filter(_p(cam, C.c_float), len(cam), _p(out, C.c_float))
        return out[:n].copy()

    def export(self, thr=None):
        """what publishMap left in risk_maps_: [V][T] (the accumulators are already cleared)"""
        out = np.zeros((self.V, self.nt), np.float32)
        self.lib.refdsp_future_risk(_p(out, C.c_float))
        return out, -1

    def particles(self):
        out = np.zeros((self.V, self.S, 8), np.float32)
        self.lib.refdsp_particles(_p(out, C.c_float))
        return out

    def voxels(self):
        out = np.zeros((self.V, 4), np.float32)
        self.lib.refdsp_voxels(_p(out, C.c_float))
        return out

    def observations(self):
        pc = np.zeros((self.NP, 100, 5), np.float32)
        cnt = np.zeros(self.NP, np.int32)
        ml = np.zeros(self.NP, np.float32)
        self.lib.refdsp_observations(_p(pc, C.c_float), _p(cnt, C.c_int32), _p(ml, C.c_float))
        return pc, cnt, ml

    def pyramids(self):
        out = np.zeros((self.NP, self.PCAP, 3), np.int32)
        self.lib.refdsp_pyramids(_p(out, C.c_int32))
        return out

    def points_with_velocity(self, cap=8192):
        out = np.zeros((cap, 7), np.float32)
        n = self.lib.refdsp_points_with_velocity(_p(out, C.c_float), cap)
        return out[:n].copy()

    def stats(self):
        """position-table, velocity-table and rand() sequence counters, update counter"""
        out = np.zeros(4, np.int32)
        self.lib.refdsp_stats(_p(out, C.c_int32))
        return out


_tmp = []


def fresh(variant="ref", **kw):
    """A new map: a private copy of the library is loaded so that the reference's static state starts from scratch."""
    path = os.path.join(REF_DIR, f"libref_dsp_{variant}.so")
    if not os.path.exists(path):
        return None
    d = tempfile.mkdtemp(prefix="refdsp_")
    _tmp.append(d)
    dst = os.path.join(d, f"libref_dsp_{variant}_{len(_tmp)}.so")
    shutil.copy(path, dst)
    return RefDsp(C.CDLL(dst), **kw)


# ---- seeded flights and the comparison object shared by the generator script, the CPU pins and the GPU test ---------
import hashlib

N_FRAMES = 10
RISK_FRAMES = (1, 4, 9)


def flight_frames(n=N_FRAMES):
    """The 10-frame flight of tests/cases.py through the 6 static + 4 moving cylinders scene: (cam cloud, pos, quat, stamp)."""
    from pred_occ_planner_b200 import worldgen
    from tests import cases
    scene = cases.dsp_scene()
    out = []
    for k in range(n):
        pos, quat, st = cases.dsp_flight(k)
        out.append((worldgen.render_depth_cloud(scene, pos, quat, t=0.1 * k), pos, quat, st))
    return out


def _sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8).copy()


class _Agent:
    """adapts pred_occ_planner_b200.dsp.DspMaps (accessors take the agent) to the single-map accessor names"""
    def __init__(self, maps, agent):
        self.m, self.a = maps, agent

    def __getattr__(self, name):
        f = getattr(self.m, name)
        return lambda *args: f(self.a, *args)


def frame_state(m):
    """Everything bit-exact about a map after one update, as digests plus a few readable numbers."""
    pc, cnt, ml = m.observations()
    mask = np.arange(pc.shape[1])[None, :] < cnt[:, None]
    pa = m.particles()
    live = pa[:, :, 0] > 0.1
    py = m.pyramids()
    pyl = py[:, :, 0] != 0
    pv = m.points_with_velocity()
    vo = m.voxels()
    return {
        "n_obs": np.int64(cnt.sum()), "n_live": np.int64(live.sum()), "n_pyr": np.int64(pyl.sum()), "n_pv": np.int64(len(pv)),
        "obs": _sha(cnt, pc[mask][:, :4], ml),          # points and C_k of every stored observation, max ranges
        "flags": _sha(pa[:, :, 0]),
        "particles": _sha(np.argwhere(live).astype(np.int32), pa[live]),
        "pyramids": _sha(np.argwhere(pyl).astype(np.int32), py[pyl]),
        "velocity": _sha(pv),
        "voxels": _sha(vo),
        "occupancy_sum": np.float64(vo[:, 0].astype(np.float64).sum()),
    }


def sparse(risk):
    r = risk.ravel()
    nz = np.flatnonzero(r).astype(np.uint32)
    return nz, r[nz].astype(np.float32)
