"""ctypes binding of oracle/_ref/libref_map_*.so — the REFERENCE's own map classes (plan_env/src/map.cpp +
fake_particle_risk_voxel.cpp or risk_base.cpp, with traj_coordinator/src/particles.cpp), compiled unmodified from
/root/reference at one grid size per library (oracle/Makefile, oracle/ref_glue_map.cpp).  Test infrastructure: pins the
oracle's SOGM builders, queries, obstacle-point scan and peer deposit.  `load(variant)` returns None where the library
was not built; the tests then fall back to tests/golden/ref_map_pins.npz (scripts/make_golden_ref_map.py)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

# variant -> (kind, nx, ny, nz, resolution, nt): must match REF_MAP_VARIANTS in oracle/Makefile
VARIANTS = {
    "fake_c1": (0, 66, 66, 20, 0.15, 9),       # the reference's simulator grid (map_parameters.h + fake_particle_risk_voxel.h:28)
    "fake_c2s": (0, 120, 120, 30, 0.1, 10),    # BASELINE config-2 geometry on a smaller grid (tests/cases.cfg_small_c2)
    "fake_c2t": (0, 160, 160, 30, 0.1, 20),    # config-2 geometry, 20 slices
    "risk_ref": (1, 66, 66, 20, 0.15, 6),      # RiskBase on the particle map's grid (map_parameters.h)
}


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefMap:
    def __init__(self, lib, cfg):
        self.lib = lib
        dims = np.zeros(5, np.int32)
        res = C.c_double(0)
        lib.refmap_dims(_p(dims, C.c_int32), C.byref(res))
        self.nx, self.ny, self.nz, self.nt, self.kind = (int(v) for v in dims)
        assert (cfg.nx, cfg.ny, cfg.nz, cfg.nt) == (self.nx, self.ny, self.nz, self.nt), "config does not match the library's grid"
        assert abs(cfg.resolution - res.value) < 1e-6
        params = np.array([cfg.time_resolution, cfg.risk_threshold, cfg.clearance, cfg.ground_height, cfg.ceiling_height,
                           cfg.risk_threshold_region, cfg.risk_region_decay, cfg.risk_voxel_decay], np.float64)
        lib.refmap_create.restype = C.c_void_p
        self.h = C.c_void_p(lib.refmap_create(_p(params, C.c_double)))
        self.cfg = cfg

    def __del__(self):
        try:
            self.lib.refmap_destroy(self.h)
        except Exception:
            pass

    def kernel(self):
        out = np.zeros((4096, 3), np.int32)
        n = self.lib.refmap_kernel(self.h, _p(out, C.c_int32), 4096)
        return out[:n].copy()

    def set_pose(self, p):
        p = np.ascontiguousarray(p, np.float64)
        self.lib.refmap_set_pose(self.h, _p(p, C.c_double))

    def set_cylinders(self, cyl):
        cyl = np.ascontiguousarray(cyl)
        self.lib.refmap_set_cylinders(self.h, C.c_void_p(cyl.ctypes.data), len(cyl))

    def set_swarm(self, n, ego_id, size, offs, counts, recs, now, replan_risk_rate=0.0, num_resample=10):
        size = np.ascontiguousarray(size, np.float64)
        offs = np.ascontiguousarray(offs, np.float64)
        counts = np.ascontiguousarray(counts, np.int32)
        recs = np.ascontiguousarray(recs)
        self.lib.refmap_set_swarm(self.h, int(n), int(ego_id), _p(size, C.c_double), C.c_float(replan_risk_rate),
                                  int(num_resample), _p(offs, C.c_double), offs.shape[1], _p(counts, C.c_int32),
                                  C.c_void_p(recs.ctypes.data), C.c_double(now))

    def update(self, cloud, now):
        cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 3)
        self.lib.refmap_update(self.h, _p(cloud, C.c_float), len(cloud), C.c_double(now))

    def future_risk(self, risk, stamp, now=None):
        risk = np.ascontiguousarray(risk, np.float32)
        self.lib.refmap_future_risk(self.h, _p(risk, C.c_float), C.c_double(stamp), C.c_double(stamp if now is None else now))

    def risk_maps(self):
        out = np.zeros((self.nx * self.ny * self.nz, self.nt), np.float32)
        self.lib.refmap_download(self.h, _p(out, C.c_float))
        return out

    def set_risk(self, risk):
        risk = np.ascontiguousarray(risk, np.float32)
        self.lib.refmap_upload(self.h, _p(risk, C.c_float))

    def set_time(self, t):
        self.lib.refmap_set_map_time(self.h, C.c_double(t))

    def query(self, pos, t):
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        n = len(pos)
        out = np.zeros(n, np.int8)
        if np.issubdtype(np.asarray(t).dtype, np.integer):
            tt = np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.int32), (n,)))
            self.lib.refmap_query(self.h, _p(pos, C.c_double), _p(tt, C.c_int32), n, _p(out, C.c_int8))
        else:
            tt = np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.float64), (n,)))
            self.lib.refmap_query_dt(self.h, _p(pos, C.c_double), _p(tt, C.c_double), n, _p(out, C.c_int8))
        return out

    def inflate_query(self, pos, t=None):
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        n = len(pos)
        out = np.zeros(n, np.int8)
        if t is None:
            mode, tt = 0, np.zeros(n)
        elif np.issubdtype(np.asarray(t).dtype, np.integer):
            mode, tt = 1, np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.float64), (n,)))
        else:
            mode, tt = 2, np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.float64), (n,)))
        self.lib.refmap_inflate_query(self.h, mode, _p(pos, C.c_double), _p(tt, C.c_double), n, _p(out, C.c_int8))
        return out

    def obstacle_points(self, t0, t1, lc, hc, cap=400000):
        lc = np.ascontiguousarray(lc, np.float64)
        hc = np.ascontiguousarray(hc, np.float64)
        out = np.zeros((cap, 3))
        n = self.lib.refmap_obstacle_points(self.h, C.c_double(t0), C.c_double(t1), _p(lc, C.c_double), _p(hc, C.c_double),
                                            _p(out, C.c_double), cap)
        assert n <= cap
        return out[:n].copy()

    def add_particles(self, pts, risks, t):
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        risks = np.ascontiguousarray(risks, np.float32)
        self.lib.refmap_add_particles(self.h, _p(pts, C.c_double), _p(risks, C.c_float), len(pts), int(t))

    def add_obstacles(self, pts, t):
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        self.lib.refmap_add_obstacles(self.h, _p(pts, C.c_double), len(pts), int(t))

    def filter_cloud(self, cam, filter_res):
        cam = np.ascontiguousarray(cam, np.float32).reshape(-1, 3)
        out = np.zeros((5000, 3), np.float32)
        n = self.lib.refmap_filter_cloud(self.h, _p(cam, C.c_float), len(cam), C.c_float(filter_res), _p(out, C.c_float))
        return out[:n].copy()

    def index_math(self, local, idx):
        local = np.ascontiguousarray(local, np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, np.int32)
        oi = np.zeros(len(local), np.int32)
        inr = np.zeros(len(local), np.int8)
        pos = np.zeros((len(idx), 3), np.float32)
        self.lib.refmap_index_math(self.h, _p(local, C.c_float), len(local), _p(oi, C.c_int32), _p(inr, C.c_int8),
                                   _p(idx, C.c_int32), len(idx), _p(pos, C.c_float))
        return oi, inr, pos


_libs = {}


def load(variant):
    """The library of one grid size, or None where it was not built (the GPU box gets the built files;
    a checkout without /root/reference has none)."""
    if variant in _libs:
        return _libs[variant]
    path = os.path.join(REF_DIR, f"libref_map_{variant}.so")
    lib = C.CDLL(path) if os.path.exists(path) else None
    _libs[variant] = lib
    return lib


def cfg_for(variant, p=None):
    """pop_config matching a variant's compile-time grid (the other fields are the reference defaults)."""
    import pred_occ_planner_b200 as pp
    kind, nx, ny, nz, res, nt = VARIANTS[variant]
    c = pp.default_config()
    c.nx, c.ny, c.nz, c.nt = nx, ny, nz, nt
    c.resolution = res
    c.cell_dtype = pp.POP_CELL_F32
    if kind == 1:
        c.kernel_shape = pp.POP_KERNEL_CUBE
        c.query_variant = pp.POP_QUERY_RISKBASE
    return c


# ---- seeded cases shared by scripts/make_golden_ref_map.py, tests/test_ref_map_pins.py and the GPU tests ------------
BODY_SIZE = [0.4, 0.4, 0.45]


def curved_trajs(rng, n, t_start, n_pieces, dur):
    """Random-walk Bezier control points (pieces need not join: only evaluation matters here)."""
    import pred_occ_planner_b200 as pp
    rec = np.zeros(n, dtype=pp.TRAJ_DTYPE)
    for i in range(n):
        rec["drone_id"][i] = i
        rec["traj_id"][i] = 1
        rec["n_pieces"][i] = n_pieces
        rec["valid"][i] = 1
        rec["start_time"][i] = t_start[i]
        p0 = rng.uniform([-2, -2, 0.6], [2, 2, 1.6])
        for q in range(n_pieces):
            rec["duration"][i, q] = dur * (1 + 0.3 * rng.uniform())
            for k in range(5):
                p0 = p0 + rng.uniform(-0.15, 0.25, 3) * [1, 1, 0.2]
                rec["cpts"][i, q * 5 + k] = p0
    return rec


def map_case(name):
    """One pinned scene: dict(variant, cfg, pose, t_map, cloud, cyl | risk, swarm | None, box)."""
    import pred_occ_planner_b200 as pp
    from pred_occ_planner_b200 import worldgen
    from tests import cases
    spec = {
        # name: (variant, seed, arena, n_static, n_moving, n_rings, peer)
        "c1_pillars_rings": ("fake_c1", 11, (10.0, 10.0, 3.0), 6, 6, 4, False),
        "c1_moving_peer": ("fake_c1", 12, (10.0, 10.0, 3.0), 4, 8, 0, True),
        "c2s_moving": ("fake_c2s", 13, (12.0, 12.0, 3.0), 0, 14, 2, True),
        "c2t_moving": ("fake_c2t", 14, (16.0, 16.0, 3.0), 4, 16, 3, False),
        "risk_peer": ("risk_ref", 15, None, 0, 0, 0, True),
    }[name]
    variant, seed, arena, n_static, n_moving, n_rings, peer = spec
    rng = np.random.default_rng(seed)
    kind = VARIANTS[variant][0]
    if variant == "fake_c1":
        cfg = cases.cfg_c1(pp.POP_CELL_F32)
    elif variant == "risk_ref":
        cfg = cases.dsp_cfg_ref()
    else:
        _, nx, ny, nz, _, nt = VARIANTS[variant]
        cfg = cases.cfg_small_c2(pp.POP_CELL_F32, 1, 1, nx, ny, nz, nt)
    case = dict(name=name, variant=variant, cfg=cfg, kind=kind, seed=seed)
    case["pose"] = rng.uniform([-1.5, -1.5, 0.8], [1.5, 1.5, 1.4])
    case["t_map"] = 50.0 + float(rng.uniform(0, 1))
    if kind == 0:
        cloud, cyl = worldgen.make_world(seed, arena=arena, n_static=n_static, n_moving=n_moving)
        if n_rings:
            rc, rings = worldgen.make_rings(seed + 1000, arena=arena, n=n_rings)
            cloud = np.concatenate([cloud, rc])
            h = n_rings // 2
            cyl = np.concatenate([rings[:h], cyl, rings[h:]])   # ring ids before and after the cylinders: ownership order matters
        case["cloud"], case["cyl"] = np.ascontiguousarray(cloud), cyl
    else:
        V = cfg.nx * cfg.ny * cfg.nz
        risk = np.zeros((V, cfg.nt), np.float32)
        idx = rng.integers(0, V, 12000)
        risk[idx, :] = rng.uniform(0, 0.5, (len(idx), cfg.nt)).astype(np.float32)
        case["risk"] = risk
    if peer:
        cfg.n_agents_total = 2
        ego = seed % 2
        ts = [case["t_map"] + rng.uniform(-1.2, 0.3), case["t_map"] + rng.uniform(-1.2, 0.3)]
        recs = curved_trajs(rng, 2, ts, int(rng.integers(2, 6)), 0.3)
        offs, counts = worldgen.body_table(2, BODY_SIZE, ego_id=ego)
        case["swarm"] = (ego, offs, counts, recs)
    else:
        case["swarm"] = None
    half = np.array([cfg.nx, cfg.ny, cfg.nz]) // 2 * cfg.resolution
    n = 6000
    pos = case["pose"] + rng.uniform(-1.1, 1.1, (n, 3)) * half
    pos[:, 2] = rng.uniform(-0.4, 3.4, n)
    case["q_pos"] = pos
    case["q_t"] = rng.integers(0, cfg.nt, n).astype(np.int32)
    case["q_dt"] = rng.uniform(0.0, (cfg.nt + 1) * cfg.time_resolution, n)
    case["q_dt_in"] = rng.uniform(0.0, (cfg.nt - 1) * cfg.time_resolution, n)   # keeps ceil(t / dt) inside the grid
    case["boxes"] = [(case["t_map"] + 0.05, case["t_map"] + 0.05 + 2.5 * cfg.time_resolution,
                      case["pose"] - [1.6, 1.6, 0.8], case["pose"] + [1.6, 1.6, 0.8]),
                     (case["t_map"] + 1.5 * cfg.time_resolution, case["t_map"] + (cfg.nt - 1.2) * cfg.time_resolution,
                      case["pose"] - [0.9, 2.4, 0.5], case["pose"] + [2.4, 0.9, 1.0])]
    return case


MAP_CASES = ["c1_pillars_rings", "c1_moving_peer", "c2s_moving", "c2t_moving", "risk_peer"]


def run_reference(case):
    """Outputs of the compiled reference for one case (None where the library is absent)."""
    lib = load(case["variant"])
    if lib is None:
        return None
    cfg = case["cfg"]
    rm = RefMap(lib, cfg)
    rm.set_pose(case["pose"])
    if case["swarm"] is not None:
        ego, offs, counts, recs = case["swarm"]
        rm.set_swarm(2, ego, BODY_SIZE, offs, counts, recs, case["t_map"])
    if case["kind"] == 0:
        rm.set_cylinders(case["cyl"])
        rm.update(case["cloud"], case["t_map"])
    else:
        rm.future_risk(case["risk"], case["t_map"])
    return collect(rm, case)


def run_oracle(orc, case):
    cfg = case["cfg"]
    om = orc.map(cfg)
    om.set_pose(case["pose"].astype(np.float32))
    if case["kind"] == 0:
        om.update_gt(case["cloud"], case["cyl"], case["t_map"])
    else:
        om.set_risk(case["risk"])
        om.set_time(case["t_map"])
    if case["swarm"] is not None:
        ego, offs, counts, recs = case["swarm"]
        om.add_other_agents(offs, counts, recs, ego)
    return collect(om, case)


def collect(m, case):
    """The comparison object: sparse grid, the three query kinds, the obstacle-point scans, getInflateOccupancy."""
    r = m.risk_maps().ravel()
    nz = np.flatnonzero(r).astype(np.uint32)
    out = {"nz": nz, "val": r[nz].astype(np.float32)}
    out["q_t"] = m.query(case["q_pos"], case["q_t"])
    out["q_dt"] = m.query(case["q_pos"], case["q_dt"])
    for k, (t0, t1, lc, hc) in enumerate(case["boxes"]):
        out[f"obs{k}"] = m.obstacle_points(t0, t1, lc, hc).astype(np.float32)   # voxel corners: fp32 values held in doubles
    if case["kind"] == 0:   # MapBase::risk_maps_ is the Fake map's grid; RiskBase shadows it (risk_base.h:46)
        out["inf0"] = m.inflate_query(case["q_pos"])
        out["inf_t"] = m.inflate_query(case["q_pos"], case["q_t"])
        out["inf_dt"] = m.inflate_query(case["q_pos"], case["q_dt_in"])
    return out


def resample_case(seed):
    """Two-agent swarm whose peer is in flight at the map time (replan_risk_rate > 0 scenes; the same inputs feed the
    live reference comparison and the GPU test)."""
    from pred_occ_planner_b200 import worldgen
    rng = np.random.default_rng(seed)
    ego = seed % 2
    t_map = 50.0
    ts = [t_map + rng.uniform(-1.5, -0.2), t_map + rng.uniform(-1.5, -0.2)]
    recs = curved_trajs(rng, 2, ts, int(rng.integers(3, 7)), 0.4)
    offs, counts = worldgen.body_table(2, BODY_SIZE, ego_id=ego)
    pose = rng.uniform([-1, -1, 0.8], [1, 1, 1.3])
    return dict(ego=ego, t_map=t_map, recs=recs, offs=offs, counts=counts, pose=pose)
