"""Seeded synthetic inputs for the replan hot path (no dataset, no network).

The obstacle model follows the reference simulator's ground-truth publisher:
`MovingCylinder` (uav_simulator/map_generator/include/map_generator/moving_cylinder.h:56-124)
— centre snapped to grid centres, solid disc fill at `res_gen` pitch from z=-0.15 m up, a
+0.01 m offset on every point, per-obstacle constant velocity — and the marker fields the
planner-side map reads (plan_env/src/fake_particle_risk_voxel.cpp:244-264: `w` is the
DIAMETER, `vx, vy` in m/s).  Everything is derived from numpy's Philox generator.
"""
from __future__ import annotations

import math

import numpy as np

from ._lib import CYL_DTYPE, POP_MAX_BODY, POP_MAX_PIECES, TRAJ_DTYPE


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(seed))


def make_world(seed: int, arena=(30.0, 30.0, 5.0), n_static=0, n_moving=40, v_max=1.0,
               w_range=(0.5, 1.0), height=4.0, res_gen=0.1, return_owner=False):
    """Returns (cloud float32 [n,3], cylinders CYL_DTYPE [n_static+n_moving]) and, with return_owner, the index of the
    cylinder each cloud point belongs to (int32 [n]) so that a caller can move the cloud with its obstacles."""
    rng = _rng(seed)
    ax, ay = arena[0] / 2 - 1.0, arena[1] / 2 - 1.0
    n = n_static + n_moving
    cyl = np.zeros(n, dtype=CYL_DTYPE)
    clouds = []
    for i in range(n):
        x = rng.uniform(-ax, ax)
        y = rng.uniform(-ay, ay)
        w = rng.uniform(*w_range)
        vx, vy = (rng.uniform(-v_max, v_max), rng.uniform(-v_max, v_max)) if i >= n_static else (0.0, 0.0)
        x = math.floor(x / res_gen) * res_gen + res_gen / 2.0
        y = math.floor(y / res_gen) * res_gen + res_gen / 2.0
        height_num = math.ceil(height / res_gen)
        wid_num = math.ceil(w / res_gen)
        r0 = int(-wid_num / 2.0)  # C truncation toward zero
        rs = np.arange(r0, math.ceil(wid_num / 2.0 - 1e-12))
        rr, ss = np.meshgrid(rs, rs, indexing="ij")
        keep = (rr * rr + ss * ss) <= (wid_num * wid_num / 4.0)
        rr, ss = rr[keep], ss[keep]
        ts = np.arange(-2, height_num)
        px = np.repeat(x + (rr + 0.5) * res_gen + 1e-2, len(ts))
        py = np.repeat(y + (ss + 0.5) * res_gen + 1e-2, len(ts))
        pz = np.tile((ts + 0.5) * res_gen + 1e-2, len(rr))
        pts = np.stack([px, py, pz], axis=1)
        pts = np.concatenate([[[x, y, height]], pts], axis=0)
        clouds.append(pts.astype(np.float32))
        cyl[i] = (3, 0, x, y, 0.0, w, height, vx, vy, 1.0, 0.0, 0.0, 0.0)
    cloud = np.concatenate(clouds, axis=0) if clouds else np.zeros((0, 3), np.float32)
    if return_owner:
        owner = np.concatenate([np.full(len(c), i, np.int32) for i, c in enumerate(clouds)]) if clouds else np.zeros(0, np.int32)
        return np.ascontiguousarray(cloud), cyl, owner
    return np.ascontiguousarray(cloud), cyl


def advance_world(cloud0: np.ndarray, cyl0: np.ndarray, owner: np.ndarray, dt: float):
    """The ground-truth world dt seconds later: every obstacle translated by its velocity (the reference's simulator
    moves them by a fixed step per 10 Hz tick and reflects at the arena walls, moving_cylinder.h:140-154; the few
    seconds a bench covers stay inside the walls, so the reflection is left out), its cloud points with it."""
    cyl = cyl0.copy()
    cyl["x"] = cyl0["x"] + cyl0["vx"] * dt
    cyl["y"] = cyl0["y"] + cyl0["vy"] * dt
    cloud = cloud0.copy()
    cloud[:, 0] += (cyl0["vx"][owner] * dt).astype(np.float32)
    cloud[:, 1] += (cyl0["vy"][owner] * dt).astype(np.float32)
    return cloud, cyl


def make_rings(seed: int, arena=(30.0, 30.0, 5.0), n=4, v_max=1.0, r_range=(0.7, 2.5), z_range=(0.8, 2.0),
               thickness=0.1, res_gen=0.1):
    """Vertical ring obstacles as the reference's simulator makes them (uav_simulator/map_generator/include/
    map_generator/moving_circle.h:43-110; "mixed" worlds of simulator_fake.launch:43-48): an ellipse of radii
    r +- thickness/2 in a vertical plane turned by a random yaw, one point every res_gen/2 rad.  Their ground-truth
    markers are of type 2 (SPHERE) with scale.x = 2 r2, the ring's pose and quaternion
    (dynamic_forest_seq_sensing.cpp:205-254), which the ground-truth map's ring test reads
    (plan_env/src/fake_particle_risk_voxel.cpp:139-150).  Returns (cloud float32 [m,3], rings CYL_DTYPE [n])."""
    rng = _rng(seed)
    ax, ay = arena[0] / 2 - 1.0, arena[1] / 2 - 1.0
    cyl = np.zeros(n, dtype=CYL_DTYPE)
    clouds = []
    for i in range(n):
        x, y, z = rng.uniform(-ax, ax), rng.uniform(-ay, ay), rng.uniform(*z_range)
        vx, vy = rng.uniform(-v_max, v_max), rng.uniform(-v_max, v_max)
        theta = rng.uniform(0.0, 2 * math.pi)
        r = rng.uniform(*r_range)
        x = math.floor(x / res_gen) * res_gen + res_gen / 2.0
        y = math.floor(y / res_gen) * res_gen + res_gen / 2.0
        z = math.floor(z / res_gen) * res_gen + res_gen / 2.0
        rot = np.array([[math.cos(theta), -math.sin(theta), 0.0], [math.sin(theta), math.cos(theta), 0.0], [0.0, 0.0, 1.0]])
        # q = Quaterniond(rot) * AngleAxisd(pi / 2, UnitY)
        w1, v1 = math.cos(theta / 2), np.array([0.0, 0.0, math.sin(theta / 2)])
        w2, v2 = math.cos(math.pi / 4), np.array([0.0, math.sin(math.pi / 4), 0.0])
        qw = w1 * w2 - float(v1 @ v2)
        qv = w1 * v2 + w2 * v1 + np.cross(v1, v2)
        r1, r2 = r + thickness / 2, r - thickness / 2
        ang = np.arange(0.0, 2 * math.pi, res_gen / 2)
        local = np.stack([np.zeros_like(ang), r1 * np.cos(ang), r2 * np.sin(ang)], axis=1)
        clouds.append((local @ rot.T + [x, y, z]).astype(np.float32))
        cyl[i] = (2, 0, x, y, z, 2 * r2, z, vx, vy, qw, qv[0], qv[1], qv[2])
    cloud = np.concatenate(clouds, axis=0) if clouds else np.zeros((0, 3), np.float32)
    return np.ascontiguousarray(cloud), cyl


def circle_agents(n: int, radius: float, z: float = 1.0):
    """Agents on a circle heading to their antipodes: (start [n,9] pos/vel/acc, goal [n,3])."""
    ang = 2 * np.pi * np.arange(n) / n
    start = np.zeros((n, 9))
    start[:, 0] = radius * np.cos(ang)
    start[:, 1] = radius * np.sin(ang)
    start[:, 2] = z
    goal = np.stack([-start[:, 0], -start[:, 1], np.full(n, z)], axis=1)
    return start, np.ascontiguousarray(goal)


def lattice_agents(n: int, arena=(50.0, 50.0), z: float = 1.0, seed: int = 0):
    """Agents on a square lattice with shuffled goals (BASELINE config 3 style)."""
    side = int(math.ceil(math.sqrt(n)))
    xs = np.linspace(-arena[0] / 2 + 2, arena[0] / 2 - 2, side)
    ys = np.linspace(-arena[1] / 2 + 2, arena[1] / 2 - 2, side)
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    p = np.stack([gx.ravel(), gy.ravel()], axis=1)[:n]
    start = np.zeros((n, 9))
    start[:, :2] = p
    start[:, 2] = z
    perm = _rng(seed).permutation(n)
    goal = np.concatenate([p[perm], np.full((n, 1), z)], axis=1)
    return start, np.ascontiguousarray(goal)


def ego_particles(size) -> np.ndarray:
    """ParticleATC::initEgoParticles (traj_coordinator/src/particles.cpp:62-75): body cube
    sampled every 0.15 m with the reference's accumulating double loop."""
    out = []
    step = 0.15
    x = -size[0] / 2
    while x <= size[0] / 2:
        y = -size[1] / 2
        while y <= size[1] / 2:
            z = -size[2] / 2
            while z <= size[2] / 2:
                out.append((x, y, z))
                z += step
            y += step
        x += step
    return np.array(out, dtype=np.float64).reshape(-1, 3)


def body_table(n_total: int, size, ego_id: int | None = None):
    """particles_buffer_ as every agent sees it: its own offsets exact, peers' offsets rounded
    through float32 (`geometry_msgs/Point32`, particles.cpp:90-123)."""
    p = ego_particles(size)
    if len(p) > POP_MAX_BODY:
        raise ValueError("too many body particles")
    offs = np.zeros((n_total, POP_MAX_BODY, 3))
    offs[:, :len(p)] = p.astype(np.float32).astype(np.float64)
    if ego_id is not None:
        offs[ego_id, :len(p)] = p
    counts = np.full(n_total, len(p), dtype=np.int32)
    return offs, counts


def straight_trajs(start_pos: np.ndarray, goal: np.ndarray, t_start: float, speed: float,
                   n_pieces: int, piece_dur: float) -> np.ndarray:
    """Constant-velocity Bézier trajectories towards each goal, as TRAJ_DTYPE records."""
    n = len(start_pos)
    rec = np.zeros(n, dtype=TRAJ_DTYPE)
    for i in range(n):
        d = goal[i] - start_pos[i]
        nrm = np.linalg.norm(d)
        v = d / nrm * speed if nrm > 1e-9 else np.zeros(3)
        rec["drone_id"][i] = i
        rec["traj_id"][i] = 1
        rec["n_pieces"][i] = n_pieces
        rec["valid"][i] = 1
        rec["start_time"][i] = t_start
        for p in range(n_pieces):
            rec["duration"][i, p] = piece_dur
            p0 = start_pos[i] + v * piece_dur * p
            for k in range(5):
                rec["cpts"][i, p * 5 + k] = p0 + v * piece_dur * k / 4.0
    return rec


def quat_yaw(yaw: float) -> np.ndarray:
    """(w, x, y, z) of a rotation about +z."""
    return np.array([math.cos(yaw / 2), 0.0, 0.0, math.sin(yaw / 2)], np.float32)


def _quat_matrix(q) -> np.ndarray:
    w, x, y, z = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def render_depth_cloud(cyl, pos, quat, t: float = 0.0, half_fov=(43.0, 29.0), pitch_deg: float = 0.5,
                       max_range: float = 6.0, ground_z: float = 0.0, noise: float = 0.0, seed: int = 0):
    """What the reference's depth camera would publish on "map/cloud" (uav_simulator/local_sensing): one
    ray per `pitch_deg` over the field of view, nearest hit with the vertical cylinders (moved to time t)
    or the ground plane, returned in the CAMERA frame (x right, y down, z forward) that
    filterPointCloud expects (risk_mapping_node.cpp:82-84: body x = z_c, y = -x_c, z = -y_c)."""
    pos = np.asarray(pos, np.float64)
    R = _quat_matrix(quat)
    az = np.deg2rad(np.arange(-half_fov[0] + pitch_deg / 2, half_fov[0], pitch_deg))
    el = np.deg2rad(np.arange(-half_fov[1] + pitch_deg / 2, half_fov[1], pitch_deg))
    A, E = np.meshgrid(az, el, indexing="ij")
    d_body = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
    d = d_body @ R.T
    best = np.full(len(d), np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (ground_z - pos[2]) / d[:, 2]
    tg = np.where((tg > 0) & np.isfinite(tg), tg, np.inf)
    best = np.minimum(best, tg)
    for c in cyl:
        cx, cy = c["x"] + c["vx"] * t, c["y"] + c["vy"] * t
        r = c["w"] / 2
        ox, oy = pos[0] - cx, pos[1] - cy
        a = d[:, 0] ** 2 + d[:, 1] ** 2
        b = 2 * (ox * d[:, 0] + oy * d[:, 1])
        cc = ox * ox + oy * oy - r * r
        disc = b * b - 4 * a * cc
        with np.errstate(divide="ignore", invalid="ignore"):
            th = (-b - np.sqrt(np.where(disc >= 0, disc, np.nan))) / (2 * a)
        zh = pos[2] + th * d[:, 2]
        ok = np.isfinite(th) & (th > 0) & (zh >= ground_z) & (zh <= c["h"])
        best = np.minimum(best, np.where(ok, th, np.inf))
    hit = best < max_range
    pts_body = d_body[hit] * best[hit, None]
    if noise > 0:
        pts_body = pts_body + _rng(seed).normal(0, noise, pts_body.shape)
    cam = np.stack([-pts_body[:, 1], -pts_body[:, 2], pts_body[:, 0]], axis=1)
    return np.ascontiguousarray(cam, np.float32)


def dsp_random_tables(n: int, seed: int, p_std: float = 0.01, v_std: float = 0.01):
    """The three noise inputs of the DSP map (dsp_dynamic.h:114-116, 1220-1231, 1687-1689): position and
    velocity Gaussian tables and the rand() stream (int32 in [0, RAND_MAX])."""
    rng = _rng(seed)
    pg = rng.normal(0.0, p_std, n).astype(np.float32)
    vg = rng.normal(0.0, v_std, n).astype(np.float32)
    ur = rng.integers(0, 2 ** 31 - 1, n, dtype=np.int64).astype(np.int32)
    return pg, vg, ur
