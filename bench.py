#!/usr/bin/env python
"""bench.py — agent-replans/s of the per-replan hot path (SOGM update → hybrid A* → corridors →
Bezier QP → peer projection/separation) on N B200s, one process per GPU.

Workload: a BASELINE.json config (default configs[1] = "c2": 64 agents per GPU, 30x30x5 m arena, 300x300x50 voxels
@0.1 m, 20 time slices of 0.1 s, 40 moving cylinders, 10 Hz replan; `--config c3|c4` for the larger ones) flown as a
CLOSED LOOP: a step = one replan cycle of every local agent at simulated time T0 + 0.1 s * step; between steps the
obstacles move with their ground-truth velocities (the cloud with them), every agent moves along the trajectory it
planned (perfect tracking, what uav_simulator/fake_drone does) and the freshly planned trajectories are exchanged.
Both arms fly the same script from the same initial state: steps 0-1 pre-roll, W warm-up steps, K timed steps.

`value` times the K steps with every input resident in HBM (world states staged on the device beforehand, agent
states advanced on the device; CUDA events on the stream the kernels run on, max over ranks).  `e2e` times the same
cycle through the C-ABI call a plan_manager would make (pop_replan_cycle) with pinned HOST buffers: cloud, obstacle
states and agent states uploaded, status and trajectories downloaded, every step, inside the timed region.
`--impl reference` flies the same closed loop with the CPU oracle (oracle/liboracle.so: the reference's path restated
and pinned against the reference's own sources compiled here; the reference as a whole needs
ROS/Eigen/PCL/OSQP/GLPK) on all host threads.  The GPU line's `cpu_baseline` is that same oracle cycle run on the
exact inputs of one GPU step, which also yields the `parity` object (statuses and control points compared).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# libpop pipelines the planning stages over up to 32 side streams; the hardware queue count must be
# chosen before torch creates the CUDA context
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "agent-replans/s (SOGM+A*+corridor+QP cycle)"
UNIT = "agent-replans/s"
DT = 0.1          # replan period (10 Hz)
LOOKAHEAD = 0.02  # traj_start_time - map stamp
# N > 1: the allgather of cycle s overlaps cycle s+1 and is consumed by cycle s+2 — a broadcast one replan period in
# flight, as the reference's "/broadcast_traj" messages are; POP_BENCH_EXCHANGE=strict makes every cycle wait for it
PIPELINED = os.environ.get("POP_BENCH_EXCHANGE", "pipelined") == "pipelined"
T0 = 100.0
PREROLL = 2

# BASELINE.json configs[1..3]; SURVEY §8d gives the synthetic inputs
CONFIGS = {
    "c2": dict(grid=(300, 300, 50, 20), arena=(30.0, 30.0, 5.0), n_static=0, n_moving=40, agents=64, layout="circle",
               radius=13.0, pitch=64.0, seed=1236,
               label="c2: 64 agents/GPU, 300x300x50 @0.1 m, T=20, 40 moving cylinders, 10 Hz"),
    "c3": dict(grid=(500, 500, 60, 20), arena=(50.0, 50.0, 6.0), n_static=60, n_moving=60, agents=256, layout="lattice",
               pitch=128.0, seed=1237,
               label="c3: 256 agents/GPU, 500x500x60 @0.1 m, T=20, 60 static + 60 moving cylinders, 10 Hz"),
    "c4": dict(grid=(500, 500, 60, 20), arena=(50.0, 50.0, 6.0), n_static=60, n_moving=60, agents=128, layout="lattice",
               pitch=128.0, seed=1238,
               label="c4: 128 agents/GPU (1 024 on 8 GPUs), c3 arena per GPU, T=20, allgather each cycle, 10 Hz"),
    "small": dict(grid=(120, 120, 30, 10), arena=(12.0, 12.0, 3.0), n_static=0, n_moving=10, agents=8, layout="circle",
                  radius=4.5, pitch=32.0, seed=1236, label="small harness smoke (not a bench number)"),
}


def arena_origin(rank: int, world: int, pitch: float):
    """Centre of rank `rank`'s arena: slot `rank` of a side x side tiling (side = ceil(sqrt(world))) centred on the
    origin, row-major.  Stacking arenas in z would keep them closer to the origin, but the reference's corridor stage
    clamps its boxes to z in [0, 4] m (baseline_fake.cpp:324-325), so every arena has to stay on the ground floor.  POP_BENCH_ARENA_SHIFT (diagnosis) moves every arena further out along x by that many pitches."""
    side = 1
    while side * side < world:
        side += 1
    ix, iy = rank % side, rank // side
    shift = int(os.environ.get("POP_BENCH_ARENA_SHIFT", "0"))
    return ((ix - 0.5 * (side - 1) + shift) * pitch, (iy - 0.5 * (side - 1)) * pitch)


class Scenario:
    """One rank's share of a config: configuration, world script, initial agent states."""

    def __init__(self, name: str, n_local: int, n_total: int, id_base: int, default_config=None):
        import pred_occ_planner_b200 as p
        k = CONFIGS[name]
        self.name, self.k = name, k
        c = (default_config or p.default_config)()   # the CPU arm passes the oracle's: it never loads libpop
        c.nx, c.ny, c.nz, c.nt = k["grid"]
        c.resolution = 0.1
        c.time_resolution = 0.1
        c.resolution_astar = 0.1
        c.time_resolution_astar = 0.2
        c.corridor_tau = 0.2
        c.cell_dtype = p.POP_CELL_U8
        c.n_agents_local, c.n_agents_total, c.agent_id_base = n_local, n_total, id_base
        c.max_obstacle_points = 16384
        c.max_cloud_points = 1 << 19
        self.cfg = c
        world = n_total // n_local
        rank = id_base // n_local
        cloud, cyl, owner = p.worldgen.make_world(k["seed"], arena=k["arena"], n_static=k["n_static"], n_moving=k["n_moving"],
                                                  return_owner=True)
        # every rank flies its own copy of the arena (weak scaling as BASELINE config 4 words it: "arena per GPU"),
        # translated to its slot of a square tiling centred on the origin; all trajectories are still exchanged and
        # every rank's peer projection scans all records.  The tiling is kept compact on purpose: the reference sets
        # OSQP's eps_rel = 1e-3 on an unscaled termination test in WORLD coordinates (iosqp.hpp:87-89), so the corridor
        # rows may be violated by 1e-3 x |coordinate| — 1.5 cm in the reference's own +-15 m arenas, 22 cm for an arena
        # 224 m out, where planned trajectories then touch their neighbours' and the separation test rejects them.
        self.origin = arena_origin(rank, world, k["pitch"])
        cloud[:, 0] += np.float32(self.origin[0])
        cloud[:, 1] += np.float32(self.origin[1])
        cyl["x"] += self.origin[0]
        cyl["y"] += self.origin[1]
        self.cloud0, self.cyl0, self.owner = cloud, cyl, owner
        self.max_abs_vy = float(np.abs(cyl["vy"]).max()) if len(cyl) else 0.0
        if k["layout"] == "circle":
            s0, g0 = p.worldgen.circle_agents(n_local, k["radius"])
        else:
            s0, g0 = p.worldgen.lattice_agents(n_local, arena=k["arena"][:2], seed=3)
        start = np.tile(s0, (world, 1))
        goal = np.tile(g0, (world, 1))
        for r in range(world):
            ox, oy = arena_origin(r, world, k["pitch"])
            for arr in (start, goal):
                arr[r * n_local:(r + 1) * n_local, 0] += ox
                arr[r * n_local:(r + 1) * n_local, 1] += oy
        self.start_all, self.goal_all = start, goal
        self.body = p.worldgen.body_table(n_total, [0.4, 0.4, 0.45])

    def t(self, step: int) -> float:
        return T0 + DT * step

    def world(self, step: int):
        """(cloud, cylinders) of this rank's arena at step `step`"""
        import pred_occ_planner_b200 as p
        return p.worldgen.advance_world(self.cloud0, self.cyl0, self.owner, DT * step)


def advance_states(orc, start, recs, t_next):
    """Perfect tracking on the host (what pop_swarm_advance does on the device, k_follow.cu:k_swarm_advance): every
    agent takes the commanded state of its own last trajectory at t_next; past its end it holds the end point; without
    one it stays."""
    out = start.copy()
    for a in range(len(start)):
        r = recs[a]
        M = int(r["n_pieces"])
        if M <= 0 or t_next < r["start_time"]:
            out[a, 3:] = 0.0
            continue
        dur = r["duration"][:M].astype(np.float64)
        tt = t_next - float(r["start_time"])
        T = 0.0
        for x in dur:
            T += float(x)
        if tt <= T:
            for d in range(3):
                out[a, 3 * d:3 * d + 3] = orc.bezier_eval(dur, r["cpts"][:5 * M], [tt], d)[0]
        else:
            out[a, :3] = orc.bezier_eval(dur, r["cpts"][:5 * M], [T], 0)[0]
            out[a, 3:] = 0.0
    return out


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_voxel_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the voxel-update launches of one cycle, parsed from the newest
    committed `ncu --set full` summary (profiles/*voxel_update.txt, written by scripts/ncu_summary.py); None if there
    is none.  Only meaningful for the c2 workload at 64 agents, which is what those captures are of."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*voxel_update.txt")))
    if not files:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_kernel = {}
    name = None
    for ln in open(files[-1]):
        if ln.startswith("## "):
            m = re.search(r"k_vox_\w+", ln)
            name = m.group(0) if m else None
        m = re.match(r"(dram__bytes_(?:read|write)\.sum)\s+([\d.,]+)\s+(\w+)", ln)
        if m and name:
            # one launch of each voxel-update kernel per cycle: keep the last capture of every kernel
            per_kernel.setdefault(name, {})[m.group(1)] = float(m.group(2).replace(",", "")) * unit.get(m.group(3), 1.0)
    total = sum(sum(v.values()) for v in per_kernel.values())
    return (int(total) if total > 0 else None), os.path.basename(files[-1])


# ------------------------------------------------------------------------------------------------
def cpu_cycle(sc, cloud, cyl, start, goal, recs, n_sample, threads, t_map, records=None):
    """Oracle cycle (map update + peers + A* + corridor + QP + separation) for the first n_sample
    local agents on `threads` host threads. Returns (seconds, status)."""
    from tests import oracle_cycle
    cfg = sc.cfg
    a0 = cfg.agent_id_base
    poses = start[:, :3].astype(np.float32)
    return oracle_cycle.cycle_batch(cfg, a0, a0 + n_sample, threads, cloud, cyl, poses, (sc.body[0], sc.body[1], recs),
                                    start, goal, t_map, t_map + LOOKAHEAD, records=records)


def run_reference(args, rank, world):
    """The reference arm: the same closed loop flown by the CPU oracle on all host threads.  At N > 1 only rank 0's
    arena is flown (the arenas are independent copies), as a bounded sample of the N-GPU workload.  No GPU, no libpop
    compute: pred_occ_planner_b200 is imported for the configuration struct and the world generator only."""
    if rank != 0:
        return
    import pred_occ_planner_b200 as p
    from tests import oracle_lib
    orc = oracle_lib.load()
    n_local = args.agents
    sc = Scenario(args.config, n_local, n_local * world, 0, default_config=oracle_lib.default_config)
    threads = os.cpu_count() or 1
    start = sc.start_all[:n_local].copy()
    goal = sc.goal_all[:n_local].copy()
    recs_all = np.zeros(n_local * world, p.TRAJ_DTYPE)
    own = np.zeros(n_local, p.TRAJ_DTYPE)
    n_steps = PREROLL + args.warmup + args.steps
    # bound the run to a few minutes: fly the whole batch if one step fits the budget, else the first n_sample agents
    cloud, cyl = sc.world(0)
    probe_n = min(n_local, threads)
    sec, _ = cpu_cycle(sc, cloud, cyl, start, goal, recs_all, probe_n, threads, sc.t(0))
    budget = 170.0 / n_steps
    n_sample = int(max(probe_n, min(n_local, budget / max(sec / probe_n, 1e-6))))
    ok = 0
    dt_timed = 0.0
    for s in range(n_steps):
        cloud, cyl = sc.world(s)
        if s > 0:
            start = advance_states(orc, start, own, sc.t(s))
        t0 = time.perf_counter()
        _, st = cpu_cycle(sc, cloud, cyl, start, goal, recs_all, n_sample, threads, sc.t(s), records=own)
        el = time.perf_counter() - t0
        recs_all[:n_local] = own    # "/broadcast_traj": peers see this cycle's trajectories in the next one
        if s >= PREROLL + args.warmup:
            dt_timed += el
            ok += int((st[:n_sample] == 0).sum())
    value = n_sample * args.steps / dt_timed
    sample = (f"{n_sample} of {n_local} agents of rank 0's arena per step, closed loop, {args.steps} timed steps after "
              f"{PREROLL}+{args.warmup} untimed, {threads} host threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt_timed / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 grid / f64 planning", "data": "synthetic",
            "config": config_dict(sc, world, None, None),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "replans_ok_frac": ok / max(1, n_sample * args.steps),
            "note": "CPU oracle = the reference's path restated in C++ and pinned against the reference's own sources "
                    "compiled here (DESIGN.md §2); the reference as a whole needs ROS/Eigen/PCL/OSQP/GLPK"}
    emit(line)


def config_dict(sc, world, cloud_points, rec_bytes):
    """The `config` object; identical keys in both arms."""
    c = sc.cfg
    V = c.nx * c.ny * c.nz
    return {"workload": sc.k["label"], "name": sc.name, "agents_per_gpu": c.n_agents_local,
            "agents_total": c.n_agents_local * world, "grid": [c.nx, c.ny, c.nz, c.nt], "cell": "u8 (GPU) / f32 (CPU)",
            "loop": "closed: obstacles move, agents track their planned trajectories, trajectories exchanged every step",
            "parallelism": f"agents sharded x{world}, one arena per GPU",
            "l2": "inputs larger than L2 (per-step SOGM writes %.2f GB per GPU)" % (c.n_agents_local * V * c.nt / 1e9),
            "collective": ("1 allgather of 1536 B trajectory records per cycle"
                           + (", pipelined: own records at once, the other ranks' arrive one cycle later "
                              "(pop_exchange_trajs_pipelined)" if PIPELINED else ", consumed by the next cycle")) if world > 1
            else "single rank: device-to-device publish"}


# ------------------------------------------------------------------------------------------------
def dsp_leg(stream, device, with_cpu, n_agents=64, frames=20, warm=5):
    """Secondary measurement: the particle-map producer of the SOGM (real-perception path)
    for 64 agents at the reference's own grid (66x66x20 @0.15 m, 6 future slices): one pop_dsp_update per frame
    with HOST camera clouds (host->device copies inside the timed region), the sequential oracle beside it."""
    import pred_occ_planner_b200 as p
    from pred_occ_planner_b200 import dsp, worldgen, scenes as cases
    import torch
    cfg = cases.dsp_cfg_ref(n_local=n_agents, n_total=n_agents)
    d = cases.dsp_params(cfg)
    eng = p.PopEngine(cfg, device=device, stream=stream.cuda_stream)
    maps = dsp.DspMaps(eng, d)
    cyl = cases.dsp_scene()
    seq = []
    for k in range(warm + frames):
        clouds, poses, quats = [], [], []
        for a in range(n_agents):
            pos, quat, st = cases.dsp_flight(k, speed=1.0 + 0.02 * (a % 16))
            pos = pos + np.array([0.0, 0.2 * (a % 8) - 0.8, 0.0], np.float32)
            clouds.append(worldgen.render_depth_cloud(cyl, pos, quat, t=0.1 * k))
            poses.append(pos)
            quats.append(quat)
        buf, counts = maps.pack(clouds)
        buf = torch.from_numpy(buf).pin_memory().numpy()   # pinned host staging, as a camera driver would provide
        seq.append((buf, counts, np.array(poses), np.array(quats), st, clouds[0]))
    for k in range(warm):
        maps.cloudCallbackPacked(*seq[k][:5])
    eng.profile_enable(1)
    eng.profile_read(True)
    l0 = eng.kernel_launches()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(warm, warm + frames):
        maps.cloudCallbackPacked(*seq[k][:5])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms, _ = eng.profile_read(True)
    out = {"workload": "particle map 66x66x20 @0.15 m, T=6, depth-camera clouds, %d agents" % n_agents,
           "agent_updates_per_s_e2e": n_agents * frames / wall, "ms_per_update_e2e": 1e3 * wall / frames,
           "ms_per_update_device": ms["dsp"] / frames, "launches_per_update": (eng.kernel_launches() - l0) / frames,
           "raw_points_per_cloud": int(np.mean(seq[-1][1])),
           "h2d_bytes_per_update": int(12 * seq[-1][1].sum() + n_agents * (12 + 16 + 4))}
    eng.close()
    if with_cpu:
        from tests import oracle_lib
        o = oracle_lib.load().dsp(cfg, d)
        o.set_random_tables(*worldgen.dsp_random_tables(d.n_randoms, 5))
        t0 = time.perf_counter()
        for k in range(warm + frames):
            o.update(seq[k][5], seq[k][2][0], seq[k][3][0], seq[k][4])
            o.export()
        cpu_ms = 1e3 * (time.perf_counter() - t0) / (warm + frames)
        out["cpu_baseline"] = {"value": 1e3 / cpu_ms, "unit": "agent-updates/s", "cores": 1, "kind": "port",
                               "sample": "agent 0's %d frames, sequential oracle" % (warm + frames)}
    return out


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import pred_occ_planner_b200 as p

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version/debug banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_total = args.agents * world
    id_base, n_local = p.swarm.shard(n_total, world, rank)
    sc = Scenario(args.config, n_local, n_total, id_base)
    cfg = sc.cfg
    sl = slice(id_base, id_base + n_local)
    start0, goal = sc.start_all[sl].copy(), sc.goal_all[sl].copy()
    W, K = args.warmup, args.steps
    n_steps = PREROLL + W + K
    # a dedicated non-default stream: libpop enqueues on it and the CUDA events below time it
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    rec_bytes = p.swarm.RECORD_BYTES

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    # the world script, staged once: pinned host copies for the e2e path, device copies for the resident path
    worlds = [sc.world(s) for s in range(n_steps + 4)]
    keep, h_clouds, h_cyls, d_clouds, d_cyls = [], [], [], [], []
    for cloud, cyl in worlds:
        tc, hc = pinned(cloud)
        ty, hy = pinned(cyl.view(np.uint8))
        keep += [tc, ty]
        h_clouds.append(hc)
        h_cyls.append(hy.view(p.CYL_DTYPE))
        d_clouds.append(tc.to(f"cuda:{local_rank}", non_blocking=True))
        d_cyls.append(ty.to(f"cuda:{local_rank}", non_blocking=True))
    torch.cuda.synchronize()
    _s, h_start = pinned(start0)
    _g, h_goal = pinned(goal)
    _p, h_poses = pinned(start0[:, :3].astype(np.float32))
    n_cloud = len(worlds[0][0])
    h2d = h_clouds[0].nbytes + h_cyls[0].nbytes + h_start.nbytes + h_goal.nbytes + h_poses.nbytes
    d2h = n_local * (4 + rec_bytes) + h_start.nbytes   # statuses + trajectory records + the agents' new states

    def new_engine():
        eng = p.PopEngine(cfg, device=local_rank, stream=stream.cuda_stream)
        eng.set_body_particles(*sc.body)
        eng.set_trajectories(np.zeros(n_total, p.TRAJ_DTYPE))
        # '/broadcast_traj': one allgather of fixed-size trajectory records per cycle, issued on the
        # current torch stream (= the engine's) straight between libpop's device buffers
        xchg = p.swarm.TrajExchange(n_total, world, rank, engine=eng, device=local_rank, pipelined=PIPELINED)
        exchange = xchg.step
        if os.environ.get("POP_BENCH_EXCHANGE") == "local":   # diagnosis only: no collective
            exchange = eng.traj_publish_local
        return eng, exchange

    # ================= e2e: host buffers in, results out, every step =================
    eng, exchange = new_engine()

    def e2e_step(s):
        if s > 0:   # odometry: the drone tracked its trajectory; the planner reads its state back
            eng.swarm_advance(sc.t(s))
            h_start[:] = eng.swarm_states()
            h_poses[:] = h_start[:, :3]
        eng.set_world_cloud(h_clouds[s])
        eng.set_cylinders(h_cyls[s])
        st, tr = eng.replan_cycle(h_start, h_goal, h_poses, sc.t(s), sc.t(s) + LOOKAHEAD)
        exchange()
        return st, tr

    for s in range(PREROLL + W):
        e2e_step(s)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_begin = time.perf_counter()
    ok_e2e = 0
    for s in range(PREROLL + W, n_steps):
        ok_e2e += int((e2e_step(s)[0] == 0).sum())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = time.perf_counter() - t_begin
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())

    # ---- parity + CPU baseline (rank 0 at N = 1): one more step, the oracle on exactly its inputs ----
    parity = cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        from tests import oracle_lib
        orc = oracle_lib.load()
        s = n_steps
        eng.swarm_advance(sc.t(s))
        h_start[:] = eng.swarm_states()
        h_poses[:] = h_start[:, :3]
        peers = eng.get_trajectories()          # the table this step's map update reads
        eng.set_world_cloud(h_clouds[s])
        eng.set_cylinders(h_cyls[s])
        st_gpu, tr_gpu = eng.replan_cycle(h_start, h_goal, h_poses, sc.t(s), sc.t(s) + LOOKAHEAD)
        threads = os.cpu_count() or 1
        own = np.zeros(n_local, p.TRAJ_DTYPE)
        sec, st_cpu = cpu_cycle(sc, worlds[s][0], worlds[s][1], h_start.copy(), goal, peers, n_local, threads, sc.t(s), records=own)
        both = (st_gpu == 0) & (st_cpu == 0)
        rel = 0.0
        for a in np.flatnonzero(both):
            M = int(own["n_pieces"][a])
            if M != int(tr_gpu["n_pieces"][a]):
                rel = float("inf")
                continue
            x, y = tr_gpu["cpts"][a][:5 * M], own["cpts"][a][:5 * M]
            rel = max(rel, float(np.abs(x - y).max() / max(1.0, np.abs(y).max())))
        parity = {"agents": int(n_local), "status_equal": int((st_gpu == st_cpu).sum()), "both_ok": int(both.sum()),
                  "max_rel_cpts": rel, "tolerance": 1e-4, "step": s,
                  "against": "CPU oracle (port) on the same cloud, obstacle states, agent states and peer table"}
        cpu_base = {"value": n_local / sec, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": f"all {n_local} agents of step {s} (the parity step), one cycle, {threads} host threads"}
    eng.close()

    # ================= device-resident: the same script, nothing crosses PCIe =================
    eng, exchange = new_engine()

    marks = []   # (before the exchange, after it) per timed step: what a rank spends waiting for the others

    def resident_step(s, mark=False):
        if s > 0:
            eng.swarm_advance(sc.t(s))
        eng.set_world_cloud_dev(d_clouds[s].data_ptr(), n_cloud)
        eng.set_cylinders_dev(d_cyls[s].data_ptr(), len(sc.cyl0), sc.max_abs_vy)
        eng.replan_cycle_resident(sc.t(s), sc.t(s) + LOOKAHEAD)
        if mark:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        exchange()
        if mark:
            e1.record(stream)
            marks.append((e0, e1))

    # step 0 places the agents (host states once); afterwards everything stays on the device
    eng.set_world_cloud(h_clouds[0])
    eng.set_cylinders(h_cyls[0])
    eng.replan_cycle(start0, goal, start0[:, :3].astype(np.float32), sc.t(0), sc.t(0) + LOOKAHEAD)
    exchange()
    # mode 2: CUDA events around the map stages only (the roofline kernel); the planning stages keep
    # running pipelined over agent groups exactly as they do with profiling off
    eng.profile_enable(2)
    for s in range(1, PREROLL + W):
        resident_step(s)
    eng.profile_read(True)
    launches0 = eng.kernel_launches()
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    host_t0 = time.perf_counter()
    for s in range(PREROLL + W, n_steps):
        resident_step(s, mark=world > 1)
    host_enqueue_ms = 1e3 * (time.perf_counter() - host_t0) / K   # host time to enqueue one cycle
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        # own work per step = the step minus the time spent inside the exchange (waiting for the slowest rank)
        own = ev0.elapsed_time(ev1) - sum(a.elapsed_time(b) for a, b in marks)
        allown = torch.zeros(world, device="cuda", dtype=torch.float64)
        allown[rank] = own / K
        dist.all_reduce(allown)
        own_ms_by_rank = [round(float(v), 3) for v in allown.tolist()]
    else:
        own_ms_by_rank = None
    stage_ms, stage_launches = eng.profile_read(True)
    launches = eng.kernel_launches() - launches0
    status, _ = eng.cycle_results()
    ok_frac = float((status == 0).mean())
    # untimed extra pass with every stage bracketed (planning stages serialised): where the time goes
    eng.profile_enable(1)
    n_serial = 3
    for i in range(n_serial):
        resident_step(n_steps + i)
    serial_ms, _ = eng.profile_read(True)
    eng.profile_enable(0)

    if rank == 0:
        ms_step = ms_total / K
        value = n_total * K / (ms_total * 1e-3)
        e2e_value = n_total * K / e2e_s
        peak, peak_src = peaks()
        cell = 1
        V = cfg.nx * cfg.ny * cfg.nz
        vox_ms = stage_ms["voxel_update"] / K
        # algorithmic bytes of the voxel update: every SOGM cell written once + cloud read
        # once per agent + scattered future-slice deposits (<= one per cloud point and slice)
        alg_bytes = n_local * (V * cfg.nt * cell + n_cloud * 12 + n_cloud * (cfg.nt - 1) * cell)
        achieved = alg_bytes / (vox_ms * 1e-3) / 1e9 if vox_ms > 0 else 0.0
        traffic, traffic_src = ncu_voxel_traffic() if (sc.name == "c2" and n_local == 64) else (None, None)
        conf = config_dict(sc, world, n_cloud, rec_bytes)
        conf["cloud_points"] = int(n_cloud)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 grid / f64 planning", "data": "synthetic",
            "config": conf,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_s / K, "replans_ok_frac": ok_e2e / (n_local * K)},
            "gpu_launches": int(launches),
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": clocks,
            "roofline": {"kernel": "voxel update (k_vox_count + k_vox_scan + k_vox_fill + k_vox_tiles)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": int(alg_bytes), "ms_per_step": vox_ms},
            "stage_ms_serialised": {k: v / n_serial for k, v in serial_ms.items()},
            "own_ms_per_step_by_rank": own_ms_by_rank,
            "stage_note": "untimed extra pass with the planning stages run one after the other over the whole batch; "
                          "in the timed region they are pipelined over agent groups on side streams",
            "replans_ok_frac": ok_frac,
            "parity": parity,
            "cpu_baseline": cpu_base,
        }
    eng.close()
    if rank == 0:
        if world == 1 and sc.name == "c2" and not args.no_dsp:
            line["dsp"] = dsp_leg(stream, local_rank, not args.no_cpu_baseline)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """stdout must carry exactly one JSON line: anything libraries print there (NCCL's version banner, torch
    warnings) is redirected to stderr for the whole run; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json workload (default configs[1])")
    ap.add_argument("--agents", type=int, default=None, help="agents per GPU (default: the config's)")
    ap.add_argument("--small", action="store_true", help="tiny grid (harness smoke only; not a bench number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dsp", action="store_true", help="skip the particle-map leg")
    args = ap.parse_args()
    if args.small:
        args.config = "small"
    if args.agents is None:
        args.agents = CONFIGS[args.config]["agents"]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
