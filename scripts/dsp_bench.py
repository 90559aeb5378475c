"""Particle-map (K1d) measurement: agent-updates/s of pop_dsp_update at the reference's grid, and the HBM
roofline of the per-update voxel sweeps on a BASELINE-config-2 sized grid.

  python scripts/dsp_bench.py [--agents 64] [--frames 30] [--big-agents 4]
Prints one JSON line per configuration.  CUDA events on the engine's stream; host->device copies of the
camera clouds are inside the timed region (the ABI call takes host buffers)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import dsp, worldgen
from pred_occ_planner_b200 import scenes as cases


def run(cfg, d, n_agents, frames, warm, label, cpu=False):
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = p.PopEngine(cfg, device=0, stream=stream.cuda_stream)
    maps = dsp.DspMaps(eng, d)
    cyl = cases.dsp_scene(arena=(10.0, 10.0, 3.0))
    seqs = []
    for k in range(warm + frames):
        clouds, poses, quats = [], [], []
        for a in range(n_agents):
            pos, quat, st = cases.dsp_flight(k, speed=1.0 + 0.02 * (a % 16))
            pos = pos + np.array([0.0, 0.2 * (a % 8) - 0.8, 0.0], np.float32)
            clouds.append(worldgen.render_depth_cloud(cyl, pos, quat, t=0.1 * k))
            poses.append(pos); quats.append(quat)
        seqs.append((clouds, np.array(poses), np.array(quats), st))
    for k in range(warm):
        maps.cloudCallback(*seqs[k])
    eng.profile_enable(1)
    eng.profile_read(True)
    l0 = eng.kernel_launches()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(warm, warm + frames):
        maps.cloudCallback(*seqs[k])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms, launches = eng.profile_read(True)
    live = int((maps.particles(0)[:, :, 0] > 0.1).sum()) if cfg.nx * cfg.ny * cfg.nz < 2_000_000 else None
    V, T = cfg.nx * cfg.ny * cfg.nz, cfg.nt
    # algorithmic bytes of the voxel sweeps per agent-update: flags read by the move, commit and occupancy
    # passes (3 x 16 B), list heads read by placement (2 rounds), commit and newborn placement (4 x 4 B),
    # occupancy + mean velocity written (16 B), SOGM accumulators zeroed (4 T B)
    alg = n_agents * V * (3 * 16 + 4 * 4 + 16 + 4 * T)
    dsp_ms = ms["dsp"] / frames
    out = {"workload": label, "agents": n_agents, "grid": [cfg.nx, cfg.ny, cfg.nz, cfg.nt], "frames": frames,
           "ms_per_update_e2e": 1e3 * wall / frames, "ms_per_update_device": dsp_ms,
           "agent_updates_per_s_e2e": n_agents * frames / wall,
           "kernel_launches_per_update": (eng.kernel_launches() - l0) / frames,
           "points_per_cloud": int(np.mean([len(c) for c in seqs[-1][0]])), "live_particles_agent0": live,
           "sweep_algorithmic_bytes": alg, "sweep_GBps_if_all_device_time": alg / (dsp_ms * 1e-3) / 1e9}
    if cpu:
        from tests import oracle_lib
        o = oracle_lib.load().dsp(cfg, d)
        o.set_random_tables(*worldgen.dsp_random_tables(d.n_randoms, 5))
        t0 = time.perf_counter()
        for k in range(warm + frames):
            o.update(seqs[k][0][0], seqs[k][1][0], seqs[k][2][0], seqs[k][3])
            o.export()
        out["cpu_oracle_ms_per_update_1_thread"] = 1e3 * (time.perf_counter() - t0) / (warm + frames)
    print(json.dumps(out), flush=True)
    eng.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=64)
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--big-agents", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    if a.agents > 0:
        cfg = cases.dsp_cfg_ref(n_local=a.agents, n_total=a.agents)
        run(cfg, cases.dsp_params(cfg), a.agents, a.frames, 5, "reference particle map 66x66x20 @0.15 m, T=6", cpu=not a.no_cpu)
    if a.big_agents > 0:
        cfg = cases.dsp_cfg_ref(n_local=a.big_agents, n_total=a.big_agents)
        cfg.nx, cfg.ny, cfg.nz, cfg.nt = 300, 300, 50, 20
        cfg.resolution = 0.1
        d = cases.dsp_params(cfg)
        d.voxel_filter_resolution = 0.1
        run(cfg, d, a.big_agents, 10, 3, "config-2 sized particle map 300x300x50 @0.1 m, T=20")
