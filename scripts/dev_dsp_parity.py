"""Frame-by-frame report of the particle-map parity (device vs sequential oracle) with timings; development aid."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import dsp, worldgen
from tests import cases, dsp_compare, oracle_lib

orc = oracle_lib.load()
nA = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = cases.dsp_cfg_ref(n_local=nA, n_total=nA)
d = cases.dsp_params(cfg)
eng = p.PopEngine(cfg, device=0)
maps = dsp.DspMaps(eng, d)
tabs = worldgen.dsp_random_tables(d.n_randoms, 5)
maps.set_random_tables(*tabs)
orcs = [orc.dsp(cfg, d) for _ in range(nA)]
for o in orcs:
    o.set_random_tables(*tabs)
cyl = cases.dsp_scene()
for k in range(frames):
    clouds, poses, quats = [], [], []
    for a in range(nA):
        pos, quat, st = cases.dsp_flight(k, speed=1.0 + 0.5 * a)
        pos = pos + np.array([0.0, 0.8 * a, 0.1 * a], np.float32)
        clouds.append(worldgen.render_depth_cloud(cyl, pos, quat, t=0.1 * k))
        poses.append(pos); quats.append(quat)
    t0 = time.time()
    try:
        ok = maps.cloudCallback(clouds, np.array(poses), np.array(quats), st)
    except Exception as e:
        print("frame", k, "device error:", e)
        break
    eng.sync()
    t1 = time.time()
    for a in range(nA):
        orcs[a].update(clouds[a], poses[a], quats[a], st)
        bad = dsp_compare.compare_frame(maps, a, orcs[a], eng.agent_map(a).risk_maps(), d.occupancy_threshold)
        print(f"frame {k} agent {a}: ok={ok[a]} gpu {1e3*(t1-t0):.2f} ms stats {maps.stats(a).tolist()}", "PARITY" if not bad else "MISMATCH")
        for b in bad:
            print("    ", b)
