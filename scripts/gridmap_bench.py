"""Raycast-map measurement: map-updates/s of pop_gridmap_update for N maps (depth-camera frames of the reference
camera at 160x120), wall clock around the ABI call (host buffers in), the sequential oracle beside it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import gridmap, worldgen
from tests import oracle_lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 20
c = gridmap.default_gridmap_config()
c.map_size[0], c.map_size[1], c.map_size[2] = 20.0, 20.0, 4.0
c.local_map_margin = 5
cam = p._lib.PopCamera()
p.load().pop_camera_default(p._lib.C.byref(cam), 160, 120)
cloud, _ = worldgen.make_world(11, arena=(10.0, 10.0, 3.0), n_static=8, n_moving=0, height=3.0)
gx, gy = np.meshgrid(np.arange(-5, 5, 0.1), np.arange(-5, 5, 0.1), indexing="ij")
cloud = np.concatenate([cloud, np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1)]).astype(np.float32)
seq = []
for k in range(frames + 3):
    pos = np.array([-4.0 + 0.1 * k, 0.0, 1.0])
    quat = np.array([np.cos(0.02 * k), 0, 0, np.sin(0.02 * k)])
    _, _, pw = oracle_lib.render(cam, pos, quat, cloud)
    seq.append((pos, pw.astype(np.float64)))
maps = gridmap.GridMaps(c, n_maps=n)
for k in range(3):
    maps.updateOccupancy([seq[k][1]] * n, np.tile(seq[k][0], (n, 1)))
t0 = time.perf_counter()
for k in range(3, 3 + frames):
    maps.updateOccupancy([seq[k][1]] * n, np.tile(seq[k][0], (n, 1)))
wall = time.perf_counter() - t0
o = oracle_lib.OracleGridMap(c)
t0 = time.perf_counter()
for k in range(3 + frames):
    o.update(seq[k][1], seq[k][0])
cpu = (time.perf_counter() - t0) / (3 + frames)
print(json.dumps({"workload": f"raycast map 200x200x40 @0.1 m, {n} maps, {len(seq[0][1])} projected points per frame",
                  "ms_per_update_e2e": 1e3 * wall / frames, "map_updates_per_s_e2e": n * frames / wall,
                  "cpu_oracle_ms_per_update_1_thread": 1e3 * cpu, "note": "python-side packing of the point buffers is inside the wall time"}))
