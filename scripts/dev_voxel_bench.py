"""Dev helper: time the voxel-update stage (K1f) for the clear variants. Not the bench."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pred_occ_planner_b200 as p
from tests import cases

n = int(os.environ.get("N_AGENTS", "64"))
cfg = cases.cfg_small_c2(p.POP_CELL_U8, n_local=n, n_total=n, nx=300, ny=300, nz=50, nt=20)
cloud, cyl = cases.world_moving(1236, (30.0, 30.0, 5.0), 0, 40)
start, goal = p.worldgen.circle_agents(n, 13.0)
eng = p.PopEngine(cfg, stream=torch.cuda.current_stream().cuda_stream)
eng.set_world_cloud(cloud); eng.set_cylinders(cyl); eng.set_poses(start[:, :3])
eng.profile_enable(True)
for _ in range(3):
    eng.map_update_gt(0.0)
eng.profile_read(True)
K = 10
for _ in range(K):
    eng.map_update_gt(0.0)
ms, ln = eng.profile_read(True)
bytes_per = cfg.nx * cfg.ny * cfg.nz * cfg.nt * n
print(os.environ.get("POP_CLEAR_MODE", "tma"), "voxel_update ms/step", ms["voxel_update"] / K,
      "GB/s", bytes_per / (ms["voxel_update"] / K * 1e-3) / 1e9, "cloud", len(cloud))
