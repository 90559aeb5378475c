"""BASELINE config 5 — fine-grid stress of the voxel-update kernel: ONE agent map of
2000x2000x160 voxels @0.05 m with 40 time slices (25.6 GB of u8 cells), cylinder count swept.
Reports the voxel-update stage time (CUDA events on the engine's stream, inputs resident) and the
achieved fraction of the measured HBM roofline for the tiled path and the clear+scatter path.

    python scripts/voxel_sweep.py [--cyl 40 400] [--steps 5]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pred_occ_planner_b200 as p  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cyl", type=int, nargs="+", default=[40, 400])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--grid", type=int, nargs=4, default=[2000, 2000, 160, 40])
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    nx, ny, nz, nt = args.grid
    for ncyl in args.cyl:
        c = p.default_config()
        c.nx, c.ny, c.nz, c.nt = nx, ny, nz, nt
        c.resolution, c.time_resolution = 0.05, 0.05
        c.cell_dtype = p.POP_CELL_U8
        c.n_agents_local = c.n_agents_total = 1
        c.max_cylinders = max(256, ncyl + 1)
        arena = (nx * 0.05, ny * 0.05, nz * 0.05)
        cloud, cyl = p.worldgen.make_world(1239, arena=arena, n_static=ncyl // 2, n_moving=ncyl - ncyl // 2)
        c.max_cloud_points = len(cloud) + 1024
        poses = np.array([[0.0, 0.0, arena[2] / 2 - 0.1]], np.float32)
        for mode in ("tiles", "scatter"):
            os.environ["POP_VOXEL_MODE"] = mode
            eng = p.PopEngine(c)
            eng.set_world_cloud(cloud)
            eng.set_cylinders(cyl)
            eng.set_poses(poses)
            for _ in range(2):
                eng.map_update_gt(0.0)
            eng.profile_enable(True)
            eng.profile_read(True)
            for i in range(args.steps):
                eng.map_update_gt(0.05 * i)
            ms, _ = eng.profile_read(True)
            t = ms["voxel_update"] / args.steps
            V = nx * ny * nz
            alg = V * nt + len(cloud) * 12 + len(cloud) * (nt - 1)
            print(json.dumps({"config": f"c5 {nx}x{ny}x{nz} T={nt} u8, {ncyl} cylinders, {len(cloud)} cloud points",
                              "path": mode, "voxel_update_ms": round(t, 3), "algorithmic_GB": round(alg / 1e9, 3),
                              "achieved_GBps": round(alg / t / 1e6, 1), "frac_of_measured_hbm": round(alg / t / 1e6 / peak, 3)}),
                  flush=True)
            eng.close()


if __name__ == "__main__":
    main()
