"""Stage-by-stage probe of BASELINE config 3 (development aid): which stage is slow / stuck."""
import sys, time, faulthandler
sys.path.insert(0, '.')
import numpy as np
import pred_occ_planner_b200 as p
from tests import cases
n = int(sys.argv[1])
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 500
nz = int(sys.argv[3]) if len(sys.argv) > 3 else 60
t=time.time()
def log(*a): print(f"[{time.time()-t:7.2f}s]", *a, flush=True)
cfg = cases.cfg_small_c2(p.POP_CELL_U8, n_local=n, n_total=n, nx=nx, ny=nx, nz=nz, nt=20)
cloud, cyl = cases.world_moving(1237, (50.0, 50.0, 6.0), 60, 60)
start, goal = p.worldgen.lattice_agents(n, arena=(50.0, 50.0), seed=3)
poses = start[:, :3].astype(np.float32)
eng = p.PopEngine(cfg)
eng.set_world_cloud(cloud); eng.set_cylinders(cyl)
eng.set_body_particles(*p.worldgen.body_table(n, [0.4, 0.4, 0.45]))
eng.set_trajectories(np.zeros(n, p.TRAJ_DTYPE))
eng.set_poses(poses); eng.map_update_gt(5.0); eng.sync(); log("map update")
faulthandler.dump_traceback_later(45, exit=True)
ret, path, ln, stats = eng.astar_search_batch(start, goal, np.full(n, 0.02)); log("astar", np.bincount(ret, minlength=6).tolist(), "iters max", stats[:, 0].max(), "len", ln.tolist()[:16])
faulthandler.cancel_dump_traceback_later(); faulthandler.dump_traceback_later(45, exit=True)
hp, rows, npoly, g = eng.corridor_batch(path, ln, start, np.full(n, 5.02)); log("corridor", npoly.tolist()[:32], "rows max", rows.max())
faulthandler.cancel_dump_traceback_later(); faulthandler.dump_traceback_later(45, exit=True)
status, traj = eng.replan_cycle(start, goal, poses, 5.0, 5.02); log("cycle", np.bincount(status, minlength=5).tolist())
