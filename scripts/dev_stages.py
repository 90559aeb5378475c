"""Dev helper: run the c2 batch stage by stage through the C ABI and print per-stage statistics."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pred_occ_planner_b200 as p
import bench

conf = os.environ.get("POP_CONFIG", "c2")
n = int(os.environ.get("N_AGENTS", str(bench.CONFIGS[conf]["agents"])))
sc = bench.Scenario(conf, n, n, 0)
cfg, start, goal, body = sc.cfg, sc.start_all, sc.goal_all, sc.body
cloud, cyl = sc.world(0)
eng = p.PopEngine(cfg)
eng.set_world_cloud(cloud); eng.set_cylinders(cyl); eng.set_body_particles(*body)
eng.set_trajectories(np.zeros(n, p.TRAJ_DTYPE))
poses = start[:, :3].astype(np.float32)
st, tr = eng.replan_cycle(start, goal, poses, 100.0, 100.02)
eng.traj_publish_local()
# a few closed-loop steps so that the batch below is a mid-flight one
for k in range(1, int(os.environ.get("POP_STEPS", "6"))):
    eng.swarm_advance(100.0 + 0.1 * k)
    cloud, cyl = sc.world(k)
    eng.set_world_cloud(cloud); eng.set_cylinders(cyl)
    eng.replan_cycle_resident(100.0 + 0.1 * k, 100.02 + 0.1 * k)
    eng.traj_publish_local()
k += 1
eng.swarm_advance(100.0 + 0.1 * k)
start = eng.swarm_states()
poses = start[:, :3].astype(np.float32)
cloud, cyl = sc.world(k)
eng.set_world_cloud(cloud); eng.set_cylinders(cyl)
eng.set_poses(poses)
eng.map_update_gt(100.0 + 0.1 * k)
eng.profile_enable(True)
ret, path, ln, stats = eng.astar_search_batch(start, goal, np.full(n, 0.02))
ms, _ = eng.profile_read(True)
print("astar ms", ms["astar"], "iters sorted tail", np.sort(stats[:, 0])[-12:], "nodes tail", np.sort(stats[:, 1])[-8:], "retried", int(stats[:, 2].sum()))
print("astar ret", np.bincount(ret, minlength=6), "iters", stats[:, 0].min(), np.median(stats[:, 0]), stats[:, 0].max(),
      "nodes max", stats[:, 1].max(), "len", np.bincount(ln.clip(0)))
hp, rows, npoly, lgoal = eng.corridor_batch(path, ln, start, np.full(n, 100.02 + 0.1 * k))
ms, _ = eng.profile_read(True)
print("corridor ms", ms["corridor"])
print("npoly", np.bincount(npoly), "rows mean", rows[rows > 0].mean(), "max", rows.max())
fin = np.zeros((n, 9)); fin[:, :6] = lgoal
dur = np.zeros((n, p.POP_MAX_PIECES))
for a in range(n): dur[a, :npoly[a]] = cfg.corridor_tau
cpts, status, iters = eng.qp_solve_batch(npoly, start, fin, dur, hp, rows, cfg.opt_max_vel, cfg.opt_max_acc)
ms, _ = eng.profile_read(True)
print("qp status", {int(s): int((status == s).sum()) for s in np.unique(status)}, "iters", np.sort(iters), "ms", ms["qp"])
prof = eng.qp_profile(n)
order = np.argsort(iters)
print("qp cycles per agent: iters | setup | factor total | loop | nfact | loop/iter | gather/it kkt/it rows/it")
for a in order[-10:]:
    it = max(1, int(iters[a]))
    print(int(iters[a]), int(prof[a, 0]), int(prof[a, 1]), int(prof[a, 2]), int(prof[a, 3]), int(prof[a, 2] / it),
          int(prof[a, 4] / it), int(prof[a, 5] / it), int(prof[a, 6] / it), '| events total', int(prof[a, 7]), 'per event', int(prof[a, 7] / max(1, it // 25)), 'update_info', int(prof[a, 0] / max(1, it // 25)), 'check', int(prof[a, 1] / max(1, it // 25)))
print("median setup", np.median(prof[:, 0]), "median cycles per factorisation", np.median(prof[:, 1] / np.maximum(prof[:, 3], 1)))
