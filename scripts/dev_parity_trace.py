"""Development aid: where does a config's worst-agent control-point difference between the CUDA cycle and the oracle
come from?  Flies the bench's closed loop on the GPU to the bench's parity step, runs the oracle on the same inputs,
then repeats the step stage by stage on both sides for the worst agents (A* path, corridor faces, QP iterations).
  POP_CONFIG=c3 POP_STEP=15 python scripts/dev_parity_trace.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pred_occ_planner_b200 as p
import bench
from tests import oracle_lib

conf = os.environ.get("POP_CONFIG", "c3")
step = int(os.environ.get("POP_STEP", "15"))
n = bench.CONFIGS[conf]["agents"]
sc = bench.Scenario(conf, n, n, 0)
cfg, goal = sc.cfg, sc.goal_all
orc = oracle_lib.load()
eng = p.PopEngine(cfg)
eng.set_body_particles(*sc.body)
eng.set_trajectories(np.zeros(n, p.TRAJ_DTYPE))
start = sc.start_all.copy()
for s in range(step):
    if s > 0:
        eng.swarm_advance(sc.t(s)); start = eng.swarm_states()
    cloud, cyl = sc.world(s)
    eng.set_world_cloud(cloud); eng.set_cylinders(cyl)
    eng.replan_cycle(start, goal, start[:, :3].astype(np.float32), sc.t(s), sc.t(s) + bench.LOOKAHEAD)
    eng.traj_publish_local()
s = step
eng.swarm_advance(sc.t(s)); start = eng.swarm_states()
poses = start[:, :3].astype(np.float32)
peers = eng.get_trajectories()
cloud, cyl = sc.world(s)
eng.set_world_cloud(cloud); eng.set_cylinders(cyl)
st_gpu, tr_gpu = eng.replan_cycle(start, goal, poses, sc.t(s), sc.t(s) + bench.LOOKAHEAD)
own = np.zeros(n, p.TRAJ_DTYPE)
sec, st_cpu = bench.cpu_cycle(sc, cloud, cyl, start.copy(), goal, peers, n, os.cpu_count() or 1, sc.t(s), records=own)
rel = np.zeros(n)
for a in np.flatnonzero((st_gpu == 0) & (st_cpu == 0)):
    M = int(own["n_pieces"][a])
    x, y = tr_gpu["cpts"][a][:5 * M], own["cpts"][a][:5 * M]
    rel[a] = np.abs(x - y).max() / max(1.0, np.abs(y).max()) if M == int(tr_gpu["n_pieces"][a]) else np.inf
print("status equal", int((st_gpu == st_cpu).sum()), "of", n, "max rel", rel.max(), "agents above 1e-4:", np.flatnonzero(rel > 1e-4).tolist())
worst = np.argsort(rel)[-3:][::-1]
# the same step stage by stage (the map of the cycle above is still current)
t_rel = np.full(n, bench.LOOKAHEAD)
ret, path, ln, stats = eng.astar_search_batch(start, goal, t_rel)
tst = np.full(n, sc.t(s) + bench.LOOKAHEAD)
hp, rows, npoly, lgoal = eng.corridor_batch(path, ln, start, tst)
fin = np.zeros((n, 9)); fin[:, :6] = lgoal
dur = np.zeros((n, p.POP_MAX_PIECES))
for a in range(n):
    dur[a, :npoly[a]] = cfg.corridor_tau
cpts, qst, qit = eng.qp_solve_batch(npoly, start, fin, dur, hp, rows, cfg.opt_max_vel, cfg.opt_max_acc)
for a in worst:
    a = int(a)
    c = cfg.copy(); c.agent_id_base = cfg.agent_id_base + a
    m = orc.map(c)
    m.set_pose(poses[a]); m.update_gt(cloud, cyl, sc.t(s))
    m.add_other_agents(sc.body[0], sc.body[1], peers, a)
    oret, opath, ostats = m.astar(start[a], goal[a], bench.LOOKAHEAD, cap=p.POP_MAX_PATH)
    same_path = len(opath) == ln[a] and np.array_equal(opath, path[a, :ln[a]])
    on, ohp, orows, ogoal = m.corridor(opath, start[a], tst[a])
    M = int(on)
    face = max((np.abs(ohp[i, :orows[i]] - hp[a, i, :orows[i]]).max() for i in range(M) if orows[i] == rows[a, i]), default=float("nan"))
    ofin = np.zeros(9); ofin[:6] = ogoal
    d = np.full(M, cfg.corridor_tau)
    st_o, cp_o, it_o = orc.qp_solve(c, start[a], ofin, d, hp=ohp, rows=orows, vmax=cfg.opt_max_vel, amax=cfg.opt_max_acc)
    st_x, cp_x, it_x = orc.qp_solve(c, start[a], fin[a], d, hp=hp[a], rows=rows[a], vmax=cfg.opt_max_vel, amax=cfg.opt_max_acc)
    g = cpts[a].reshape(-1, 3)[:5 * M]
    print(f"agent {a}: cycle rel {rel[a]:.3e} | A* ret {oret}/{ret[a]} path identical {same_path} | corridors {M}/{npoly[a]} rows equal "
          f"{np.array_equal(orows[:M], rows[a, :M])} max face diff {face:.3e} goal diff {np.abs(ogoal - lgoal[a]).max():.1e} | "
          f"QP iterations: oracle on its faces {it_o}, GPU on its faces {int(qit[a])}, oracle on the GPU's faces {it_x} | "
          f"cpts: GPU vs oracle-on-GPU-faces {np.abs(g - cp_x).max():.2e}, GPU vs oracle {np.abs(g - cp_o).max():.2e}")
