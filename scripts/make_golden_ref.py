"""Regenerates tests/golden/ref_pins.npz: outputs of the REFERENCE's own code (oracle/_ref/libref.so =
plan_env/src/raycast.cpp, traj_utils/sdlp.hpp, sfc_gen/lbfgs.hpp, sfc_gen/firi.hpp, traj_utils/src/bernstein.cpp, traj_opt/src/bezier_optimizer.cpp, path_searching/src/{fake_,}risk_hybrid_a_star.cpp, traj_coordinator/src/particles.cpp, plan_manager/src/baseline{_fake,}.cpp compiled unmodified from
/root/reference; see oracle/Makefile) on the seeded cases of tests/ref_lib.py.  tests/test_ref_pins.py checks the
oracle against these vectors wherever the library itself is not available.  Needs /root/reference.

    python scripts/make_golden_ref.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_RAYS, N_LP, N_COST, N_LBFGS, N_MVIE, N_FIRI, N_BEZ, N_QP, N_ASTAR, N_REPLAN = 300, 200, 60, 40, 40, 45, 40, 12, 40, 30


def main():
    from tests import ref_lib as rl
    ref = rl.load()
    assert ref is not None, "oracle/_ref/libref.so cannot be built here (no /root/reference)"
    out = {}
    paths, oks = [], []
    for s, e in rl.ray_cases(N_RAYS):
        p, ok = ref.raycaster(s, e)
        paths.append(p.astype(np.int16))
        oks.append(ok)
    out["ray_path"], out["ray_offs"] = rl.pack(paths)
    out["ray_ok"] = np.array(oks, np.int8)
    vals, xs = [], []
    for c, A, b in rl.lp_cases(N_LP):
        v, x = ref.linprog(c, A, b)
        vals.append(v)
        xs.append(np.pad(x, (0, 4 - len(x))))
    out["lp_val"], out["lp_x"] = np.array(vals), np.array(xs)
    cg = []
    for A, x in rl.mvie_cost_cases(N_COST):
        c, g = ref.cost_mvie(A, x)
        cg.append(np.r_[c, g])
    out["cost"] = np.array(cg)
    lb = []
    for A, x0 in rl.lbfgs_cases(N_LBFGS):
        ret, x, f = ref.lbfgs_mvie(A, x0)
        lb.append(np.r_[ret, f, x])
    out["lbfgs"] = np.array(lb)
    mv = []
    for h in rl.poly_cases(N_MVIE):
        ok, R, p, r = ref.mvie(h, np.eye(3), np.zeros(3), np.full(3, 0.3))
        vol, Q, viol = rl.ellipsoid(R, p, r, h)
        mv.append(np.r_[ok, vol, viol, p, Q.ravel()])
    out["mvie"] = np.array(mv)
    for it in (1, 2, 4):
        hs = [ref.firi(bd, pc, a, b, it, r0)[0] for bd, pc, a, b, r0, lo, hi in rl.firi_cases(N_FIRI)]
        out[f"firi{it}"], out[f"firi{it}_offs"] = rl.pack(hs)
    bz, rates = [], []
    for dur, cp, t in rl.bezier_cases(N_BEZ):
        bz.append(np.stack([ref.bezier_eval(dur, cp, t, d) for d in (0, 1, 2)]))
        rates.append([ref.bezier_max_rate(dur, cp, 1), ref.bezier_max_rate(dur, cp, 2)])
    out["bezier"], out["bezier_offs"] = rl.pack(bz)
    out["bezier_rates"] = np.array(rates)
    qps = []
    for init, goal, dur, polys, vmax, amax in rl.qp_cases(N_QP, max_pieces=4):
        qps.append(np.concatenate([a.ravel() for a in ref.qp_setup(init, goal, dur, polys, vmax, amax)]))
    out["qp"], out["qp_offs"] = rl.pack(qps)
    from tests import oracle_lib
    orc = oracle_lib.load()
    for key, (cfg, (cloud, cyl)), risk, seed in (("astar", rl.astar_world(), False, 8), ("astar_risk", rl.astar_world_risk(), True, 9)):
        paths, heads = [], []
        for pose, start, goal, t0 in rl.astar_cases(N_ASTAR, seed=seed):
            m = orc.map(cfg)
            m.set_pose(pose)
            m.update_gt(cloud, cyl, 0.0)
            ret, pth, st = ref.astar(cfg, m, pose, start, goal, t0, risk=risk)
            paths.append(pth)
            heads.append([ret, st[0], st[1], st[2]])
        out[key + "_path"], out[key + "_offs"] = rl.pack(paths)
        out[key + "_head"] = np.array(heads, np.int32)
    from tests import oracle_cycle
    for key, (cfg, (cloud, cyl)), risk, seed in (("replan", rl.astar_world(), False, 67), ("replan_risk", rl.astar_world_risk(), True, 79)):
        heads, cps = [], []
        for pose, start, goal, ts in rl.astar_cases(N_REPLAN, seed=seed):
            m = orc.map(cfg)
            m.set_pose(pose)
            m.update_gt(cloud, cyl, 0.0)
            m.set_time(100.0)
            # without a single valid corridor both planners index an empty list (`hPolys.size() - 1` underflows,
            # baseline_fake.cpp:366 / baseline.cpp:363): undefined behaviour, not run
            if oracle_cycle.replan(m, None, 0, start, goal, 100.0 + ts)[0] == 2:
                heads.append([-1, 0, 0, 0, 0])
                cps.append(np.zeros((0, 3)))
                continue
            ok, M, cp, st = ref.replan(cfg, m, pose, 100.0, 100.0 + ts, start, goal, risk=risk)
            heads.append([ok, M, st[0], st[1], st[3]])
            cps.append(cp)
        out[key + "_head"] = np.array(heads, np.int32)
        out[key + "_cpts"], out[key + "_offs"] = rl.pack(cps)
    # one batch with a common plan time: what pop_replan_cycle can be compared with directly (tests/test_gpu_zz_...)
    cfg, (cloud, cyl) = rl.astar_world()
    heads, cps = [], []
    for pose, start, goal, _ in rl.astar_cases(24, seed=83):
        m = orc.map(cfg)
        m.set_pose(pose)
        m.update_gt(cloud, cyl, 100.0)
        if oracle_cycle.replan(m, None, 0, start, goal, 100.02)[0] == 2:
            heads.append([-1, 0, 0, 0, 0])
            cps.append(np.zeros((0, 3)))
            continue
        ok, M, cp, st = ref.replan(cfg, m, pose, 100.0, 100.02, start, goal)
        heads.append([ok, M, st[0], st[1], st[3]])
        cps.append(cp)
    out["batch_head"] = np.array(heads, np.int32)
    out["batch_cpts"], out["batch_offs"] = rl.pack(cps)
    cfg, size, ego, offs, counts, recs, qa, qt = rl.peer_case()
    out["peer_ego"] = ref.ego_particles(size)
    pts, rk, on = ref.particles_with_risk(cfg.n_agents_total, ego, size, offs, counts, recs, 10.0, qa, qt)
    out["peer_pts"], out["peer_risks"], out["peer_n"] = pts, rk, on
    cpts, dur, M, nows = rl.sep_case()
    sets = [rl.ref_sep_sets(ref, cfg.n_agents_total, ego, size, offs, counts, recs, now, cpts, dur, M) for now in nows]
    out["sep_n"] = np.array([s[0] for s in sets], np.int32)
    out["sep_sets"] = np.stack([s[1] for s in sets])
    out["traj_safe"] = np.concatenate([ref.traj_safe(cfg, m, pose, t_map, rec, cfg.corridor_tau, now, T)
                                       for cfg, m, pose, t_map, rec, now, T in rl.traj_safe_cases(orc)])
    path = os.path.join(ROOT, "tests", "golden", "ref_pins.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
