"""GPU parity of the corridor stage (K3) against the oracle on identical A* paths: corridor
count, face count per corridor, faces within 1e-6 (L-BFGS runs on CUDA vs libm `log`), local goal."""
import numpy as np
import pytest

import pred_occ_planner_b200 as p
from tests import cases

pytestmark = pytest.mark.gpu


def _case(oracle, cfg, cloud, cyl, start, goal, t_map, t_rel):
    n = cfg.n_agents_local
    poses = start[:, :3].astype(np.float32)
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    eng.set_poses(poses)
    eng.map_update_gt(t_map)
    ret, path, ln, stats = eng.astar_search_batch(start, goal, np.full(n, t_rel))
    tst = np.full(n, t_map + t_rel)
    hp, rows, npoly, lgoal = eng.corridor_batch(path, ln, start, tst)
    bad = []
    n_checked = 0
    for a in range(n):
        if ret[a] == 0:
            assert npoly[a] == 0 or ln[a] >= 2
            continue
        c = cfg.copy()
        c.agent_id_base = cfg.agent_id_base + a      # LP shuffle seed = global agent id
        m = oracle.map(c)
        m.set_pose(poses[a])
        m.update_gt(cloud, cyl, t_map)
        on, ohp, orows, ogoal = m.corridor(path[a, :ln[a]], start[a], tst[a])
        ok = on == npoly[a]
        if ok and on > 0:
            ok = np.array_equal(orows[:on], rows[a, :on]) and np.allclose(ogoal, lgoal[a], rtol=0, atol=1e-9)
            if ok:
                for i in range(on):
                    ok = ok and np.allclose(ohp[i, :orows[i]], hp[a, i, :orows[i]], rtol=1e-6, atol=1e-6)
        n_checked += 1
        if not ok:
            bad.append((a, int(on), int(npoly[a]), orows[:on].tolist(), rows[a, :max(on, npoly[a])].tolist()))
    return bad, n_checked, npoly


def test_c1_corridors(oracle):
    n = 24
    cfg = cases.cfg_c1(p.POP_CELL_U8, n_local=n, n_total=n)
    cloud, cyl = cases.world_c1(seed=99)
    rng = np.random.default_rng(8)
    start = np.zeros((n, 9))
    start[:, :2] = rng.uniform(-4.5, 4.5, (n, 2))
    start[:, 2] = rng.uniform(0.7, 1.5, n)
    start[:, 3:5] = rng.uniform(-0.8, 0.8, (n, 2))
    goal = np.concatenate([rng.uniform(-4.5, 4.5, (n, 2)), rng.uniform(0.8, 1.5, (n, 1))], axis=1)
    bad, checked, npoly = _case(oracle, cfg, cloud, cyl, start, goal, 2.0, 0.02)
    assert not bad, bad
    assert checked >= n // 2 and (npoly > 0).sum() >= n // 3


def test_moving_world_corridors(oracle):
    n = 16
    cfg = cases.cfg_small_c2(p.POP_CELL_U8, n_local=n, n_total=n, nx=160, ny=160, nz=30, nt=20)
    cloud, cyl = cases.world_moving(21, (16.0, 16.0, 3.0), 6, 18)
    start, goal = p.worldgen.circle_agents(n, 6.0)
    bad, checked, npoly = _case(oracle, cfg, cloud, cyl, start, goal, 5.0, 0.02)
    assert not bad, bad
    assert (npoly > 0).sum() >= n // 2
