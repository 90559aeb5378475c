"""The CUDA path against outputs of the REFERENCE's own code — no oracle in between.

First the hybrid A*, then the whole replan cycle (below).

tests/golden/ref_pins.npz holds what path_searching/src/fake_risk_hybrid_a_star.cpp (compiled unmodified into
oracle/_ref/libref.so, see oracle/Makefile and scripts/make_golden_ref.py) returned for 40 seeded problems on a grid
with 12 pillars and 8 moving cylinders: ASTAR_RET, expansions, nodes used, whether the init=false retry ran, and the
path sampled by getPathWithVel(corridor_tau).  Here the same 40 problems are one batch of pop_astar_search_batch
(one agent per problem, each with its own pose-centred grid and its own start time) and must give the same answers:
integers identical, path states within 1e-9 (identical unless CUDA and libm round a transcendental differently
— the tolerance test_gpu_astar.py uses).  The grids themselves are the CUDA voxel update's, not the oracle's.
(The file sorts after the other GPU tests on purpose: they compare with the oracle, which tests/test_ref_pins.py
ties to the same vectors on the CPU.)"""
import os

import numpy as np
import pytest

import pred_occ_planner_b200 as p
from tests import cases, ref_lib as rl

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_pins.npz")


@pytest.mark.parametrize("cell", [p.POP_CELL_F32, p.POP_CELL_U8])
def test_astar_batch_equals_reference_outputs(cell):
    gold = np.load(GOLD)
    heads = gold["astar_head"]
    want = rl.unpack(gold["astar_path"], gold["astar_offs"], 6)
    n = len(heads)
    probs = rl.astar_cases(n)
    _, (cloud, cyl) = rl.astar_world()
    cfg = cases.cfg_c1(cell, n_local=n, n_total=n)
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    eng.set_poses(np.stack([pr[0] for pr in probs]).astype(np.float32))
    eng.map_update_gt(0.0)
    start = np.stack([pr[1] for pr in probs])
    goal = np.stack([pr[2] for pr in probs])
    t0 = np.array([pr[3] for pr in probs])
    ret, path, ln, stats = eng.astar_search_batch(start, goal, t0)
    eng.close()
    bad = []
    for a in range(n):
        ok = ([int(ret[a]), int(stats[a, 0]), int(stats[a, 1]), int(stats[a, 2])] == heads[a].tolist()
              and int(ln[a]) == len(want[a])
              and np.allclose(path[a, :ln[a]], want[a], rtol=0, atol=1e-9))
        if not ok:
            bad.append((a, int(ret[a]), stats[a].tolist(), heads[a].tolist(), int(ln[a]), len(want[a])))
    assert not bad, bad
    assert (ret != 0).sum() >= n // 2


def test_replan_cycle_against_the_reference_planner():
    """pop_replan_cycle (CUDA voxel update, A*, corridors, QP, record) for 24 agents at one plan time, against what the
    reference's own FakeBaselinePlanner::replan returned for each of them (compiled unmodified into oracle/_ref/libref.so;
    recorded in tests/golden/ref_pins.npz by scripts/make_golden_ref.py, where the map and the OSQP solve behind the
    reference's wrapper are the oracle's).  The same replans succeed, with the same number of pieces; control points
    within 1e-2 relative — the solver runs at eps 1e-3 and the corridor faces of two correct runs differ by ~1e-4
    (FIRI's L-BFGS stopping rule; LP shuffles: mt19937 there, an LCG here).  One borderline status flip is tolerated."""
    gold = np.load(GOLD)
    heads = gold["batch_head"]
    wants = rl.unpack(gold["batch_cpts"], gold["batch_offs"], 3)
    n = len(heads)
    probs = rl.astar_cases(n, seed=83)
    _, (cloud, cyl) = rl.astar_world()
    cfg = cases.cfg_c1(p.POP_CELL_U8, n_local=n, n_total=n)
    start = np.stack([pr[1] for pr in probs])
    goal = np.stack([pr[2] for pr in probs])
    poses = np.stack([pr[0] for pr in probs]).astype(np.float32)
    eng = p.PopEngine(cfg)
    eng.set_world_cloud(cloud)
    eng.set_cylinders(cyl)
    eng.set_body_particles(*p.worldgen.body_table(n, list(cfg.drone_size)))
    eng.set_trajectories(np.zeros(n, p.TRAJ_DTYPE))        # nobody has published: no peers, as in the recorded runs
    status, traj = eng.replan_cycle(start, goal, poses, 100.0, 100.02)
    eng.close()
    flips, rel, n_ok = [], [], 0
    for a in range(n):
        if heads[a, 0] < 0:                                # no valid corridor: the reference was not run (undefined there)
            assert status[a] != 0
            continue
        if (status[a] == 0) != bool(heads[a, 0]):
            flips.append((a, int(status[a]), heads[a].tolist()))
            continue
        if status[a] == 0:
            M = int(heads[a, 1])
            assert int(traj["n_pieces"][a]) == M, (a, int(traj["n_pieces"][a]), M)
            got = np.asarray(traj["cpts"][a]).reshape(-1, 3)[:5 * M]
            rel.append(np.abs(got - wants[a]).max() / max(1.0, np.abs(wants[a]).max()))
            n_ok += 1
    assert len(flips) <= 1, flips
    assert n_ok >= n // 3 and max(rel) <= 1e-2 and np.median(rel) <= 1e-3, (n_ok, max(rel), float(np.median(rel)))
