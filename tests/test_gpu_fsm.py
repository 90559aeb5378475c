"""The FSM bookkeeping that follows a replan (pop_fsm_update; FiniteStateMachineFake::FSMCallback,
plan_manager/src/fake_plan_manager.cpp:117-197, publishEmptyTrajectory :389-408): failure counter, the hover ("empty")
trajectory after too many failed replans or a failed fresh plan, goal detection — in a closed loop where one agent flies
freely, one gets walled in after its first plan and one starts next to its goal."""
import numpy as np
import pytest

import pred_occ_planner_b200 as p
from tests import cases

pytestmark = pytest.mark.gpu

NEW_PLAN, EXEC_TRAJ, REPLAN, GOAL_REACHED = 0, 1, 2, 3


def _expect(state, fails, ok, dist, max_failures, tol):
    """the reference's rule, restated on the host"""
    hover = False
    if state == GOAL_REACHED:
        return state, fails, False
    if state == NEW_PLAN:
        if ok:
            state = EXEC_TRAJ
        else:
            hover = True
    else:
        if ok:
            fails, state = 0, EXEC_TRAJ
        else:
            fails += 1
            state = REPLAN
            if fails > max_failures:
                state, hover = NEW_PLAN, True
    if state == EXEC_TRAJ and dist < tol:
        state = GOAL_REACHED
    return state, fails, hover


def test_failure_counter_hover_and_goal():
    n = 3
    cfg = cases.cfg_small_c2(p.POP_CELL_U8, n_local=n, n_total=n)
    eng = p.PopEngine(cfg)
    start = np.zeros((n, 9))
    start[:, :3] = [[-4.0, -3.0, 1.0], [-4.0, 0.0, 1.0], [2.0, 3.0, 1.0]]
    goal = np.array([[4.0, -3.0, 1.0], [4.0, 0.0, 1.0], [2.6, 3.0, 1.0]])   # agent 2 starts 0.6 m from its goal
    eng.set_body_particles(*p.worldgen.body_table(n, [0.4, 0.4, 0.45]))
    eng.set_trajectories(np.zeros(n, p.TRAJ_DTYPE))
    eng.set_cylinders(np.zeros(0, p.CYL_DTYPE))
    empty = np.zeros((0, 3), np.float32)
    state, fails = [NEW_PLAN] * n, [0] * n
    max_failures, tol = 2, 1.0
    hovered = 0
    prev_tr = None
    for k in range(8):
        t = 30.0 + 0.1 * k
        if k > 0:
            eng.swarm_advance(t)
        s = eng.swarm_states() if k > 0 else start
        if k >= 2:   # wall agent 1 in: a shell of points around wherever it is now
            c = s[1, :3]
            g = np.mgrid[-6:7, -6:7, -4:5].reshape(3, -1).T * 0.1
            shell = g[(np.abs(g).max(axis=1) >= 0.35) | (np.abs(g[:, 2]) >= 0.3)]
            cloud = (c + shell[np.linalg.norm(shell, axis=1) < 0.75]).astype(np.float32)
        else:
            cloud = empty
        eng.set_world_cloud(cloud)
        st, tr = eng.replan_cycle(s, goal, s[:, :3].astype(np.float32), t, t + 0.02)
        got_state, got_fails = eng.fsm_update(t, replan_max_failures=max_failures, goal_tolerance=tol)
        _, send = eng.cycle_results()
        for a in range(n):
            dist = float(np.linalg.norm(s[a, :3] - goal[a]))
            state[a], fails[a], hover = _expect(state[a], fails[a], st[a] == 0, dist, max_failures, tol)
            assert (got_state[a], got_fails[a]) == (state[a], fails[a]), (k, a, st.tolist(), got_state.tolist(), got_fails.tolist())
            if hover:
                hovered += 1
                r = send[a]
                assert r["n_pieces"] == 1 and r["valid"] == 1 and r["start_time"] == t and abs(r["duration"][0] - 1.5) < 1e-6
                assert np.array_equal(r["cpts"][:5], np.tile(s[a, :3], (5, 1))) and not r["cpts"][5:].any()
                if prev_tr is not None:
                    assert r["traj_id"] == prev_tr["traj_id"][a] + 1
        eng.traj_publish_local()
        prev_tr = send
    assert state[0] == EXEC_TRAJ and state[2] == GOAL_REACHED
    assert hovered >= 1 and max(fails) >= 1, (hovered, fails, state)
    # a hovering agent stays where it is
    before = eng.swarm_states()[1, :3].copy()
    eng.swarm_advance(31.2)
    after = eng.swarm_states()[1]
    if state[1] == NEW_PLAN:
        assert np.allclose(after[:3], before, atol=1e-12) and np.abs(after[3:]).max() < 1e-12   # derivatives of a constant curve: rounding only
    eng.close()
