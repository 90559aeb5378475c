"""The trajectory exchange behind the C ABI (csrc/pop_comm.cu): pop_comm_init + pop_exchange_trajs.

Single GPU: a world of one needs no NCCL id and the exchange is the device-to-device publish.  Two GPUs (skipped on
smaller boxes): two processes, one handle each, the NCCL id handed over through a file; after every cycle both ranks
hold the same table, and entry i of it is exactly the record agent i's owner planned."""
import os
import subprocess
import sys

import numpy as np
import pytest

import pred_occ_planner_b200 as p
from tests import cases
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def test_single_rank_exchange_is_the_local_publish():
    n = 4
    cfg = cases.cfg_small_c2(p.POP_CELL_U8, n_local=n, n_total=n)
    eng = p.PopEngine(cfg)
    eng.comm_init(None, 0, 1)
    assert eng.comm_info()[:2] == (0, 1)
    start, goal = p.worldgen.circle_agents(n, 4.0)
    eng.set_world_cloud(np.zeros((0, 3), np.float32))
    eng.set_cylinders(np.zeros(0, p.CYL_DTYPE))
    eng.set_body_particles(*p.worldgen.body_table(n, [0.4, 0.4, 0.45]))
    eng.set_trajectories(np.zeros(n, p.TRAJ_DTYPE))
    st, tr = eng.replan_cycle(start, goal, start[:, :3].astype(np.float32), 10.0, 10.02)
    eng.exchange_trajs()
    eng.sync()
    tab = eng.get_trajectories()
    assert (st == 0).all() and np.array_equal(tab.view(np.uint8), tr.view(np.uint8))
    # a handle that holds only part of the swarm must be given a communicator first
    cfg2 = cases.cfg_small_c2(p.POP_CELL_U8, n_local=2, n_total=4)
    e2 = p.PopEngine(cfg2)
    with pytest.raises(p.PopError) as err:
        e2.exchange_trajs()
    assert err.value.code == -5
    with pytest.raises(p.PopError):
        e2.comm_init(None, 1, 2)       # rank 1 of 2 does not own ids 0..1
    e2.close()
    eng.close()


def test_two_rank_allgather(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "exchange_worker.py"), str(r), "2", str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [pr.communicate(timeout=600)[0] for pr in procs]
    assert all(pr.returncode == 0 for pr in procs), "\n".join(outs)[-4000:]
    r0, r1 = (np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(2))
    rec = p.TRAJ_DTYPE.itemsize
    for k in ("table0", "table1"):
        assert np.array_equal(r0[k], r1[k])                      # every rank holds the same table
    for k in (0, 1):
        tab = r0[f"table{k}"]
        assert np.array_equal(tab[:3 * rec], r0[f"own{k}"]) and np.array_equal(tab[3 * rec:], r1[f"own{k}"])
    t1 = r0["table1"].view(p.TRAJ_DTYPE)
    assert (t1["n_pieces"] > 0).all() and sorted(t1["drone_id"].tolist()) == list(range(6))
    # pipelined variant (pop_exchange_trajs_pipelined): a rank's own block is the cycle it just planned, the other
    # rank's block is what that rank planned one call earlier (before the first pipelined call: the last strict table)
    ranks = (r0, r1)
    for me in (0, 1):
        other = 1 - me
        mine = slice(me * 3 * rec, (me + 1) * 3 * rec)
        theirs = slice(other * 3 * rec, (other + 1) * 3 * rec)
        for k in range(3):
            tab = ranks[me][f"ptable{k}"]
            assert np.array_equal(tab[mine], ranks[me][f"pown{k}"])
            expect = ranks[other][f"pown{k - 1}"] if k > 0 else ranks[other]["own1"]
            assert np.array_equal(tab[theirs], expect), (me, k)
