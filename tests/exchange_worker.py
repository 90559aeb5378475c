"""Worker of tests/test_gpu_exchange.py: one rank of a two-GPU swarm that plans its block of agents and exchanges the
trajectory records through libpop's own communicator (pop_comm_init / pop_exchange_trajs).  The 128-byte NCCL id
travels through a file — no torch.distributed, as a C++ host would do it with MPI or a socket.
usage: exchange_worker.py RANK WORLD DIR"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pred_occ_planner_b200 as p   # noqa: E402
from tests import cases             # noqa: E402


def main():
    rank, world, d = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    n_local = 3
    n_total = n_local * world
    cfg = cases.cfg_small_c2(p.POP_CELL_U8, n_local=n_local, n_total=n_total)
    cfg.agent_id_base = rank * n_local
    eng = p.PopEngine(cfg, device=rank)
    id_path = os.path.join(d, "nccl_id.bin")
    if rank == 0:
        uid = eng.comm_unique_id()
        with open(id_path + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(id_path + ".tmp", id_path)
    else:
        for _ in range(600):
            if os.path.exists(id_path):
                break
            time.sleep(0.05)
        uid = open(id_path, "rb").read()
    eng.comm_init(uid if world > 1 else None, rank, world)
    assert eng.comm_info()[:2] == (rank, world)
    start, goal = p.worldgen.circle_agents(n_total, 4.0)
    sl = slice(rank * n_local, (rank + 1) * n_local)
    eng.set_world_cloud(np.zeros((0, 3), np.float32))
    eng.set_cylinders(np.zeros(0, p.CYL_DTYPE))
    eng.set_body_particles(*p.worldgen.body_table(n_total, [0.4, 0.4, 0.45]))
    eng.set_trajectories(np.zeros(n_total, p.TRAJ_DTYPE))
    tables = []
    own = []
    for k in range(2):   # two cycles: the second one plans against the peers' first trajectories
        st, tr = eng.replan_cycle(start[sl], goal[sl], start[sl, :3].astype(np.float32), 10.0 + 0.1 * k, 10.02 + 0.1 * k)
        eng.exchange_trajs()
        eng.sync()
        tables.append(eng.get_trajectories())
        own.append(tr)
        assert (st == 0).all(), st
    # three more cycles through the pipelined variant: own block fresh, the other rank's block one call behind
    ptables, pown = [], []
    for k in range(2, 5):
        st, tr = eng.replan_cycle(start[sl], goal[sl], start[sl, :3].astype(np.float32), 10.0 + 0.1 * k, 10.02 + 0.1 * k)
        eng.exchange_trajs(pipelined=True)
        eng.sync()
        ptables.append(eng.get_trajectories())
        pown.append(tr)
    np.savez(os.path.join(d, f"rank{rank}.npz"), table0=tables[0].view(np.uint8), table1=tables[1].view(np.uint8),
             own0=own[0].view(np.uint8), own1=own[1].view(np.uint8),
             **{f"ptable{k}": t.view(np.uint8) for k, t in enumerate(ptables)},
             **{f"pown{k}": t.view(np.uint8) for k, t in enumerate(pown)})
    eng.close()


if __name__ == "__main__":
    main()
