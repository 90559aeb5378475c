"""Host-side multi-rank logic on CPU: world_size-2 gloo process group (SURVEY 8e — agents
block-partitioned over ranks, ONE allgather of trajectory records per replan cycle).

The two ranks run the CPU oracle's replan for their own agents (the oracle is the checker here:
this is not the product path), exchange the records with `swarm.TrajExchange`, and every rank
must end up with exactly the table a single process computes for the whole swarm — including the
one-cycle delay of peer trajectories.
"""
import os
import socket
import sys

import numpy as np
import pytest

import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import swarm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_partition():
    assert swarm.shard(1024, 8, 0) == (0, 128)
    assert swarm.shard(1024, 8, 7) == (896, 128)
    assert [swarm.owner(a, 64, 4) for a in (0, 15, 16, 63)] == [0, 0, 1, 3]
    covered = []
    for r in range(4):
        b, n = swarm.shard(64, 4, r)
        covered += list(range(b, b + n))
    assert covered == list(range(64))
    with pytest.raises(ValueError):
        swarm.shard(10, 4, 0)
    with pytest.raises(ValueError):
        swarm.shard(8, 4, 4)
    assert swarm.RECORD_BYTES == 1536


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _swarm_inputs(n_total):
    from tests import cases
    cfg = cases.cfg_c1(p.POP_CELL_F32, n_local=n_total, n_total=n_total)
    cloud, cyl = cases.world_c1(seed=5)
    rng = np.random.default_rng(11)
    start = np.zeros((n_total, 9))
    ang = np.linspace(0, 2 * np.pi, n_total, endpoint=False)
    start[:, 0], start[:, 1], start[:, 2] = 3.5 * np.cos(ang), 3.5 * np.sin(ang), 1.0 + 0.1 * rng.random(n_total)
    goal = -start[:, :3].copy()
    goal[:, 2] = 1.0
    body = p.worldgen.body_table(n_total, [0.3, 0.3, 0.4])
    return cfg, cloud, cyl, start, goal, body


def _plan_cycle(cfg, cloud, cyl, body, recs, start, goal, ids, t_map, t_plan):
    """Oracle cycle (map update incl. peers + replan) for the given global agent ids."""
    from tests import oracle_cycle, oracle_lib
    orc = oracle_lib.load()
    out = np.zeros(len(ids), p.TRAJ_DTYPE)
    status = np.zeros(len(ids), np.int32)
    for k, a in enumerate(ids):
        om = orc.map(cfg)
        pose = start[a, :3].astype(np.float32)
        om.set_pose(pose)
        om.update_gt(cloud, cyl, t_map)
        om.add_other_agents(body[0], body[1], recs, a)
        st, rec, _ = oracle_cycle.replan(om, (body[0], body[1], recs), a, start[a], goal[a], t_plan)
        status[k] = st
        if st == 0:
            out[k] = rec
        else:   # nothing published: peers keep the previous trajectory
            out[k] = recs[a]
    return status, out


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        cfg, cloud, cyl, start, goal, body = _swarm_inputs(n_total)
        id_base, n_local = swarm.shard(n_total, world, rank)
        # body particles: each rank contributes its own agents' rows once at start-up
        offs, counts = swarm.gather_body_particles(body[0][id_base:id_base + n_local],
                                                   body[1][id_base:id_base + n_local], world)
        assert np.array_equal(offs, body[0]) and np.array_equal(counts, body[1])
        x = swarm.TrajExchange(n_total, world, rank)
        table = np.zeros(n_total, p.TRAJ_DTYPE)
        ids = list(range(id_base, id_base + n_local))
        tables = []
        for c in range(2):
            st, mine = _plan_cycle(cfg, cloud, cyl, (offs, counts), table, start, goal, ids, 1.0 + 0.1 * c, 1.02 + 0.1 * c)
            table = x.step(mine)
            tables.append(table.copy())
        q.put((rank, tables))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_matches_single_process():
    import torch.multiprocessing as mp
    n_total, world = 4, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    # single-process reference: all agents planned against the same (delayed) table
    cfg, cloud, cyl, start, goal, body = _swarm_inputs(n_total)
    table = np.zeros(n_total, p.TRAJ_DTYPE)
    for c in range(2):
        st, table = _plan_cycle(cfg, cloud, cyl, body, table, start, goal, list(range(n_total)), 1.0 + 0.1 * c, 1.02 + 0.1 * c)
        for r in range(world):
            assert got[r][c].tobytes() == table.tobytes(), f"rank {r} cycle {c}: gathered table differs"
    assert (table["n_pieces"] > 0).any(), "no agent produced a trajectory: the test exercised nothing"
    ids = table["drone_id"][table["n_pieces"] > 0]
    assert np.array_equal(ids, np.nonzero(table["n_pieces"] > 0)[0])


def test_single_rank_exchange_is_identity():
    x = swarm.TrajExchange(3, 1, 0)
    recs = np.zeros(3, p.TRAJ_DTYPE)
    recs["drone_id"] = [0, 1, 2]
    recs["n_pieces"] = 2
    out = x.step(recs)
    assert out.tobytes() == recs.tobytes()
    with pytest.raises(ValueError):
        x.step(recs[:2])
