"""Multi-rank plumbing of the swarm: one process per GPU, agents block-partitioned over ranks,
ONE allgather of fixed-size trajectory records per replan cycle.

The reference exchanges trajectories by ROS broadcast: every planner publishes its freshly
optimised Bezier trajectory on "/broadcast_traj" (plan_manager/src/fake_plan_manager.cpp:364-399)
and every peer stores the latest one it has heard (traj_coordinator/src/particles.cpp:131-191);
a map update then projects whatever arrived before it (fake_particle_risk_voxel.cpp:178-218).
Here the same one-cycle-delayed semantics come from an allgather of `pop_traj_record`s
(include/pop.h) placed between cycle k's QP and cycle k+1's map update.  Body particles
("/broadcast_particles", particles.cpp:25,62-87) are static and are gathered once at start-up.

On GPUs the allgather is issued by libpop itself (pop_comm_init / pop_exchange_trajs, csrc/pop_comm.cu: NCCL over
NVLink on the handle's stream); torch.distributed only carries the 128-byte communicator id at start-up, and on CPU
(gloo) the host-array path the tests exercise.
"""
from __future__ import annotations

import numpy as np

from ._lib import POP_MAX_BODY, TRAJ_DTYPE

RECORD_BYTES = TRAJ_DTYPE.itemsize


def shard(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Block partition of agent ids over ranks: (agent_id_base, n_agents_local).
    The allgather needs equal shares, so n_total must be a multiple of world."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    if n_total <= 0 or n_total % world:
        raise ValueError(f"{n_total} agents do not split evenly over {world} ranks")
    n_local = n_total // world
    return rank * n_local, n_local


def owner(agent_id: int, n_total: int, world: int) -> int:
    """Rank that plans for a global agent id."""
    return agent_id // (n_total // world)


class TrajExchange:
    """The per-cycle "/broadcast_traj" step.

    device mode (engine given): the exchange lives behind the C ABI — `pop_comm_init` gives the engine's handle an
    NCCL communicator (the 128-byte id is created on rank 0 by `pop_comm_unique_id` and broadcast here over the
    default torch.distributed group, whatever its backend) and `step()` is `pop_exchange_trajs`: one ncclAllGather
    on the engine's stream between libpop's own buffers; world == 1 is the device-to-device publish.
    host mode (engine None): numpy records in, numpy records out over the default process group
    (gloo) — the path the CPU tests exercise.
    """

    def __init__(self, n_total: int, world: int, rank: int, engine=None, device: int | None = None, unique_id: bytes | None = None,
                 pipelined: bool = False):
        self.world, self.rank, self.n_total = world, rank, n_total
        self.pipelined = pipelined   # device mode: pop_exchange_trajs_pipelined (remote records one cycle in flight)
        self.id_base, self.n_local = shard(n_total, world, rank)
        self.engine = engine
        self.bytes_per_cycle = self.n_local * RECORD_BYTES
        if engine is not None:
            if device is not None and device != engine.device:
                raise ValueError(f"engine lives on cuda:{engine.device}, not cuda:{device}")
            if world > 1 and unique_id is None:
                unique_id = self._broadcast_id(engine, rank)
            engine.comm_init(unique_id, rank, world)

    @staticmethod
    def _broadcast_id(engine, rank: int) -> bytes:
        import torch
        import torch.distributed as dist
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = torch.frombuffer(bytearray(engine.comm_unique_id()), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            buf = buf.to(f"cuda:{engine.device}")
        dist.broadcast(buf, src=0)
        return bytes(buf.cpu().numpy().tobytes())

    def step(self, local_records: np.ndarray | None = None) -> np.ndarray | None:
        if self.engine is not None:
            self.engine.exchange_trajs(self.pipelined)
            return None
        import torch
        import torch.distributed as dist
        local_records = np.ascontiguousarray(local_records, dtype=TRAJ_DTYPE)
        if len(local_records) != self.n_local:
            raise ValueError(f"expected {self.n_local} local records, got {len(local_records)}")
        out = np.zeros(self.n_total, TRAJ_DTYPE)
        if self.world == 1:
            out[:] = local_records
            return out
        send = torch.from_numpy(local_records.view(np.uint8).reshape(-1))
        recv = torch.from_numpy(out.view(np.uint8).reshape(-1))
        dist.all_gather_into_tensor(recv, send)
        return out


def gather_body_particles(offs_local: np.ndarray, counts_local: np.ndarray, world: int):
    """One-off exchange of every agent's body-particle offsets (host arrays, default group)."""
    offs_local = np.ascontiguousarray(offs_local, np.float64).reshape(-1, POP_MAX_BODY, 3)
    counts_local = np.ascontiguousarray(counts_local, np.int32)
    if world == 1:
        return offs_local.copy(), counts_local.copy()
    import torch
    import torch.distributed as dist
    n_local = len(counts_local)
    offs = np.zeros((n_local * world, POP_MAX_BODY, 3))
    counts = np.zeros(n_local * world, np.int32)
    dist.all_gather_into_tensor(torch.from_numpy(offs.reshape(-1)), torch.from_numpy(offs_local.reshape(-1)))
    dist.all_gather_into_tensor(torch.from_numpy(counts), torch.from_numpy(counts_local))
    return offs, counts
