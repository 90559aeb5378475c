"""pred-occ-planner_b200 — B200-native per-replan hot path of a spatio-temporal-occupancy swarm
planner: SOGM voxel update, hybrid-A* collision lookups, safe corridors, Bézier QP, peer
projection.  The compute lives in libpop.so (hand-written sm_100a CUDA behind include/pop.h);
this package is the thin host mirror of the reference's interfaces."""
from ._lib import (POP_CELL_F32, POP_CELL_U8, POP_KERNEL_CUBE, POP_KERNEL_DISC_Z0, POP_MAX_BODY,
                   POP_MAX_FACES, POP_MAX_PATH, POP_MAX_PIECES, POP_QUERY_FAKE, POP_QUERY_RISKBASE,
                   CYL_DTYPE, TRAJ_DTYPE, PopConfig, PopDspConfig, PopGridmapConfig, PopError, default_config, load)
from .engine import AgentMap, PopEngine
from . import dsp, gridmap, scenes, swarm, wire, worldgen

__all__ = ["PopEngine", "AgentMap", "PopConfig", "PopError", "default_config", "load", "dsp", "gridmap", "swarm", "wire", "worldgen"]
