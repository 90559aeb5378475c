"""The CUDA map path against outputs of the REFERENCE's own map classes — no oracle in between.

tests/golden/ref_map_pins.npz holds what plan_env/src/fake_particle_risk_voxel.cpp, risk_base.cpp and map.cpp
(compiled unmodified into oracle/_ref/libref_map_*.so, see oracle/Makefile and scripts/make_golden_ref_map.py) produced
on the seeded scenes of tests/ref_map_lib.py: the SOGM after updateMap (static and moving cylinders, vertical rings, one
peer's projected trajectory) or after futureRiskCallback + addOtherAgents, getClearOcccupancy by slice index and by
time, getObstaclePoints on two boxes (emission order), getInflateOccupancy in its three forms.  Here the same scenes go
through libpop's C ABI (pop_map_update_gt / pop_map_future_risk, pop_map_query*, pop_map_obstacle_points,
pop_map_query_inflate) and every output must be identical — voxel indices, cell values, query answers, point order.
Both cell types where the scene allows (u8 cells hold the ground-truth map's 0/1 + unit peer risks exactly; float
risks need f32)."""
import os

import numpy as np
import pytest

import pred_occ_planner_b200 as p
from tests import ref_map_lib as rml

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_map_pins.npz")
PARAMS = [(n, c) for n in rml.MAP_CASES for c in (p.POP_CELL_F32, p.POP_CELL_U8) if not (n == "risk_peer" and c == p.POP_CELL_U8)]


@pytest.mark.parametrize("name,cell", PARAMS)
def test_map_path_equals_reference_outputs(name, cell):
    gold = np.load(GOLD)
    case = rml.map_case(name)
    cfg = case["cfg"]
    cfg.cell_dtype = cell
    cfg.n_agents_local = 1
    cfg.max_obstacle_points = 40000
    if case["swarm"] is not None:
        ego, offs, counts, recs = case["swarm"]
        cfg.n_agents_total, cfg.agent_id_base = 2, ego
    eng = p.PopEngine(cfg)
    if case["swarm"] is not None:
        eng.set_body_particles(offs, counts)
        eng.set_trajectories(recs)
    pose = case["pose"].astype(np.float32)[None]
    if case["kind"] == 0:
        eng.set_world_cloud(case["cloud"])
        eng.set_cylinders(case["cyl"])
        eng.set_poses(pose)
        eng.map_update_gt(case["t_map"])
    else:
        eng.map_future_risk(case["risk"][None], case["t_map"], pose)
    am = eng.agent_map(0)
    r = am.risk_maps().ravel()
    nz = np.flatnonzero(r).astype(np.uint32)
    assert np.array_equal(nz, gold[f"{name}.nz"]), "occupied (voxel, slice) set differs from the reference's"
    assert np.array_equal(r[nz], gold[f"{name}.val"]), "cell values differ from the reference's"
    assert np.array_equal(am.getClearOcccupancy(case["q_pos"], case["q_t"]), gold[f"{name}.q_t"])
    assert np.array_equal(am.getClearOcccupancy(case["q_pos"], case["q_dt"]), gold[f"{name}.q_dt"])
    for k, (t0, t1, lc, hc) in enumerate(case["boxes"]):
        pts = am.getObstaclePoints(t0, t1, lc, hc)
        assert np.array_equal(pts.astype(np.float32), gold[f"{name}.obs{k}"]) and np.array_equal(pts, pts.astype(np.float32))
    if case["kind"] == 0:
        assert np.array_equal(am.getInflateOccupancy(case["q_pos"]), gold[f"{name}.inf0"])
        assert np.array_equal(am.getInflateOccupancy(case["q_pos"], case["q_t"]), gold[f"{name}.inf_t"])
        assert np.array_equal(am.getInflateOccupancy(case["q_pos"], case["q_dt_in"]), gold[f"{name}.inf_dt"])
    eng.close()
