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


def test_peer_resampling_equals_oracle_and_reference_noise():
    """swarm/replan_risk_rate > 0: the Gaussian copies of a peer's body particles (particles.cpp:365-408) deposited by
    k_peer_scatter, fed the noise sequence libstdc++ hands the reference for a given clock value.  The oracle this is
    compared with is pinned bit for bit against the compiled reference (tests/test_ref_map_pins.py::
    test_peer_resampling_live).  Support identical; weights to 1e-6 relative (the device rounds exp once from double,
    glibc's expf is correctly rounded in all but rare cases; deposits into one cell are summed by atomics)."""
    from tests import cases, oracle_lib
    orc = oracle_lib.load()
    cfg = cases.cfg_c1(p.POP_CELL_F32)
    cfg.n_agents_local, cfg.n_agents_total, cfg.replan_risk_rate, cfg.num_resample = 1, 2, 0.1, 10
    cells = exact = 0
    try:
        for seed in range(6):
            z = oracle_lib.libstdcxx_normal_triples(1000 + 77 * seed, 64 * cfg.num_resample)
            oracle_lib.set_resample_noise(z)
            case = rml.resample_case(seed)
            cfg.agent_id_base = case["ego"]
            eng = p.PopEngine(cfg)
            eng.set_resample_noise(z)
            eng.set_body_particles(case["offs"], case["counts"])
            eng.set_trajectories(case["recs"])
            eng.set_world_cloud(np.zeros((0, 3), np.float32))
            eng.set_cylinders(np.zeros(0, p.CYL_DTYPE))
            eng.set_poses(case["pose"].astype(np.float32)[None])
            eng.map_update_gt(case["t_map"])
            got = eng.agent_map(0).risk_maps()
            om = orc.map(cfg)
            om.set_pose(case["pose"].astype(np.float32))
            om.update_gt(np.zeros((0, 3), np.float32), np.zeros(0, p.CYL_DTYPE), case["t_map"])
            om.add_other_agents(case["offs"], case["counts"], case["recs"], case["ego"])
            want = om.risk_maps()
            assert np.array_equal(got != 0, want != 0), seed
            assert np.allclose(got, want, rtol=1e-6, atol=1e-7), (seed, float(np.abs(got - want).max()))
            cells += np.count_nonzero(want)
            exact += int(np.count_nonzero(got == want))
            eng.close()
    finally:
        oracle_lib.set_resample_noise(np.zeros((0, 3), np.float32))
    assert cells > 2000
