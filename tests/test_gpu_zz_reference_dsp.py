"""The CUDA particle map against the state of the REFERENCE's own node — no oracle in between.

tests/golden/ref_dsp_pins.npz holds, for every frame of the seeded 10-frame flight of tests/ref_dsp_lib.py, what
plan_env/src/risk_mapping_node.cpp + plan_env/include/plan_env/dsp_dynamic.h (compiled unmodified into
oracle/_ref/libref_dsp_ref.so, see oracle/Makefile and scripts/make_golden_ref_dsp.py) held after cloudCallback:
observation bins with C_k, every particle, the pyramid tables, per-point cluster velocities, voxel occupancy and mean
velocity (SHA-256 digests of the arrays: these are bit-exact quantities), the noise-table sequence counters, and the
published future-risk grid of three frames.  Here the same clouds, odometry and noise tables go through
pop_dsp_update and must reproduce every digest; the future risk agrees to 1e-5 relative (its fp32 deposits are summed
by atomics on the device, so the order differs from the sequential sweep) with identical support."""
import os

import numpy as np
import pytest

from tests import cases
from tests import ref_dsp_lib as rdl

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_dsp_pins.npz")
EXACT = ("n_obs", "n_live", "n_pyr", "n_pv", "obs", "flags", "particles", "pyramids", "velocity", "voxels", "occupancy_sum")


def test_particle_map_flight_equals_reference_state(pop):
    from pred_occ_planner_b200 import dsp, worldgen
    gold = np.load(GOLD)
    cfg = cases.dsp_cfg_ref()
    d = cases.dsp_params(cfg)
    eng = pop.PopEngine(cfg, device=0)
    maps = dsp.DspMaps(eng, d)
    maps.set_random_tables(*worldgen.dsp_random_tables(d.n_randoms, 5))
    view = rdl._Agent(maps, 0)
    V, T = cfg.nx * cfg.ny * cfg.nz, cfg.nt
    for k, (cam, pos, quat, st) in enumerate(rdl.flight_frames()):
        assert maps.cloudCallback([cam], pos[None], quat[None], st).tolist() == [int(gold[f"f{k}.ok"])]
        got = rdl.frame_state(view)
        for name in EXACT:
            assert np.array_equal(got[name], gold[f"f{k}.{name}"]), (k, name, got["n_live"], gold[f"f{k}.n_live"])
        assert np.array_equal(maps.stats(0)[4:8], gold[f"f{k}.stats"]), k
        if k in rdl.RISK_FRAMES:
            want = np.zeros(V * T, np.float32)
            want[gold[f"f{k}.risk_nz"]] = gold[f"f{k}.risk_val"]
            risk = eng.agent_map(0).risk_maps().ravel()
            assert np.array_equal(risk != 0, want != 0), k
            assert np.allclose(risk, want, rtol=1e-5, atol=1e-6), (k, float(np.abs(risk - want).max()))
    eng.close()
