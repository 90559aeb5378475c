"""Regenerates tests/golden/ref_dsp_pins.npz: the state of the REFERENCE's own particle-map node (oracle/_ref/
libref_dsp_ref.so = plan_env/src/risk_mapping_node.cpp + plan_env/include/plan_env/dsp_dynamic.h compiled unmodified
from /root/reference; oracle/Makefile, oracle/ref_glue_dsp.cpp) after each frame of the seeded 10-frame flight of
tests/ref_dsp_lib.py: observation bins with C_k, every particle, pyramid tables, per-point cluster velocities, voxel
occupancy (as SHA-256 digests — these are compared bit for bit), table sequence counters, and the published future
risk of three frames (sparse).  Needs /root/reference.

    python scripts/make_golden_ref_dsp.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from pred_occ_planner_b200 import worldgen
    from tests import cases, ref_dsp_lib as rdl
    dcfg = cases.dsp_cfg_ref()
    dpar = cases.dsp_params(dcfg)
    rd = rdl.fresh(sigma_observation=dpar.sigma_observation, filter_res=dpar.voxel_filter_resolution,
                   risk_threshold=dpar.occupancy_threshold)
    assert rd is not None, "oracle/_ref/libref_dsp_ref.so missing (no /root/reference?)"
    rd.set_random_tables(*worldgen.dsp_random_tables(dpar.n_randoms, 5))
    out = {}
    for k, (cam, pos, quat, st) in enumerate(rdl.flight_frames()):
        out[f"f{k}.ok"] = np.int32(rd.update(cam, pos, quat, st))
        for name, v in rdl.frame_state(rd).items():
            out[f"f{k}.{name}"] = v
        out[f"f{k}.stats"] = rd.stats()
        risk, _ = rd.export()
        if k in rdl.RISK_FRAMES:
            out[f"f{k}.risk_nz"], out[f"f{k}.risk_val"] = rdl.sparse(risk)
        print(k, "live", int(out[f"f{k}.n_live"]), "obs", int(out[f"f{k}.n_obs"]), "stats", out[f"f{k}.stats"].tolist(),
              "risk cells", int(np.count_nonzero(risk)))
    path = os.path.join(ROOT, "tests", "golden", "ref_dsp_pins.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
