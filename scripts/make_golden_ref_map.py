"""Regenerates tests/golden/ref_map_pins.npz: outputs of the REFERENCE's own map classes (oracle/_ref/libref_map_*.so =
plan_env/src/map.cpp + fake_particle_risk_voxel.cpp | risk_base.cpp + traj_coordinator/src/particles.cpp compiled
unmodified from /root/reference; oracle/Makefile) on the seeded scenes of tests/ref_map_lib.py: the SOGM after
updateMap / futureRiskCallback (+ one peer's projected trajectory), getClearOcccupancy by slice and by time,
getObstaclePoints on two boxes, getInflateOccupancy.  tests/test_ref_map_pins.py checks the oracle against these
vectors, tests/test_gpu_zz_reference_maps.py the CUDA path.  Needs /root/reference.

    python scripts/make_golden_ref_map.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from tests import ref_map_lib as rml
    out = {}
    for name in rml.MAP_CASES:
        case = rml.map_case(name)
        got = rml.run_reference(case)
        assert got is not None, f"oracle/_ref/libref_map_{case['variant']}.so missing (no /root/reference?)"
        for k, v in got.items():
            out[f"{name}.{k}"] = v
        print(name, "occupied", len(got["nz"]), "obstacle points", len(got["obs0"]), len(got["obs1"]))
    path = os.path.join(ROOT, "tests", "golden", "ref_map_pins.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
