"""Regenerates tests/golden/oracle_dsp.npz: outputs of the CPU oracle for the depth sensor (K0) and the
particle map (K1d) on a fixed 6-frame flight — what `tests/test_gpu_golden.py` compares libpop against
without the oracle at run time, and what `tests/test_oracle_dsp.py` uses to notice accidental changes of the
oracle itself.

    python scripts/make_golden_dsp.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def scenario():
    """Inputs shared by the generator and the tests (all derived from fixed seeds)."""
    import pred_occ_planner_b200 as p
    from pred_occ_planner_b200 import worldgen
    from tests import cases
    cfg = cases.dsp_cfg_ref()
    d = cases.dsp_params(cfg)
    cam = p._lib.PopCamera()
    p.load().pop_camera_default(p._lib.C.byref(cam), 160, 120)
    cloud, cyl = worldgen.make_world(11, arena=(10.0, 10.0, 3.0), n_static=8, n_moving=0, height=3.0)
    gx, gy = np.meshgrid(np.arange(-5, 5, 0.1), np.arange(-5, 5, 0.1), indexing="ij")
    ground = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1).astype(np.float32)
    cloud = np.concatenate([cloud, ground]).astype(np.float32)
    tabs = worldgen.dsp_random_tables(d.n_randoms, 5)
    frames = []
    for k in range(6):
        pos = np.array([-4.0 + 0.1 * k, -0.5 + 0.05 * k, 1.0])
        yaw = 0.15 * np.sin(0.5 * k)
        quat = np.array([np.cos(yaw / 2), 0.0, 0.0, np.sin(yaw / 2)])
        frames.append((pos, quat, 70.0 + 0.1 * k))
    return cfg, d, cam, cloud, tabs, frames


def main():
    from tests import oracle_lib
    cfg, d, cam, cloud, tabs, frames = scenario()
    orc = oracle_lib.load()
    m = orc.dsp(cfg, d)
    m.set_random_tables(*tabs)
    out = {}
    for k, (pos, quat, st) in enumerate(frames):
        depth, pc, _ = oracle_lib.render(cam, pos, quat, cloud)
        assert m.update(pc, pos.astype(np.float32), quat.astype(np.float32), st) == 1
        risk, nocc = m.export()
        if k in (0, 5):
            out[f"depth_{k}"] = depth.astype(np.int32)
            out[f"ncloud_{k}"] = np.int32(len(pc))
            out[f"filtered_{k}"] = m.filter(pc)
        out[f"stats_{k}"] = m.stats()
        out[f"nocc_{k}"] = np.int32(nocc)
    parts = m.particles()
    live = np.argwhere(parts[:, :, 0] > 0.1).astype(np.int32)
    out["live_slots"] = live
    out["live_state"] = parts[live[:, 0], live[:, 1]]
    vox = m.voxels()
    occ = np.nonzero(vox[:, 0] > 0)[0].astype(np.int32)
    out["vox_idx"] = occ
    out["vox_val"] = vox[occ]
    nz = np.argwhere(risk != 0).astype(np.int32)
    out["risk_idx"] = nz
    out["risk_val"] = risk[nz[:, 0], nz[:, 1]]
    path = os.path.join(ROOT, "tests", "golden", "oracle_dsp.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(live), "live particles,", len(nz), "non-zero risk cells")


if __name__ == "__main__":
    main()
