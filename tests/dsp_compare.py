"""Stage-by-stage comparison of the CUDA particle map with the sequential oracle (used by the GPU parity
tests and by scripts/dev_dsp_parity.py).  Everything is bit-exact except the future-risk accumulators, whose
fp32 deposits are summed by atomics on the device (order differs from the sequential sweep)."""
import numpy as np

RISK_RTOL = 1e-5   # relative, on the fp32 future-risk sums
RISK_ATOL = 1e-6


def compare_frame(gpu, agent, orc, risk_gpu, thr):
    """Returns a list of human-readable mismatches (empty = parity)."""
    bad = []
    pc_g, cnt_g, ml_g = gpu.observations(agent)
    pc_o, cnt_o, ml_o = orc.observations()
    if not np.array_equal(cnt_g, cnt_o):
        bad.append(f"observation counts differ in {np.count_nonzero(cnt_g != cnt_o)} pyramids")
    else:
        m = np.arange(pc_o.shape[1])[None, :] < cnt_o[:, None]
        if not np.array_equal(pc_g[m][:, :3], pc_o[m][:, :3]):
            bad.append("observation points differ")
        if not np.array_equal(pc_g[m][:, 3], pc_o[m][:, 3]):
            d = np.abs(pc_g[m][:, 3] - pc_o[m][:, 3])
            bad.append(f"C_k differs in {np.count_nonzero(d)} of {d.size} observations, max {d.max():.3e}")
    if not np.array_equal(ml_g, ml_o):
        bad.append("pyramid max lengths differ")
    pv_g, pv_o = gpu.points_with_velocity(agent), orc.points_with_velocity()
    if pv_g.shape != pv_o.shape or not np.array_equal(pv_g, pv_o):
        n = min(len(pv_g), len(pv_o))
        first = int(np.argmax((pv_g[:n] != pv_o[:n]).any(axis=1))) if n else -1
        bad.append(f"points with velocity differ: {pv_g.shape} vs {pv_o.shape}, first row {first}: "
                   f"{pv_g[first] if n else None} vs {pv_o[first] if n else None}")
    py_g, py_o = gpu.pyramids(agent), orc.pyramids()
    py_o = py_o.copy()
    py_o[py_o[:, :, 0] == 0] = 0          # stale (voxel, slot) of cleared entries are not kept on the device
    if not np.array_equal(py_g, py_o):
        bad.append(f"pyramid tables differ in {np.count_nonzero((py_g != py_o).any(axis=(1, 2)))} pyramids")
    st_g, st_o = gpu.stats(agent)[:7], orc.stats()[:7]
    if not np.array_equal(st_g, st_o):
        bad.append(f"stats differ: gpu {st_g.tolist()} oracle {st_o.tolist()}")
    pa_g, pa_o = gpu.particles(agent), orc.particles()
    live_g, live_o = pa_g[:, :, 0] > 0.1, pa_o[:, :, 0] > 0.1
    if not np.array_equal(pa_g[:, :, 0], pa_o[:, :, 0]):
        bad.append(f"particle flags differ in {np.count_nonzero(pa_g[:, :, 0] != pa_o[:, :, 0])} slots "
                   f"(live {live_g.sum()} vs {live_o.sum()})")
    elif not np.array_equal(pa_g[live_o], pa_o[live_o]):
        d = (pa_g[live_o] != pa_o[live_o])
        bad.append(f"particle states differ in {np.count_nonzero(d.any(axis=1))} of {live_o.sum()} particles; columns "
                   f"{np.nonzero(d.any(axis=0))[0].tolist()}")
    vo_g, vo_o = gpu.voxels(agent), orc.voxels()
    if not np.array_equal(vo_g, vo_o):
        bad.append(f"voxel occupancy/velocity differ in {np.count_nonzero((vo_g != vo_o).any(axis=1))} voxels")
    risk_o, _ = orc.export()
    if not np.allclose(risk_gpu, risk_o, rtol=RISK_RTOL, atol=RISK_ATOL):
        d = np.abs(risk_gpu - risk_o)
        bad.append(f"future risk differs: max abs {d.max():.3e}, cells {(d > RISK_ATOL + RISK_RTOL * np.abs(risk_o)).sum()}")
    else:
        clear = np.abs(risk_o - thr) > 1e-4
        if not np.array_equal((risk_gpu > thr)[clear], (risk_o > thr)[clear]):
            bad.append("occupied (voxel, slice) sets differ away from the threshold")
        if not np.array_equal(risk_gpu != 0, risk_o != 0):
            bad.append("support of the future risk differs")
    return bad
