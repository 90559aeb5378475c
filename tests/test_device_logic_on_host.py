"""Device logic checked on the CPU: the product's own `__device__` source compiled for the host.

pred_occ_planner_b200/csrc/pop_lp.cuh (Seidel LP) and pop_corridor_finish.cuh (per-agent tail of the corridor stage:
intersection cut, minimum corridor count, goal search — for both planners of the reference, see the header) contain
no thread indices or intrinsics, so tests/csrc/corridor_finish_host.cpp builds the SAME source with g++ and this file
compares it with the oracle on seeded cases — bit for bit, since both sides draw the LP shuffles from the same LCG.
pop_bezier.cuh uses CUDA's round-to-nearest intrinsics only; tests/csrc/bezier_host.cpp maps them to the plain
operators and checks the evaluation against the reference's own outputs; tests/csrc/geom_host.cpp does the same
for pop_geom.cuh (voxel index math and the occupancy query that k_query runs and k_astar inlines).
It is how kernel-side logic gets exercised in `-m "not gpu"` runs; the GPU parity tests still go through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import POP_MAX_FACES, POP_MAX_PIECES
from tests import ref_lib as rl

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "corridor_finish_host.cpp")
OUT = os.path.join(HERE, "_build", "libdevhost.so")
DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int32)


CSRC = os.path.join(os.path.dirname(HERE), "pred_occ_planner_b200", "csrc")


def _build(src, out, headers):
    deps = [src] + [os.path.join(CSRC, h) for h in headers]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wl,--exclude-libs,ALL",
                        "-I" + cuda_inc, "-o", out, src], check=True)
    return C.CDLL(out)


@pytest.fixture(scope="module")
def dev():
    lib = _build(SRC, OUT, ["pop_lp.cuh", "pop_corridor_finish.cuh"])
    lib.dev_linprog.restype = C.c_double
    return lib


@pytest.fixture(scope="module")
def dev_bezier():
    return _build(os.path.join(HERE, "csrc", "bezier_host.cpp"), os.path.join(HERE, "_build", "libdevbezier.so"),
                  ["pop_bezier.cuh"])


def test_device_bezier_source_against_reference_vectors(dev_bezier, oracle):
    """pop_bezier.cuh (what the peer projection, the trajectory server and the separation check evaluate trajectories
    with) compiled for the host, against the outputs of the reference's own bernstein.cpp (tests/golden/ref_pins.npz)
    and against the oracle: 1e-14 relative, almost always identical (the device source raises s to integer powers
    with an error-free product instead of pow())."""
    gold = np.load(os.path.join(HERE, "golden", "ref_pins.npz"))
    wants = rl.unpack(gold["bezier"], gold["bezier_offs"])
    same = total = 0
    for (dur, cp, t), want in zip(rl.bezier_cases(len(wants)), wants):
        want = want.reshape(3, len(t), 3)
        for d in (0, 1, 2):
            out = np.zeros((len(t), 3))
            dev_bezier.dev_bezier_eval(dur.ctypes.data_as(DP), len(dur), cp.ctypes.data_as(DP), t.ctypes.data_as(DP),
                                       len(t), d, out.ctypes.data_as(DP))
            assert np.abs(out - want[d]).max() <= 1e-14 * max(1.0, np.abs(want[d]).max())
            ora = oracle.bezier_eval(dur, cp, t, deriv=d)
            assert np.abs(out - ora).max() <= 1e-14 * max(1.0, np.abs(ora).max())
            same += int(np.array_equal(out, want[d]))
            total += 1
    assert same >= 0.9 * total


def test_device_lp_equals_oracle_lp(dev, oracle):
    """poplp::linprog<d> (device source) and the oracle's linprog share algorithm and shuffle: identical results."""
    n_fin = 0
    for i, (c, A, b) in enumerate(rl.lp_cases(400, seed=71)):
        x = np.zeros(len(c))
        v = dev.dev_linprog(len(c), c.ctypes.data_as(DP), A.ctypes.data_as(DP), b.ctypes.data_as(DP), len(b),
                            x.ctypes.data_as(DP), C.c_uint64(i + 1))
        vo, xo = oracle.linprog(c, A, b, seed=i + 1)
        assert (v == vo or (np.isnan(v) and np.isnan(vo))) and np.array_equal(x, xo), (i, v, vo)
        n_fin += int(np.isfinite(v))
    assert n_fin > 200


def _box(lo, hi):
    h = np.zeros((6, 4))
    for k in range(3):
        h[2 * k, k], h[2 * k, 3] = 1.0, -hi[k]
        h[2 * k + 1, k], h[2 * k + 1, 3] = -1.0, lo[k]
    return h


def _chain(rng):
    """A path of n samples and one corridor per segment: boxes around the segment with random cuts; now and then a
    corridor is invalid, shifted away from its neighbour, or too short to contain the path's end."""
    n = int(rng.integers(2, 10))
    pts = np.cumsum(rng.uniform(0.2, 0.9, (n, 3)) * [1, rng.choice([-1, 1]), 0.1], axis=0)
    path = np.zeros((n, 6))
    path[:, :3] = pts
    path[:, 3:] = rng.normal(size=(n, 3))
    hp = np.zeros((POP_MAX_PIECES, POP_MAX_FACES, 4))
    rows = np.zeros(POP_MAX_PIECES, np.int32)
    valid = np.zeros(POP_MAX_PIECES, np.int32)
    for i in range(n - 1):
        lo = np.minimum(pts[i], pts[i + 1]) - rng.uniform(0.15, 0.6, 3)
        hi = np.maximum(pts[i], pts[i + 1]) + rng.uniform(0.15, 0.6, 3)
        u = rng.random()
        if u < 0.12:                      # does not reach its neighbour
            hi = lo + 0.3 * (hi - lo)
        elif u < 0.2:                     # far away: no intersection with the previous one
            lo, hi = lo + 5.0, hi + 5.0
        h = [_box(lo, hi)]
        for _ in range(int(rng.integers(0, 8))):
            nv = rng.normal(size=3)
            nv /= np.linalg.norm(nv)
            mid = 0.5 * (pts[i] + pts[i + 1])
            h.append(np.array([[*nv, -(nv @ mid) - rng.uniform(0.2, 0.8)]]))
        h = np.concatenate(h)
        hp[i, :len(h)] = h
        rows[i] = len(h)
        valid[i] = 0 if rng.random() < 0.06 else 1
    start = np.zeros(9)
    start[:3] = pts[0]
    return n, path, hp, rows, valid, start


@pytest.mark.parametrize("variant", [p.POP_QUERY_FAKE, p.POP_QUERY_RISKBASE])
def test_corridor_tail_device_source_equals_oracle(dev, oracle, variant):
    rng = np.random.default_rng(5 + variant)
    kept, cut, failed, moved = 0, 0, 0, 0
    for k in range(1500):
        n, path, hp, rows, valid, start = _chain(rng)
        g1, g2 = np.full(6, np.nan), np.full(6, np.nan)
        args = (variant, C.c_uint64(k), n, path.ctypes.data_as(DP), hp.ctypes.data_as(DP), rows.ctypes.data_as(IP),
                valid.ctypes.data_as(IP), start.ctypes.data_as(DP))
        a = dev.dev_corridor_finish(*args, g1.ctypes.data_as(DP))
        b = oracle.lib.orc_corridor_tail(*args, g2.ctypes.data_as(DP))
        assert a == b, (k, a, b)
        if a:
            assert np.array_equal(g1, g2), (k, g1, g2)
            n_valid = int(np.argmin(np.r_[valid[:n - 1], 0]))
            kept += int(a == n_valid)
            cut += int(a < n_valid)
            moved += int(not np.array_equal(g1[:3], path[a - 1, :3]))
        else:
            failed += 1
    # every branch is visited: nothing cut, cut by an intersection or the goal search, failed, goal left moved
    assert kept > 100 and cut > 100 and failed > 100 and moved >= 3, (kept, cut, failed, moved)


def test_the_two_planners_differ_where_the_reference_does(dev):
    """One corridor left: the simulator planner goes on (baseline_fake.cpp:384), the real planner gives up
    (baseline.cpp:381).  Corridors 2 and 3 apart: the first keeps 0..2 (:374), the second 0..1 (:371)."""
    pts = np.array([[0, 0, 1.0], [1, 0, 1], [2, 0, 1], [3, 0, 1], [4, 0, 1]])
    path = np.zeros((5, 6))
    path[:, :3] = pts
    hp = np.zeros((POP_MAX_PIECES, POP_MAX_FACES, 4))
    rows = np.zeros(POP_MAX_PIECES, np.int32)
    valid = np.zeros(POP_MAX_PIECES, np.int32)
    for i in range(4):
        lo, hi = pts[i] - 0.5, pts[i + 1] + 0.5
        if i == 3:
            lo, hi = lo + [0, 5, 0], hi + [0, 5, 0]
        hp[i, :6], rows[i], valid[i] = _box(lo, hi), 6, 1
    start = np.zeros(9)
    start[:3] = pts[0]

    def run(variant, n):
        g = np.zeros(6)
        return dev.dev_corridor_finish(variant, C.c_uint64(1), n, path.ctypes.data_as(DP), hp.ctypes.data_as(DP),
                                       rows.ctypes.data_as(IP), valid.ctypes.data_as(IP), start.ctypes.data_as(DP),
                                       g.ctypes.data_as(DP)), g
    assert run(p.POP_QUERY_FAKE, 2)[0] == 1 and run(p.POP_QUERY_RISKBASE, 2)[0] == 0
    assert run(p.POP_QUERY_FAKE, 5)[0] == 3 and run(p.POP_QUERY_RISKBASE, 5)[0] == 2
    assert np.array_equal(run(p.POP_QUERY_FAKE, 5)[1][:3], pts[2]) and np.array_equal(run(p.POP_QUERY_RISKBASE, 5)[1][:3], pts[1])


@pytest.fixture(scope="module")
def dev_geom():
    return _build(os.path.join(HERE, "csrc", "geom_host.cpp"), os.path.join(HERE, "_build", "libdevgeom.so"),
                  ["pop_geom.cuh", "pop_internal.h"])


@pytest.mark.parametrize("variant,kernel,clearance,cell", [
    (p.POP_QUERY_FAKE, p.POP_KERNEL_DISC_Z0, 0.45, p.POP_CELL_F32),      # simulator map: 5 x 5 x 1 footprint
    (p.POP_QUERY_FAKE, p.POP_KERNEL_DISC_Z0, 0.45, p.POP_CELL_U8),
    (p.POP_QUERY_RISKBASE, p.POP_KERNEL_CUBE, 0.15, p.POP_CELL_F32),     # RiskBase: 3 x 3 x 3 cube, decaying threshold
    (p.POP_QUERY_RISKBASE, p.POP_KERNEL_CUBE, 0.3, p.POP_CELL_F32),      # 5 x 5 x 5
    (p.POP_QUERY_FAKE, p.POP_KERNEL_DISC_Z0, 0.6, p.POP_CELL_U8),        # a footprint without a specialised plane sum
])
def test_device_query_source_equals_oracle(dev_geom, oracle, variant, kernel, clearance, cell):
    """geom_query (what k_query runs and k_astar inlines) on the host, on a grid built by the oracle from a world with
    moving cylinders: every answer equal — inside, on the borders, outside the box, below ground, above the ceiling,
    at every slice."""
    cfg, (cloud, cyl) = rl.astar_world()
    cfg.query_variant, cfg.kernel_shape, cfg.clearance, cfg.cell_dtype = variant, kernel, clearance, cell
    if variant == p.POP_QUERY_RISKBASE:
        cfg.risk_threshold_region, cfg.risk_region_decay = 0.9, 0.05
    rng = np.random.default_rng(11)
    for k in range(3):
        pose = np.array([rng.uniform(-3, 3), rng.uniform(-3, 3), 1.0 + 0.1 * k], np.float32)
        m = oracle.map(cfg)
        m.set_pose(pose)
        m.update_gt(cloud, cyl, 0.0)
        risk = m.risk_maps()                                             # [V][T]
        grid = np.ascontiguousarray(risk.T)                              # device layout [t][z][y][x]
        if cell == p.POP_CELL_U8:
            assert np.isin(risk, (0.0, 1.0)).all()
            grid = grid.astype(np.uint8)
        n = 20000
        pos = pose.astype(np.float64) + rng.uniform(-5.6, 5.6, (n, 3)) * [1, 1, 0.4]
        pos[::50, 2] = rng.uniform(-1, 5, len(pos[::50]))                # below ground / above the ceiling
        t = rng.integers(0, cfg.nt, n).astype(np.int32)
        got = np.zeros(n, np.int8)
        dev_geom.dev_query(C.byref(cfg), grid.ctypes.data_as(C.c_void_p), pose.ctypes.data_as(C.POINTER(C.c_float)),
                           pos.ctypes.data_as(DP), t.ctypes.data_as(IP), n, got.ctypes.data_as(C.POINTER(C.c_int8)))
        want = m.query(pos, t)
        assert np.array_equal(got, want), np.flatnonzero(got != want)[:5]
        assert (want == 1).sum() > 50 and (want == 0).sum() > 1000 and ((want == -1).sum() > 100 or variant)
