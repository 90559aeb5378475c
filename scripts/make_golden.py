"""Regenerates tests/golden/:

  reference_kats.json  the known-answer inputs/expectations of the reference's OWN tests, parsed
                       from its test sources under /root/reference (run in the build container;
                       the GPU box only ever sees the committed JSON):
                         traj_utils/test/test_bernstein.cpp   (Bezier evaluation, exact equality)
                         traj_opt/test/test_bezier_opt.cpp    (3-cube corridor QP fixture)
                         utils/separator/src/test_separator.cpp (separable 4-vs-5 point instance)
  oracle_c1.npz        outputs of the CPU oracle on BASELINE config 1 (66x66x20 @0.15 m, 20 static
                       pillars, one replan): occupied (voxel, slice) set, A* path, corridors,
                       QP control points, the published trajectory record.  The -m gpu tests
                       compare libpop against these committed vectors as well as against the
                       oracle run live.

    python scripts/make_golden.py [--reference /root/reference]
"""
import argparse
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
NUM = r"-?\d+(?:\.\d+)?"


def _strip_comments(src):
    return re.sub(r"//[^\n]*", "", src)


def _stream(src, name, after=0):
    """numbers of the Eigen comma initialiser `name << a, b, ...;` at/after offset `after`"""
    m = re.compile(r"\b" + re.escape(name) + r"\s*<<([^;]*);").search(src, after)
    assert m, f"initialiser of {name} not found"
    return [float(x) for x in re.findall(NUM, m.group(1))], m.end()


def parse_reference(ref):
    kat = {}
    # ---- Bezier known answers ----
    path = "traj_utils/test/test_bernstein.cpp"
    src = _strip_comments(open(os.path.join(ref, path)).read())
    piece, pos = _stream(src, "cpts")
    t0t1 = re.search(r"BernsteinPiece\(cpts,\s*(" + NUM + r"),\s*(" + NUM + r")\)", src)
    curve, _ = _stream(src, "cpts", pos)
    ts = re.search(r"std::vector<double>\s+ts\s*=\s*\{([^}]*)\}", src)
    coeff, _ = _stream(src, "A")
    vc, _ = _stream(src, "v_cpts")
    ac, _ = _stream(src, "a_cpts")
    expects = re.findall(r"EXPECT_EQ\(bp\.(get\w+)\((" + NUM + r")\),\s*Eigen::Vector3d(?:::Zero\(\)|\(([^)]*)\))\)", src)
    kat["bernstein"] = {
        "source": path,
        "piece_cpts": np.reshape(piece, (5, 3)).tolist(), "piece_t0": float(t0t1.group(1)), "piece_t1": float(t0t1.group(2)),
        "piece_expect": [{"fn": fn, "t": float(t), "value": [float(x) for x in re.findall(NUM, v)] if v else [0.0, 0.0, 0.0]}
                         for fn, t, v in expects],
        "coeff_mat": np.reshape(coeff, (5, 5)).tolist(),
        "vel_cpts": np.reshape(vc, (4, 3)).tolist(), "acc_cpts": np.reshape(ac, (3, 3)).tolist(),
        "curve_durations": [float(x) for x in re.findall(NUM, ts.group(1))],
        "curve_cpts": np.reshape(curve, (15, 3)).tolist(),
        "curve_expect": {"duration": 6.0, "pieces": 3, "max_vel_rate_gt": 1.5},
    }
    # ---- QP fixture ----
    path = "traj_opt/test/test_bezier_opt.cpp"
    src = _strip_comments(open(os.path.join(ref, path)).read())
    at = src.index("class BezierOptTest2")
    cubes = []
    for name in ("cube1", "cube2", "cube3"):
        v, _ = _stream(src, name, at)
        cubes.append(np.reshape(v, (6, 4)).tolist())
    start, _ = _stream(src, "start", at)
    end, _ = _stream(src, "end", at)
    t = [float(x) for x in re.findall(r"t\.push_back\((" + NUM + r")\)", src[at:])][:3]
    kat["bezier_opt"] = {"source": path, "polys": cubes, "start": start, "end": end, "t": t,
                         "tolerance": 1e-3, "note": "reference asserts boundary conditions and continuity at 1e-3"}
    # ---- separator instance ----
    path = "utils/separator/src/test_separator.cpp"
    src = _strip_comments(open(os.path.join(ref, path)).read())
    pts = {k: [[float(x) for x in re.findall(NUM, v)] for v in re.findall(r"points" + k + r"\.col\(\d\)\s*=\s*Eigen::Vector3d\(([^)]*)\)", src)]
           for k in ("A", "B")}
    kat["separator"] = {"source": path, "points_a": pts["A"], "points_b": pts["B"], "expect_solved": True}
    return kat


def oracle_c1():
    import pred_occ_planner_b200 as p
    from tests import cases, oracle_cycle, oracle_lib
    orc = oracle_lib.load()
    cfg = cases.cfg_c1(p.POP_CELL_F32)
    cloud, cyl = cases.world_c1()
    start = np.zeros(9)
    start[:3] = [-4, 0, 1.0]
    goal = np.array([4.0, 0, 1.0])
    om = orc.map(cfg)
    om.set_pose(start[:3].astype(np.float32))
    om.update_gt(cloud, cyl, 10.0)
    grid = om.risk_maps()
    occ = np.argwhere(grid > 0).astype(np.int32)          # rows (voxel index, slice)
    ret, path, stats = om.astar(start, goal, 0.02)
    npoly, hp, rows, lgoal = om.corridor(path, start, 10.02)
    fin = np.zeros(9)
    fin[:6] = lgoal
    st, X, it = orc.qp_solve(cfg, start, fin, [cfg.corridor_tau] * npoly, hp=hp, rows=rows,
                             vmax=cfg.opt_max_vel, amax=cfg.opt_max_acc)
    rst, rec, _ = oracle_cycle.replan(om, None, 0, start, goal, 10.02)
    rng = np.random.default_rng(0)
    qpos = rng.uniform([-9, -5, 0], [1, 5, 3], (512, 3))
    qt = rng.integers(0, cfg.nt, 512).astype(np.int32)
    return dict(start=start, goal=goal, t_map=10.0, t_plan=10.02, occupied=occ, astar_ret=ret, astar_path=path,
                n_poly=npoly, hpoly=hp[:npoly], rows=rows[:npoly], local_goal=lgoal, qp_status=st, qp_iters=it,
                qp_cpts=X, replan_status=rst, record_cpts=rec["cpts"], record_n_pieces=rec["n_pieces"],
                query_pos=qpos, query_t=qt, query_out=om.query(qpos, qt))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    kat = parse_reference(args.reference)
    with open(os.path.join(OUT, "reference_kats.json"), "w") as f:
        json.dump(kat, f, indent=1)
    np.savez_compressed(os.path.join(OUT, "oracle_c1.npz"), **oracle_c1())
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
