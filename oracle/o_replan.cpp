// o_replan.cpp — oracle: one replan cycle for one agent (TEST INFRASTRUCTURE).
// Follows FakeBaselinePlanner::replan (plan_manager/src/baseline_fake.cpp:253-453) and
// ParticleATC::isSafeAfterOpt (traj_coordinator/src/particles.cpp:223-283); the trajectory record
// is what FiniteStateMachineFake::publishTrajectory broadcasts
// (plan_manager/src/fake_plan_manager.cpp:364-399; durations narrowed to float32 by
// traj_utils/msg/BezierTraj.msg:8).
// PINNED end to end against the reference's own planners compiled into oracle/_ref/libref.so (baseline_fake.cpp and
// baseline.cpp, with the map and the OSQP solve answered through callbacks by this oracle): same successes and failures,
// A* expansions and piece counts, control points at the solver's tolerance — tests/test_ref_pins.py.
#include "oracle.hpp"

namespace orc {

std::vector<double>* g_sep_trace = nullptr;

// isSafeAfterOpt: ego control points vs every active peer's remaining control points
bool safe_after_opt(const Swarm& s, int ego_id, const std::vector<double>& cpts, int M, double t0) {
  std::vector<V3d> A;
  for (int i = 0; i < 5 * M; i++) A.push_back(V3d{cpts[i * 3], cpts[i * 3 + 1], cpts[i * 3 + 2]});
  for (int id = 0; id < s.n_agents; id++) {
    const pop_traj_record& r = s.trajs[id];
    if (id == ego_id || r.n_pieces <= 0) continue;
    double t_end = r.start_time;
    for (int i = 0; i < r.n_pieces; i++) t_end += r.duration[i];
    if (r.start_time < t0 && t0 < t_end) {
      double rem = t0 - r.start_time;  // locatePiece
      int piece = r.n_pieces - 1;
      for (int i = 0; i < r.n_pieces; i++) {
        rem -= static_cast<double>(r.duration[i]);
        if (rem < 0) { piece = i; break; }
      }
      std::vector<V3d> B;
      for (int i = piece * 5; i < r.n_pieces * 5; i++) B.push_back(V3d{r.cpts[i * 3], r.cpts[i * 3 + 1], r.cpts[i * 3 + 2]});
      if (g_sep_trace) {   // test hook: the point sets instead of the LP (see orc_safe_after_opt_sets)
        g_sep_trace->push_back(static_cast<double>(A.size()));
        g_sep_trace->push_back(static_cast<double>(B.size()));
        for (const V3d& p : B) { g_sep_trace->push_back(p.x); g_sep_trace->push_back(p.y); g_sep_trace->push_back(p.z); }
        continue;
      }
      if (!separable(A, B)) return false;
    }
  }
  return true;
}

// BaselinePlanner::isTrajSafe(T) (plan_manager/src/baseline.cpp:45-68 = baseline_fake.cpp:53-76): the planner's own
// trajectory (pieces of `tau` seconds each — the double corridor_tau the optimiser was given, not the float32 of the
// wire record) sampled every 0.1 s from `t_now` to min(T, duration); false at the first sample whose occupancy query
// answers 1.  The sample times accumulate (`t += 0.1`), so they are reproduced by the same running sum.
bool traj_safe(const Grid& g, const pop_traj_record& r, double tau, double t_now, double T, double* t_unsafe) {
  if (t_unsafe) *t_unsafe = -1.0;
  double t0 = t_now - r.start_time;
  if (t0 < 0) t0 = 0;
  if (t0 > T) return true;
  if (r.n_pieces <= 0) return true;   // an empty traj_ has duration 0: the loop below never runs
  const Bezier traj(std::vector<double>(r.n_pieces, tau), std::vector<double>(r.cpts, r.cpts + 15 * r.n_pieces));
  const double dur = traj.T;
  T = T > dur ? dur : T;
  for (double t = t0; t < T; t += 0.1) {
    const V3d pos = traj.pos(t);
    const double dt = t + r.start_time - g.t_map;
    if (query_dt(g, pos, dt) == 1) {
      if (t_unsafe) *t_unsafe = t;
      return false;
    }
  }
  return true;
}

// returns POP_REPLAN_*; fills `out` on success
int replan(const Grid& g, const pop_config& c, const Swarm* swarm, int ego_id, const double start[9],
           const double goal[3], double t_plan, pop_traj_record* out, int* stats) {
  AstarResult ar = astar_replan(g, c, start, goal, t_plan - g.t_map, 0);
  if (stats) { stats[0] = ar.ret; stats[1] = ar.iters; stats[2] = 0; stats[3] = 0; }
  if (ar.ret == POP_ASTAR_NO_PATH || ar.path.empty()) return POP_REPLAN_FAIL_ASTAR;
  CorridorResult cr = corridor_stage(g, c, ar.path, start, t_plan);
  if (!cr.ok || cr.polys.empty()) return POP_REPLAN_FAIL_CORRIDOR;
  const int M = static_cast<int>(cr.polys.size());
  if (stats) stats[2] = M;
  double fin[9] = {cr.goal[0], cr.goal[1], cr.goal[2], cr.goal[3], cr.goal[4], cr.goal[5], 0, 0, 0};
  QpProblem qp;
  qp_setup(qp, start, fin, std::vector<double>(M, c.corridor_tau), cr.polys, c.opt_max_vel, c.opt_max_acc);
  OsqpSettings S;
  S.eps_abs = c.osqp_eps_abs; S.eps_rel = c.osqp_eps_rel; S.max_iter = c.osqp_max_iter;
  S.check_termination = c.osqp_check_termination; S.adaptive_rho_interval = c.osqp_adaptive_rho_interval;
  OsqpResult qr = osqp_solve(qp, S);
  if (stats) stats[3] = qr.iters;
  if (!(qr.status == 1 || qr.status == 2)) return POP_REPLAN_FAIL_QP;
  if (swarm && !safe_after_opt(*swarm, ego_id, qr.x, M, t_plan)) return POP_REPLAN_FAIL_SEPARATION;
  if (out) {
    std::memset(out, 0, sizeof(*out));
    out->drone_id = ego_id;
    out->n_pieces = M;
    out->valid = 1;
    out->start_time = t_plan;
    for (int i = 0; i < M; i++) out->duration[i] = static_cast<float>(c.corridor_tau);
    for (int i = 0; i < 15 * M; i++) out->cpts[i] = qr.x[i];
  }
  return POP_REPLAN_OK;
}

}  // namespace orc
