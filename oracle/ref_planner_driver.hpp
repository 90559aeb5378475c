// ref_planner_driver.hpp — shared by ref_glue_planner.cpp (FakeBaselinePlanner) and ref_glue_planner_risk.cpp
// (BaselinePlanner): the two planner headers pull in A* headers with the same include guard, so each lives in its own
// translation unit.  TEST INFRASTRUCTURE.  Included after the planner's header (with private members opened).
//
// One replan(t, start, goal) of drone cfg->agent_id_base in a swarm whose peers' body particles and trajectories have
// been received (recs may be NULL: no peers).  Parameters come from cfg as the launch files would set them; `now` is
// what ros::Time::now() returns during the call.  Returns 1 / 0 = replan's bool; on success M pieces, durations and
// 5M x 3 control points.  stats = A* iterations, A* nodes, pieces, QP iterations.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include "../include/pop.h"

extern "C" void ref_feed_atc(ParticleATC& atc, int n, int ego_id, const double* offs, int max_body, const int32_t* counts,
                  const pop_traj_record* recs, double now);   // ref_glue_particles.cpp

typedef int (*dense_osqp_fn)(const pop_config* cfg, int n, int m, const double* Q, const double* A, const double* l,
                             const double* u, double* x, int* iters);
static dense_osqp_fn g_osqp = nullptr;
static const pop_config* g_cfg = nullptr;
static int g_qp_iters = 0;
static int solve_adapter(int n, int m, const double* P, const double*, const double* A, const double* l, const double* u,
                         double, double, double* x) {
  return g_osqp(g_cfg, n, m, P, A, l, u, x, &g_qp_iters);
}


static void fill_params(ros::NodeHandle& nh, const pop_config& c) {
  nh.values["search/max_tau"] = c.max_tau; nh.values["search/init_max_tau"] = c.max_tau;
  nh.values["search/max_vel"] = c.max_vel; nh.values["search/max_acc"] = c.max_acc;
  nh.values["search/w_time"] = c.w_time; nh.values["search/horizon"] = c.horizon;
  nh.values["search/resolution_astar"] = c.resolution_astar; nh.values["search/time_resolution"] = c.time_resolution_astar;
  nh.values["search/lambda_heu"] = c.lambda_heu; nh.values["search/allocate_num"] = c.allocate_num;
  nh.values["search/check_num"] = c.check_num; nh.values["search/tolerance"] = c.tolerance;
  nh.values["planner/corridor_tau"] = c.corridor_tau; nh.values["corridor/init_range"] = c.init_range;
  nh.values["corridor/shrink_size"] = c.shrink_size;
  nh.values["optimizer/max_vel_optimization"] = c.opt_max_vel; nh.values["optimizer/max_acc_optimization"] = c.opt_max_acc;
  nh.values["drone_id"] = c.agent_id_base; nh.values["swarm/num_robots"] = c.n_agents_total;
  nh.values["swarm/drone_size_x"] = c.drone_size[0]; nh.values["swarm/drone_size_y"] = c.drone_size[1];
  nh.values["swarm/drone_size_z"] = c.drone_size[2];
  nh.values["swarm/replan_risk_rate"] = c.replan_risk_rate; nh.values["swarm/num_resample"] = c.num_resample;
}

template <class Planner, class Params>
static int run_replan(const pop_config* cfg, FakeParticleRiskVoxel::query_dt_fn query, FakeParticleRiskVoxel::obstacle_fn obstacles,
                      dense_osqp_fn osqp, void* map, const float* pose3, double map_time, const double* offs,
                      const int32_t* counts, const pop_traj_record* recs, double now, double t, const double* start9,
                      const double* goal3, int* out_M, double* out_dur, double* out_cpts, int* stats) {
  std::ostringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  ros::Time::set_now(now);
  ros::NodeHandle nh;
  const pop_config& c = *cfg;
  fill_params(nh, c);
  Params prm(nh);
  Planner planner(nh, nh, nh, prm);
  planner.init();
  planner.map_->query = query; planner.map_->obstacles = obstacles; planner.map_->map = map;
  planner.map_->pose = Eigen::Vector3f(pose3[0], pose3[1], pose3[2]);
  planner.map_->map_time = map_time;
  if (recs) ref_feed_atc(*planner.collision_avoider_, c.n_agents_total, c.agent_id_base, offs, POP_MAX_BODY, counts, recs, now);
  g_osqp = osqp; g_cfg = cfg; g_qp_iters = 0;
  pop_iosqp_solver_slot() = &solve_adapter;
  const Eigen::Vector3d sp(start9[0], start9[1], start9[2]), sv(start9[3], start9[4], start9[5]),
      sa(start9[6], start9[7], start9[8]), gp(goal3[0], goal3[1], goal3[2]);
  const bool ok = planner.replan(t, sp, sv, sa, gp);
  pop_iosqp_solver_slot() = nullptr;
  std::cout.rdbuf(old);
  if (getenv("POP_REF_LOG")) std::cerr << sink.str();   // the reference's own console log of this replan
  int M = 0;
  if (ok) {
    const Bernstein::Bezier traj = planner.getTrajectory();
    M = traj.getNumPieces();
    const Eigen::MatrixXd cp = traj.getCtrlPoints();
    for (int i = 0; i < 5 * M; i++) for (int k = 0; k < 3; k++) out_cpts[3 * i + k] = cp(i, k);
    for (int i = 0; i < M; i++) out_dur[i] = traj.t_[i];
  }
  if (out_M) *out_M = M;
  if (stats) { stats[0] = planner.a_star_->iter_num_; stats[1] = planner.a_star_->use_node_num_; stats[2] = M; stats[3] = g_qp_iters; }
  return ok ? 1 : 0;
}

// isTrajSafe(T) (baseline_fake.cpp:53-76 / baseline.cpp:45-68) of a planner that holds the trajectory (dur[M], cpts 5M x 3,
// started at traj_start_time) over the map behind `query`, asked at nq wall-clock times now[q] with horizons T[q].
template <class Planner, class Params>
static void run_traj_safe(const pop_config* cfg, FakeParticleRiskVoxel::query_dt_fn query, void* map, const float* pose3,
                          double map_time, int M, const double* dur, const double* cpts, double traj_start_time, int nq,
                          const double* now, const double* T, int8_t* out) {
  std::ostringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  ros::NodeHandle nh;
  fill_params(nh, *cfg);
  Params prm(nh);
  Planner planner(nh, nh, nh, prm);
  planner.init();
  planner.map_->query = query; planner.map_->map = map;
  planner.map_->pose = Eigen::Vector3f(pose3[0], pose3[1], pose3[2]);
  planner.map_->map_time = map_time;
  Eigen::MatrixX3d c(5 * M, 3);
  for (int i = 0; i < 5 * M; i++) for (int k = 0; k < 3; k++) c(i, k) = cpts[3 * i + k];
  planner.traj_ = Bernstein::Bezier(std::vector<double>(dur, dur + M), c);
  planner.traj_start_time_ = traj_start_time;
  for (int q = 0; q < nq; q++) {
    ros::Time::set_now(now[q]);
    out[q] = planner.isTrajSafe(T[q]) ? 1 : 0;
  }
  std::cout.rdbuf(old);
}
