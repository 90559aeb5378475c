// ref_astar_driver.hpp — shared by ref_glue.cpp (FakeRiskHybridAstar) and ref_glue_risk.cpp (RiskHybridAstar): the two
// reference headers use the same include guard, so each variant lives in its own translation unit.  TEST INFRASTRUCTURE.
//
// Drives a hybrid A* class as (Fake)BaselinePlanner::replan does (baseline_fake.cpp:280-305, baseline.cpp:280-305):
// reset, search with init = true, once more with init = false on NO_PATH, then getPathWithVel(corridor_tau).
// prm = max_tau, init_max_tau, max_vel, max_acc, w_time, horizon, resolution_astar, time_resolution, lambda_heu,
// allocate_num, check_num, tolerance, corridor_tau (search/... of fake_risk_hybrid_a_star.cpp:62-82).
// start9 = pos, vel, acc.  out_path n x 6; stats = iterations, nodes used, retried, occupancy queries.  Returns ASTAR_RET.
#pragma once
#include <iostream>
#include <sstream>

template <class Astar, class Map>
static int run_astar(const double* prm, FakeParticleRiskVoxel::query_dt_fn query, void* map, const float* pose3,
                     const double* start9, const double* goal3, double t_start, double* out_path, int cap, int* out_n,
                     int* stats) {
  ros::NodeHandle nh;
  const char* names[] = {"search/max_tau", "search/init_max_tau", "search/max_vel", "search/max_acc", "search/w_time",
                         "search/horizon", "search/resolution_astar", "search/time_resolution", "search/lambda_heu",
                         "search/allocate_num", "search/check_num", "search/tolerance"};
  for (int i = 0; i < 12; i++) nh.values[names[i]] = prm[i];
  std::shared_ptr<Map> m(new Map());
  m->query = query; m->map = map;
  m->pose = Eigen::Vector3f(pose3[0], pose3[1], pose3[2]);
  std::ostringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  Astar a;
  a.setParam(nh);
  a.setEnvironment(m);
  a.init(Eigen::Vector3d::Zero(), Eigen::Vector3d(10, 10, 4));          // baseline_fake.cpp:28
  const Eigen::Vector3d sp(start9[0], start9[1], start9[2]), sv(start9[3], start9[4], start9[5]),
      sa(start9[6], start9[7], start9[8]), gp(goal3[0], goal3[1], goal3[2]);
  a.reset();
  int retried = 0;
  ASTAR_RET rst = a.search(sp, sv, sa, gp, Eigen::Vector3d(0, 0, 0), true, true, t_start);
  if (rst == 0) {
    a.reset();
    rst = a.search(sp, sv, sa, gp, Eigen::Vector3d(0, 0, 0), false, true, t_start);
    retried = 1;
  }
  int n = 0;
  if (rst != NO_PATH) {
    const std::vector<Eigen::Matrix<double, 6, 1>> path = a.getPathWithVel(prm[12]);
    n = static_cast<int>(path.size());
    for (int i = 0; i < n && i < cap; i++) for (int k = 0; k < 6; k++) out_path[6 * i + k] = path[i](k);
  }
  std::cout.rdbuf(old);
  if (out_n) *out_n = n;
  if (stats) { stats[0] = a.iter_num_; stats[1] = a.use_node_num_; stats[2] = retried; stats[3] = static_cast<int>(m->n_queries); }
  return static_cast<int>(rst);
}

