// ref_glue_planner_risk.cpp — the reference's real-perception planner, BaselinePlanner::replan (plan_manager/src/
// baseline.cpp: RiskHybridAstar on a RiskBase map), in oracle/_ref/libref.so.  TEST INFRASTRUCTURE; see
// ref_glue_planner.cpp and ref_planner_driver.hpp.
#define private public
#define protected public
#include <plan_manager/baseline.h>
#undef private
#undef protected
#include "ref_planner_driver.hpp"

extern "C" int ref_replan_risk(const pop_config* cfg, FakeParticleRiskVoxel::query_dt_fn query, FakeParticleRiskVoxel::obstacle_fn obstacles,
                               dense_osqp_fn osqp, void* map, const float* pose3, double map_time, const double* offs,
                               const int32_t* counts, const pop_traj_record* recs, double now, double t, const double* start9,
                               const double* goal3, int* out_M, double* out_dur, double* out_cpts, int* stats) {
  return run_replan<BaselinePlanner, BaselineParameters>(cfg, query, obstacles, osqp, map, pose3, map_time, offs, counts, recs, now,
                                                         t, start9, goal3, out_M, out_dur, out_cpts, stats);
}

extern "C" void ref_traj_safe_risk(const pop_config* cfg, FakeParticleRiskVoxel::query_dt_fn query, void* map, const float* pose3,
                                   double map_time, int M, const double* dur, const double* cpts, double traj_start_time, int nq,
                                   const double* now, const double* T, int8_t* out) {
  run_traj_safe<BaselinePlanner, BaselineParameters>(cfg, query, map, pose3, map_time, M, dur, cpts, traj_start_time, nq, now, T, out);
}
