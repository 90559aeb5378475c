// ref_glue_risk.cpp — the second hybrid A* of the reference, RiskHybridAstar (path_searching/src/risk_hybrid_a_star.cpp,
// the planner over RiskBase maps = the real-perception path), in its own translation unit because its header shares
// the include guard of fake_risk_hybrid_a_star.h.  TEST INFRASTRUCTURE (part of oracle/_ref/libref.so); see ref_glue.cpp.
#define private public
#define protected public
#include <path_searching/risk_hybrid_a_star.h>
#undef private
#undef protected
#include "ref_astar_driver.hpp"

extern "C" int ref_astar_risk(const double* prm, FakeParticleRiskVoxel::query_dt_fn query, void* map, const float* pose3,
                              const double* start9, const double* goal3, double t_start, double* out_path, int cap,
                              int* out_n, int* stats) {
  return run_astar<RiskHybridAstar, RiskBase>(prm, query, map, pose3, start9, goal3, t_start, out_path, cap, out_n, stats);
}
