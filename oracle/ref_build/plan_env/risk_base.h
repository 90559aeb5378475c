// plan_env/risk_base.h — a STAND-IN for the map class RiskHybridAstar queries (the real one needs ROS and PCL): the
// calls risk_hybrid_a_star.cpp makes, occupancy forwarded to a callback.  getClearOcccupancy(pos) is the current
// slice, as RiskBase::getClearOcccupancy(pos) = (pos, 0) (risk_base.cpp:252-254).  TEST INFRASTRUCTURE (oracle/_ref).
#ifndef POP_RISK_BASE_STANDIN_H
#define POP_RISK_BASE_STANDIN_H
#include <plan_env/fake_particle_risk_voxel.h>
class RiskBase : public FakeParticleRiskVoxel {
 public:
  typedef std::shared_ptr<RiskBase> Ptr;
  using FakeParticleRiskVoxel::getClearOcccupancy;
  int getClearOcccupancy(const Eigen::Vector3d& pos) { return getClearOcccupancy(pos, 0.0); }
};
#endif
