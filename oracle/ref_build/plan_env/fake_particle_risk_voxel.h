// plan_env/fake_particle_risk_voxel.h — a STAND-IN for the map class the hybrid A* queries (the real one needs ROS and
// PCL).  It offers the three calls fake_risk_hybrid_a_star.cpp makes and forwards the occupancy query to a callback,
// so that the reference's search runs UNMODIFIED on whatever grid the test supplies (the oracle's, whose queries are
// tested separately).  TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_FAKE_PARTICLE_RISK_VOXEL_STANDIN_H
#define POP_FAKE_PARTICLE_RISK_VOXEL_STANDIN_H
#include <Eigen/Eigen>
#include <cstdint>
#include <memory>
#include <ros/ros.h>

class FakeParticleRiskVoxel {
 public:
  typedef std::shared_ptr<FakeParticleRiskVoxel> Ptr;
  // getClearOcccupancy(pos, double t) (fake_particle_risk_voxel.cpp:341-346) for n positions
  typedef void (*query_dt_fn)(void* map, const double* pos, const double* dt, int n, int8_t* out);
  query_dt_fn query = nullptr;
  void* map = nullptr;
  Eigen::Vector3f pose;
  long long n_queries = 0;

  Eigen::Vector3f getMapCenter() const { return pose; }
  int getClearOcccupancy(const Eigen::Vector3d& pos, double t) {
    const double p[3] = {pos(0), pos(1), pos(2)};
    int8_t out = 0;
    query(map, p, &t, 1, &out);
    n_queries++;
    return out;
  }
  int getInflateOccupancy(const Eigen::Vector3d&) { return 0; }   // checkOccupancy() is never called on the path
};
#endif
