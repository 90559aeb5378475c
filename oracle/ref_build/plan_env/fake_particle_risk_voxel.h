// plan_env/fake_particle_risk_voxel.h — a STAND-IN for the map class the hybrid A* queries (the real one needs ROS and
// PCL).  It offers the calls fake_risk_hybrid_a_star.cpp and baseline_fake.cpp make and forwards the occupancy query
// and the obstacle-point scan to callbacks,
// so that the reference's search runs UNMODIFIED on whatever grid the test supplies (the oracle's, whose queries are
// tested separately).  TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_FAKE_PARTICLE_RISK_VOXEL_STANDIN_H
#define POP_FAKE_PARTICLE_RISK_VOXEL_STANDIN_H
#include <Eigen/Eigen>
#include <cstdint>
#include <memory>
#include <ros/ros.h>
#include <traj_coordinator/particle.hpp>
#include <vector>

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

  // what FakeBaselinePlanner calls besides the queries (baseline_fake.cpp:19-38, 283, 341)
  // getObstaclePoints(points, t0, t1, lower, higher) (map.cpp:463-531): returns the count, writes up to cap points
  typedef int (*obstacle_fn)(void* map, double t0, double t1, const double* lc, const double* hc, double* out, int cap);
  obstacle_fn obstacles = nullptr;
  double map_time = 0.0;
  void init(ros::NodeHandle&) {}
  void setCoordinator(ParticleATC::Ptr) {}
  ros::Time getMapTime() const { return ros::Time(map_time); }
  void getObstaclePoints(std::vector<Eigen::Vector3d>& pts, double t0, double t1, const Eigen::Vector3d& lc,
                         const Eigen::Vector3d& hc) {
    const double l[3] = {lc(0), lc(1), lc(2)}, h[3] = {hc(0), hc(1), hc(2)};
    std::vector<double> buf(3 * 4096);
    int n = obstacles(map, t0, t1, l, h, buf.data(), 4096);
    if (n > 4096) { buf.resize(3 * static_cast<size_t>(n)); n = obstacles(map, t0, t1, l, h, buf.data(), n); }
    for (int i = 0; i < n; i++) pts.push_back(Eigen::Vector3d(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]));
    // The planner views the whole vector as one 3 x N matrix through &pts[0] (baseline_fake.cpp:343), which relies on
    // real Eigen's fixed-size vectors being contiguous PODs.  The stand-in's vectors own heap storage, so the points
    // are additionally laid out contiguously behind the first element's storage (its shape stays 3 x 1).
    if (n > 0) pts[0].v_.assign(buf.begin(), buf.begin() + 3 * static_cast<size_t>(n));
  }
};
#endif
