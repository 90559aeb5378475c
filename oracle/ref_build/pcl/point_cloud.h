// pcl/point_cloud.h — a STAND-IN: the container the planner's rviz helpers fill (baseline_fake.cpp:87-119).  TEST INFRASTRUCTURE.
#ifndef POP_PCL_POINT_CLOUD_STANDIN_H
#define POP_PCL_POINT_CLOUD_STANDIN_H
#include <memory>
#include <vector>
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
template <class P> struct PointCloud {
  std::vector<P> points;
  typedef std::shared_ptr<PointCloud<P>> Ptr;
  void push_back(const P& p) { points.push_back(p); }
};
}  // namespace pcl
#endif
