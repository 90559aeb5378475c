// pcl/point_cloud.h — a STAND-IN: the container of plan_env's clouds.  TEST INFRASTRUCTURE (oracle/_ref only).
// (Shadows oracle/ref_build/pcl/point_cloud.h in the map builds; same guard, so only one of them is seen.)
#ifndef POP_PCL_POINT_CLOUD_STANDIN_H
#define POP_PCL_POINT_CLOUD_STANDIN_H
#include <pcl/point_types.h>
#include <memory>
#include <string>
#include <vector>
namespace pcl {
template <class P> struct PointCloud {
  std::vector<P> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  struct Header { std::string frame_id; uint64_t stamp = 0; uint32_t seq = 0; } header;
  typedef std::shared_ptr<PointCloud<P>> Ptr;
  typedef std::shared_ptr<const PointCloud<P>> ConstPtr;
  void push_back(const P& p) { points.push_back(p); }
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = height = 0; }
  typename std::vector<P>::iterator begin() { return points.begin(); }
  typename std::vector<P>::iterator end() { return points.end(); }
  typename std::vector<P>::const_iterator begin() const { return points.begin(); }
  typename std::vector<P>::const_iterator end() const { return points.end(); }
  P& operator[](size_t i) { return points[i]; }
  const P& operator[](size_t i) const { return points[i]; }
};
}  // namespace pcl
#endif
