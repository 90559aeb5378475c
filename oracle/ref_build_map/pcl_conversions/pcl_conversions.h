// pcl_conversions/pcl_conversions.h — a STAND-IN: fromROSMsg unpacks the floats of the stand-in PointCloud2; toROSMsg
// only feeds rviz.  TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_PCL_CONVERSIONS_STANDIN_H
#define POP_PCL_CONVERSIONS_STANDIN_H
#include <pcl/point_cloud.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
template <class P> void toROSMsg(const PointCloud<P>&, sensor_msgs::PointCloud2&) {}
template <class P> void fromROSMsg(const sensor_msgs::PointCloud2& m, PointCloud<P>& c) {
  c.points.resize(m.xyz.size() / 3);
  for (size_t i = 0; i < c.points.size(); i++) { c.points[i].x = m.xyz[3 * i]; c.points[i].y = m.xyz[3 * i + 1]; c.points[i].z = m.xyz[3 * i + 2]; }
  c.width = static_cast<uint32_t>(c.points.size());
  c.height = 1;
}
}  // namespace pcl
#endif
