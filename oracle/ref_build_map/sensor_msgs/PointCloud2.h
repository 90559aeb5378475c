// sensor_msgs/PointCloud2.h — a STAND-IN that carries the xyz points as floats (what pcl::fromROSMsg would decode).
// TEST INFRASTRUCTURE (oracle/_ref only).  Same guard as oracle/ref_build/sensor_msgs/PointCloud2.h, which it shadows.
#ifndef POP_POINTCLOUD2_STANDIN_H
#define POP_POINTCLOUD2_STANDIN_H
#include <std_msgs/Header.h>
#include <memory>
#include <vector>
namespace sensor_msgs {
struct PointCloud2 {
  std_msgs::Header header;
  std::vector<float> xyz;
  typedef std::shared_ptr<const PointCloud2> ConstPtr;
  typedef std::shared_ptr<PointCloud2> Ptr;
};
typedef PointCloud2::ConstPtr PointCloud2ConstPtr;
}  // namespace sensor_msgs
#endif
