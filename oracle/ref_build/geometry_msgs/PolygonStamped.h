// geometry_msgs/PolygonStamped.h — a STAND-IN: the fields traj_coordinator/src/particles.cpp reads and writes
// (std_msgs/Header, geometry_msgs/Polygon of float32 Point32).  TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_POLYGON_STAMPED_STANDIN_H
#define POP_POLYGON_STAMPED_STANDIN_H
#include <memory>
#include <ros/ros.h>
#include <string>
#include <vector>
namespace geometry_msgs {
struct Point32 { float x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Polygon { std::vector<Point32> points; };
struct Header { ros::Time stamp; std::string frame_id; };
struct PolygonStamped {
  Header header;
  Polygon polygon;
  typedef std::shared_ptr<const PolygonStamped> ConstPtr;
};
}  // namespace geometry_msgs
#endif
