// traj_utils/BezierTraj.h — a STAND-IN for the header ROS generates from traj_utils/msg/BezierTraj.msg (same field
// names and types: int32 drone_id, traj_id; time start_time, pub_time; uint8 order; float32[] duration;
// geometry_msgs/Point[] cpts).  TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_BEZIER_TRAJ_STANDIN_H
#define POP_BEZIER_TRAJ_STANDIN_H
#include <cstdint>
#include <geometry_msgs/PolygonStamped.h>
#include <memory>
#include <vector>
namespace traj_utils {
struct BezierTraj {
  int32_t drone_id = 0, traj_id = 0;
  ros::Time start_time, pub_time;
  uint8_t order = 0;
  std::vector<float> duration;
  std::vector<geometry_msgs::Point> cpts;
  typedef std::shared_ptr<const BezierTraj> ConstPtr;
};
}  // namespace traj_utils
#endif
