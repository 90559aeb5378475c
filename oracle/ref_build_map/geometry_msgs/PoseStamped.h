// geometry_msgs/PoseStamped.h — a STAND-IN with the pose fields MapBase::poseCallback reads (map.cpp:139-151).
// TEST INFRASTRUCTURE (oracle/_ref only).  Same guard as oracle/ref_build/geometry_msgs/PoseStamped.h, which it shadows.
#ifndef POP_POSE_STAMPED_STANDIN_H
#define POP_POSE_STAMPED_STANDIN_H
#include <geometry_msgs/PolygonStamped.h>
namespace geometry_msgs {
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; };
struct PoseStamped { Header header; Pose pose; typedef std::shared_ptr<const PoseStamped> ConstPtr; };
typedef std::shared_ptr<const PoseStamped> PoseStampedConstPtr;
}  // namespace geometry_msgs
#endif
