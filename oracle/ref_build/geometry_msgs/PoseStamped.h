// geometry_msgs/PoseStamped.h — a STAND-IN (named in a planner header, never used on the path).  TEST INFRASTRUCTURE.
#ifndef POP_POSE_STAMPED_STANDIN_H
#define POP_POSE_STAMPED_STANDIN_H
#include <geometry_msgs/PolygonStamped.h>
namespace geometry_msgs { struct PoseStamped { Header header; typedef std::shared_ptr<const PoseStamped> ConstPtr; }; }
#endif
