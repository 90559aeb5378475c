// nav_msgs/Odometry.h — a STAND-IN (named in a planner header, never used on the path).  TEST INFRASTRUCTURE.
#ifndef POP_ODOMETRY_STANDIN_H
#define POP_ODOMETRY_STANDIN_H
#include <geometry_msgs/PolygonStamped.h>
namespace nav_msgs { struct Odometry { geometry_msgs::Header header; typedef std::shared_ptr<const Odometry> ConstPtr; }; }
#endif
