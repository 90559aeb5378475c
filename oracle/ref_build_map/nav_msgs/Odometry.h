// nav_msgs/Odometry.h — a STAND-IN with the pose fields MapBase::odomCallback reads (map.cpp:158-170).
// TEST INFRASTRUCTURE (oracle/_ref only).  Same guard as oracle/ref_build/nav_msgs/Odometry.h, which it shadows.
#ifndef POP_ODOMETRY_STANDIN_H
#define POP_ODOMETRY_STANDIN_H
#include <geometry_msgs/PoseStamped.h>
namespace geometry_msgs { struct Twist { Vector3 linear, angular; }; struct TwistWithCovariance { Twist twist; }; }
namespace nav_msgs { struct Odometry { geometry_msgs::Header header; geometry_msgs::PoseWithCovariance pose; geometry_msgs::TwistWithCovariance twist; typedef std::shared_ptr<const Odometry> ConstPtr; }; typedef std::shared_ptr<const Odometry> OdometryConstPtr; }
#endif
