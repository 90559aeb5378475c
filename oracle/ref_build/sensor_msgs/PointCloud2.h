// sensor_msgs/PointCloud2.h — a STAND-IN: the planner only fills a header and publishes it for rviz.  TEST INFRASTRUCTURE.
#ifndef POP_POINTCLOUD2_STANDIN_H
#define POP_POINTCLOUD2_STANDIN_H
#include <geometry_msgs/PolygonStamped.h>
namespace sensor_msgs { struct PointCloud2 { geometry_msgs::Header header; typedef std::shared_ptr<const PointCloud2> ConstPtr; }; }
#endif
