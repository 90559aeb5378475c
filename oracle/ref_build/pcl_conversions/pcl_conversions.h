// pcl_conversions/pcl_conversions.h — a STAND-IN: toROSMsg only feeds rviz.  TEST INFRASTRUCTURE.
#ifndef POP_PCL_CONVERSIONS_STANDIN_H
#define POP_PCL_CONVERSIONS_STANDIN_H
#include <pcl/point_cloud.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl { template <class P> void toROSMsg(const PointCloud<P>&, sensor_msgs::PointCloud2&) {} }
#endif
