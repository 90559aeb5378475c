// pcl/point_types.h — a STAND-IN: the point structs plan_env names.  TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_PCL_POINT_TYPES_STANDIN_H
#define POP_PCL_POINT_TYPES_STANDIN_H
#include <cstdint>
namespace pcl {
struct PointXYZ {
  float x = 0, y = 0, z = 0;
  PointXYZ() {}
  PointXYZ(float a, float b, float c) : x(a), y(b), z(c) {}
};
struct PointXYZRGB { float x = 0, y = 0, z = 0; uint8_t r = 0, g = 0, b = 0; };
struct PointXYZRGBA { float x = 0, y = 0, z = 0; uint8_t r = 0, g = 0, b = 0, a = 0; };
struct PointNormal { float x = 0, y = 0, z = 0, normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0; };
struct PointXYZINormal { float x = 0, y = 0, z = 0, intensity = 0, normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0; };
}  // namespace pcl
#endif
