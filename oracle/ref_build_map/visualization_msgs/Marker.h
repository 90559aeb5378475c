// visualization_msgs/Marker.h — a STAND-IN with the fields FakeParticleRiskVoxel::groundTruthStateCallback reads
// (fake_particle_risk_voxel.cpp:244-264).  TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_MARKER_STANDIN_H
#define POP_MARKER_STANDIN_H
#include <geometry_msgs/PoseStamped.h>
#include <vector>
namespace visualization_msgs {
struct Marker {
  geometry_msgs::Header header;
  int id = 0, type = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std::vector<geometry_msgs::Point> points;
};
}  // namespace visualization_msgs
#endif
