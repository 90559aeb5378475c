// traj_utils/visualizer.hpp — a STAND-IN for the rviz publisher class: the four calls the planner makes, doing nothing.
// TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_VISUALIZER_STANDIN_H
#define POP_VISUALIZER_STANDIN_H
#include <Eigen/Eigen>
#include <memory>
#include <ros/ros.h>
#include <string>
#include <vector>
namespace visualizer {
class Visualizer {
 public:
  typedef std::shared_ptr<Visualizer> Ptr;
  Visualizer(ros::NodeHandle&, const std::string&) {}
  template <class T> void visualizePolytope(const T&) {}
  void visualizeStartGoal(const Eigen::Vector3d&) {}
  void visualizeAstarPath(const std::vector<Eigen::Vector3d>&) {}
  void visualizeAstarPathXYT(const std::vector<Eigen::Vector3d>&, double) {}
};
}  // namespace visualizer
#endif
