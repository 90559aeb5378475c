// ros/ros.h — a STAND-IN for the slice of roscpp that path_searching's headers and fake_risk_hybrid_a_star.cpp touch:
// NodeHandle::param (parameter lookup with a default) and the ROS_* logging macros (discarded).  TEST INFRASTRUCTURE
// (oracle/_ref only); ROS itself is absent from this image.
#ifndef POP_ROS_STANDIN_H
#define POP_ROS_STANDIN_H
#include <map>
#include <string>

namespace ros {
class NodeHandle {
 public:
  std::map<std::string, double> values;
  template <class T>
  bool param(const std::string& name, T& var, const T& dflt) const {
    auto it = values.find(name);
    if (it == values.end()) { var = dflt; return false; }
    var = static_cast<T>(it->second);
    return true;
  }
};
}  // namespace ros

#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_ERROR_STREAM(x) ((void)0)
#endif
