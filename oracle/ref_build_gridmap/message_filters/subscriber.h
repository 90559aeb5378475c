// message_filters — stand-in (TEST INFRASTRUCTURE) for what plan_env/src/grid_map.cpp names while wiring its topics:
// Subscriber<M>, Synchronizer<Policy>, the two policies, registerCallback.  Inert: the test calls the callbacks itself.
#pragma once
#include <memory>
#include <string>

#include <ros/ros.h>

namespace message_filters {
template <class M>
class Subscriber {
 public:
  Subscriber() {}
  Subscriber(ros::NodeHandle&, const std::string&, int, const ros::TransportHints& = ros::TransportHints()) {}
};
namespace sync_policies {
template <class A, class B> struct ApproximateTime { explicit ApproximateTime(int) {} };
template <class A, class B> struct ExactTime { explicit ExactTime(int) {} };
}  // namespace sync_policies
template <class Policy>
class Synchronizer {
 public:
  template <class A, class B> Synchronizer(const Policy&, A&, B&) {}
  template <class F> void registerCallback(const F&) {}
};
template <class A, class B>
class TimeSynchronizer {
 public:
  template <class... X> explicit TimeSynchronizer(X&&...) {}
  template <class F> void registerCallback(const F&) {}
};
}  // namespace message_filters
