// ros/ros.h — a STAND-IN for the slice of roscpp that path_searching's and traj_coordinator's sources touch:
// NodeHandle::param (parameter lookup with a default), inert subscribe / advertise / createTimer, ros::Time with a
// settable now(), and the ROS_* logging macros (discarded).  TEST INFRASTRUCTURE
// (oracle/_ref only); ROS itself is absent from this image.
#ifndef POP_ROS_STANDIN_H
#define POP_ROS_STANDIN_H
#include <iostream>   // roscpp's headers bring <iostream>, <math.h> and <stdlib.h>; the reference relies on that:
#include <math.h>     // map.cpp:477 prints with std::cout, fake_particle_risk_voxel.cpp:145 calls the global abs() on a
#include <stdlib.h>   // double — with libstdc++'s <stdlib.h>/<math.h> wrappers visible that is std::abs(double), not abs(int)
#include <map>
#include <string>
#include <type_traits>

namespace ros {
// time: now() is whatever the test set (ros::Time::set_now), so the reference's wall-clock reads are inputs
class Time {
 public:
  double sec_ = 0.0;
  Time() {}
  explicit Time(double s) : sec_(s) {}
  double toSec() const { return sec_; }
  Time& fromSec(double s) { sec_ = s; return *this; }
  static double& now_value() { static double v = 0.0; return v; }
  static void set_now(double s) { now_value() = s; }
  static Time now() { return Time(now_value()); }
};
class Duration {
 public:
  double sec_ = 0.0;
  Duration() {}
  explicit Duration(double s) : sec_(s) {}
  double toSec() const { return sec_; }
};
inline Duration operator-(const Time& a, const Time& b) { return Duration(a.sec_ - b.sec_); }
inline Time operator+(const Time& a, const Duration& d) { return Time(a.sec_ + d.sec_); }
struct TimerEvent {};
struct TransportHints { TransportHints& tcpNoDelay() { return *this; } };
class Subscriber {};
class Timer {};
class Publisher {
 public:
  template <class M> void publish(const M&) const {}
  int getNumSubscribers() const { return 0; }
};

class NodeHandle {
 public:
  std::map<std::string, double> values;
  // parameters of node handles the test cannot reach (a `ros::NodeHandle nh("~")` local to the reference's main())
  static std::map<std::string, double>& global_values() { static std::map<std::string, double> g; return g; }
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  template <class T>
  bool param(const std::string& name, T& var, const T& dflt) const {
    auto it = values.find(name);
    if (it == values.end()) {
      it = global_values().find(name);
      if (it == global_values().end()) { var = dflt; return false; }
    }
    if constexpr (std::is_arithmetic<T>::value) var = static_cast<T>(it->second);
    else var = dflt;   // strings (frame ids) are not numbers: the default stands
    return true;
  }
  template <class T>
  bool getParam(const std::string& name, T& var) const {
    auto it = values.find(name);
    if (it == values.end()) return false;
    var = static_cast<T>(it->second);
    return true;
  }
  // topics and timers are never driven: the tests call the callbacks directly
  template <class... A> Subscriber subscribe(const std::string&, int, A&&...) { return Subscriber(); }
  // node.subscribe<Msg>(topic, queue, &Class::callback, this): the message type is given, the rest deduced
  template <class M, class C, class Arg> Subscriber subscribe(const std::string&, int, void (C::*)(Arg), C*,
                                                               const TransportHints& = TransportHints()) { return Subscriber(); }
  template <class M> Publisher advertise(const std::string&, int, bool = false) { return Publisher(); }
  template <class... A> Timer createTimer(Duration, A&&...) { return Timer(); }
};
// process start-up and spinning: nothing to do, the tests call the callbacks
inline void init(int&, char**, const std::string&) {}
class AsyncSpinner { public: explicit AsyncSpinner(int) {} void start() {} void stop() {} };
inline void waitForShutdown() {}
inline void spin() {}
inline void spinOnce() {}
inline bool ok() { return false; }
}  // namespace ros

#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_ERROR_STREAM(x) ((void)0)
#endif
