// Stub of <ros/ros.h> for building the reference's CPU evaluator WITHOUT ROS.
// TEST INFRASTRUCTURE ONLY (oracle/_ref build). Written from scratch: just enough
// surface for the reference headers to parse; nothing here runs on the hot path.
#pragma once
#include <chrono>
#include <iostream>
#include <sstream>
#include <string>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <cmath>
#include <algorithm>
namespace ros {
struct Duration {
  double sec_ = 0.0;
  Duration() = default;
  explicit Duration(double s) : sec_(s) {}
  double toSec() const { return sec_; }
};
// Test clock: >= 0 makes ros::Time::now() return this value (the motion-update harness steps it by hand).
inline double& tsdf_stub_clock() { static double t = -1.0; return t; }
struct Time {
  double sec_ = 0.0;
  Time() = default;
  explicit Time(double s) : sec_(s) {}
  static Time now() {
    if (tsdf_stub_clock() >= 0.0) return Time(tsdf_stub_clock());
    using namespace std::chrono;
    return Time(duration<double>(steady_clock::now().time_since_epoch()).count());
  }
  double toSec() const { return sec_; }
  Duration operator-(const Time& o) const { return Duration(sec_ - o.sec_); }
};
// Just enough of the node API for the reference's own benchmark program (src/num_particles_eval.cpp:41-62) to run unmodified:
// private parameters come from environment variables ROSPARAM_<name>; an unset one leaves the program's default in place,
// like an unset ROS parameter.
inline void init(int&, char**, const std::string&) {}
class NodeHandle {
 public:
  NodeHandle() = default;
  explicit NodeHandle(const std::string&) {}
  template <typename T>
  bool getParam(const std::string& name, T& value) const {
    const char* e = std::getenv(("ROSPARAM_" + name).c_str());
    if (!e || !*e) return false;
    std::istringstream in(e);
    double v = 0.0;
    if (!(in >> v)) return false;
    value = static_cast<T>(v);
    return true;
  }
};
}  // namespace ros
#define TSDF_STUB_LOG(x) do { std::ostringstream _s; _s << x; std::cerr << _s.str() << std::endl; } while (0)
#define ROS_INFO_STREAM(x)  TSDF_STUB_LOG(x)
#define ROS_WARN_STREAM(x)  TSDF_STUB_LOG(x)
#define ROS_ERROR_STREAM(x) TSDF_STUB_LOG(x)
#define ROS_DEBUG_STREAM(x) do {} while (0)
