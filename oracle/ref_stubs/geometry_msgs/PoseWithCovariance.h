#pragma once
#include <array>
#include <geometry_msgs/Pose.h>
namespace geometry_msgs { struct PoseWithCovariance { Pose pose; std::array<double, 36> covariance{}; }; }
