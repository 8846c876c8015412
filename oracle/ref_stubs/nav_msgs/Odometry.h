#pragma once
#include <geometry_msgs/PoseWithCovariance.h>
#include <geometry_msgs/TransformStamped.h>
namespace geometry_msgs { struct Twist { Vector3 linear, angular; }; struct TwistWithCovariance { Twist twist; std::array<double, 36> covariance{}; }; }
namespace nav_msgs { struct Odometry { std_msgs::Header header; std::string child_frame_id; geometry_msgs::PoseWithCovariance pose; geometry_msgs::TwistWithCovariance twist; }; }
