#pragma once
#include <geometry_msgs/TransformStamped.h>
namespace tf2_ros { class TransformBroadcaster { public: void sendTransform(const geometry_msgs::TransformStamped&) {} }; }
