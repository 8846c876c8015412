#pragma once
#include <string>
#include <ros/ros.h>
#include <tf2_geometry_msgs/tf2_geometry_msgs.h>
namespace tf2_ros {
// No TF tree exists outside ROS: every lookup fails the way an unknown frame would.
class Buffer {
 public:
  geometry_msgs::TransformStamped lookupTransform(const std::string& target, const std::string& source, const ros::Time&, const ros::Duration&) const {
    throw tf2::TransformException("stub tf2_ros::Buffer: no transform " + source + " -> " + target);
  }
};
class TransformListener { public: explicit TransformListener(Buffer&) {} };
}
