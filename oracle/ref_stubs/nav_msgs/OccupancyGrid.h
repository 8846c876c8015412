#pragma once
#include <vector>
#include <geometry_msgs/Pose.h>
#include <geometry_msgs/TransformStamped.h>
namespace nav_msgs {
struct MapMetaData { float resolution = 0; unsigned width = 0, height = 0; geometry_msgs::Pose origin; };
struct OccupancyGrid { std_msgs::Header header; MapMetaData info; std::vector<signed char> data; };
}
