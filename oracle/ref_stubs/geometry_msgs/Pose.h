#pragma once
#include <geometry_msgs/Point.h>
#include <geometry_msgs/Quaternion.h>
namespace geometry_msgs { struct Pose { Point position; Quaternion orientation; }; }
