#pragma once
// Stub (test infrastructure): the tf2 value types live in the tf2_geometry_msgs stub.
#include <tf2_geometry_msgs/tf2_geometry_msgs.h>
