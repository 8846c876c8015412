#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <geometry_msgs/TransformStamped.h>
namespace sensor_msgs {
struct PointField { std::string name; uint32_t offset = 0; uint8_t datatype = 0; uint32_t count = 1; };
// Minimal PointCloud2: a packed byte buffer with named fields at byte offsets.
struct PointCloud2 {
  std_msgs::Header header;
  uint32_t height = 1, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = true;
};
}
