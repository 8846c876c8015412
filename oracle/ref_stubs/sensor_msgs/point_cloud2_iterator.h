#pragma once
#include <stdexcept>
#include <sensor_msgs/PointCloud2.h>
namespace sensor_msgs {
// Walks one named field of a packed PointCloud2; it[k] reads the k-th T after the field offset.
template <typename T>
class PointCloud2ConstIterator {
  const uint8_t* cur_ = nullptr;
  const uint8_t* end_ = nullptr;
  uint32_t step_ = 0;
 public:
  PointCloud2ConstIterator() = default;
  PointCloud2ConstIterator(const PointCloud2& c, const std::string& field) {
    uint32_t off = 0; bool found = false;
    for (const auto& f : c.fields) if (f.name == field) { off = f.offset; found = true; }
    if (!found) throw std::runtime_error("Field " + field + " does not exist");
    step_ = c.point_step;
    cur_ = c.data.data() + off;
    end_ = c.data.data() + off + static_cast<size_t>(c.width) * c.height * c.point_step;
  }
  const T& operator[](size_t i) const { return *(reinterpret_cast<const T*>(cur_) + i); }
  const T& operator*() const { return *reinterpret_cast<const T*>(cur_); }
  PointCloud2ConstIterator& operator++() { cur_ += step_; return *this; }
  PointCloud2ConstIterator end() const { PointCloud2ConstIterator e; e.cur_ = end_; e.end_ = end_; e.step_ = step_; return e; }
  bool operator!=(const PointCloud2ConstIterator& o) const { return cur_ != o.cur_; }
};
}
