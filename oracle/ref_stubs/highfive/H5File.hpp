// Stub of <highfive/H5File.hpp> (HighFive 2.2.2 is vendored by the reference but needs libhdf5, absent here).
// TEST INFRASTRUCTURE ONLY: an in-memory "file" so that the reference's createTSDFMap (map/map_util.h:17-154) compiles
// and runs verbatim. A file is a map group-name -> dataset-name -> uint32 payload; listObjectNames() returns the names in
// increasing (strcmp) order, which is the order HDF5's default name index yields.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
namespace HighFive {
using StubGroupData = std::map<std::string, std::vector<uint32_t>>;
inline std::map<std::string, std::map<std::string, StubGroupData>>& stub_files() {
  static std::map<std::string, std::map<std::string, StubGroupData>> files;
  return files;
}
class DataSet {
  const std::vector<uint32_t>* d_;
 public:
  explicit DataSet(const std::vector<uint32_t>* d) : d_(d) {}
  template <typename T> void read(std::vector<T>& out) const { out.assign(d_->begin(), d_->end()); }
};
class Group {
  const StubGroupData* g_;
 public:
  explicit Group(const StubGroupData* g) : g_(g) {}
  std::vector<std::string> listObjectNames() const {
    std::vector<std::string> names;
    for (const auto& kv : *g_) names.push_back(kv.first);
    return names;
  }
  DataSet getDataSet(const std::string& name) const {
    auto it = g_->find(name);
    if (it == g_->end()) throw std::runtime_error("no dataset " + name);
    return DataSet(&it->second);
  }
};
class File {
  const std::map<std::string, StubGroupData>* f_;
 public:
  enum : unsigned { ReadOnly = 0 };
  File(const std::string& name, unsigned) {
    auto it = stub_files().find(name);
    if (it == stub_files().end()) throw std::runtime_error("no such file " + name);
    f_ = &it->second;
  }
  Group getGroup(const std::string& name) const {
    auto it = f_->find(name);
    if (it == f_->end()) throw std::runtime_error("no group " + name);
    return Group(&it->second);
  }
};
}  // namespace HighFive
