// tsdfloc_host_map.h — internal: the host map object behind the tsdfloc_map_* calls, shared by host_map.cpp and the GPU
// ingest in tsdfloc_api.cu. Not part of the public interface (include/tsdfloc.h only forward-declares the struct).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/tsdfloc.h"

struct tsdfloc_host_map
{
  tsdfloc_map_desc desc{};
  std::vector<int32_t> grid_occ;
  std::vector<float> data;
  std::vector<float> free_points;  // x y z of the free-space voxels createTSDFMap collects (map_util.h:131-145)
};

namespace tsdfloc_host
{
// Empty host map with createTSDFMap's bounding box for these chunks (map_util.h:23-78) + the order in which the reference
// visits the datasets (increasing name). TSDFLOC_OK or an error (e.g. a chunk twice).
int begin_chunk_map(const int32_t* chunk_pos, uint64_t n_chunks, float sigma, tsdfloc_host_map** out, std::vector<uint64_t>& order);
// likelihood^3 of every representable in-band TSDF value, indexed by value_mm + 599
std::vector<float> likelihood_lut(float sigma);
}  // namespace tsdfloc_host
