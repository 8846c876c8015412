// host_map.cpp — host-side two-level sparse TSDF map builder and likelihood LUT of libtsdfloc.so.
//
// Stand-alone counterpart of the reference's CudaSubVoxelMap<float,float> constructor / setData
// (include/tsdf_localization/cuda/cuda_sub_voxel_map.tcc:4-48, 170-230) and of createTSDFMap's value transform
// (include/tsdf_localization/map/map_util.h:68-71, 124-126). It produces exactly the arrays the reference class
// would (1 m upper cells -> dense sub_dim^3 bricks, bricks laid out in increasing upper-cell index), because the flat
// voxel numbering of that layout is what parity is measured in. Unlike the reference it rejects cells outside the
// bounding box instead of invoking undefined float->unsigned conversions.
#include "../../include/tsdfloc.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <numeric>
#include <string>
#include <vector>

#include "tsdfloc_host_map.h"

namespace
{

// every division below must be a single IEEE fp32 operation, like the reference's
inline float div32(float a, float b)
{
  volatile float q = a / b;
  return q;
}

// upper cell / in-cell voxel of one axis offset, same arithmetic as cuda_sub_voxel_map.tcc:74-84,132-134
inline void split_axis(float off, float res, uint64_t& up, uint64_t& sub)
{
  up = static_cast<uint64_t>(off);  // off >= 0 checked by the caller
  volatile float pos = off - static_cast<float>(up);
  sub = static_cast<uint64_t>(div32(pos, res));
}

}  // namespace

namespace tsdfloc_host
{

// Empty host map with createTSDFMap's bounding box for these chunks (map_util.h:23-78) and the order in which the reference
// visits the datasets (increasing name = HDF5's default name index). Returns TSDFLOC_OK or an error (e.g. a chunk twice).
int begin_chunk_map(const int32_t* chunk_pos, uint64_t n_chunks, float sigma, tsdfloc_host_map** out, std::vector<uint64_t>& order)
{
  constexpr int kChunk = 64;          // CHUNK_SIZE (grid_map.h:18-19)
  constexpr int kResMm = 64;          // MAP_RESOLUTION in millimetres (grid_map.h:21-22)
  // bounding box over the chunk coordinates, starting from 0 like the reference's min(3, 0) / max(3, 0) (:23-59)
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  for (uint64_t c = 0; c < n_chunks; ++c)
    for (int a = 0; a < 3; ++a)
    {
      const float v = static_cast<float>(chunk_pos[3 * c + a]);
      if (v < lo[a]) lo[a] = v;
      if (v > hi[a]) hi[a] = v;
    }
  float mn[3], mx[3];
  for (int a = 0; a < 3; ++a)
  {
    // float * int * int stays fp32, the final * 0.001 is a double product stored back into the float (:61-65)
    volatile float t0 = lo[a] * kChunk;
    volatile float t1 = t0 * kResMm;
    mn[a] = static_cast<float>(t1 * 0.001);
    volatile float u0 = hi[a] * kChunk;
    volatile float u1 = u0 + kChunk;
    volatile float u2 = u1 * kResMm;
    mx[a] = static_cast<float>(u2 * 0.001);
  }
  const float res = static_cast<float>(kResMm * 0.001);
  tsdfloc_host_map* m = nullptr;
  int rc = tsdfloc_map_create(mn, mx, res, tsdfloc_likelihood_init(sigma), &m);
  if (rc != TSDFLOC_OK) return rc;

  // datasets are visited in increasing name order (HDF5's default name index), which fixes the order of the free points
  std::vector<std::string> tags(n_chunks);
  for (uint64_t c = 0; c < n_chunks; ++c)
    tags[c] = std::to_string(chunk_pos[3 * c]) + "_" + std::to_string(chunk_pos[3 * c + 1]) + "_" + std::to_string(chunk_pos[3 * c + 2]);
  order.resize(n_chunks);
  std::iota(order.begin(), order.end(), 0ull);
  std::sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) { return tags[x] < tags[y]; });
  for (uint64_t k = 1; k < n_chunks; ++k)
    if (tags[order[k]] == tags[order[k - 1]])
    {
      delete m;
      return TSDFLOC_E_BAD_ARG;  // the same chunk twice
    }

  *out = m;
  return TSDFLOC_OK;
}

// likelihood^3 of every representable in-band TSDF value, indexed by value_mm + 599
std::vector<float> likelihood_lut(float sigma)
{
  std::vector<float> lut(1199);
  for (int mm = -599; mm <= 599; ++mm) lut[mm + 599] = tsdfloc_likelihood_value(static_cast<float>(mm), sigma);
  return lut;
}

}  // namespace tsdfloc_host

extern "C"
{

int tsdfloc_map_create(const float mn[3], const float mx[3], float res, float init_value, tsdfloc_host_map** out)
{
  if (!out || !mn || !mx) return TSDFLOC_E_BAD_ARG;
  *out = nullptr;
  if (!(res > 0.0f) || !std::isfinite(res)) return TSDFLOC_E_BAD_ARG;
  tsdfloc_host_map* m = new (std::nothrow) tsdfloc_host_map;
  if (!m) return TSDFLOC_E_BAD_ARG;
  tsdfloc_map_desc& d = m->desc;
  for (int a = 0; a < 3; ++a)
  {
    if (!std::isfinite(mn[a]) || !std::isfinite(mx[a])) { delete m; return TSDFLOC_E_BAD_ARG; }
    volatile float extent = std::fabs(mx[a] - mn[a]);
    d.dim[a] = static_cast<uint64_t>(std::ceil(div32(extent, res)));
    d.up_dim[a] = static_cast<uint64_t>(std::ceil(extent));  // 1 m upper cells (cuda_sub_voxel_map.h:15)
    d.min[a] = mn[a];
    d.max[a] = mx[a];
  }
  d.resolution = res;
  d.init_value = init_value;
  d.up_dim_2 = d.up_dim[0] * d.up_dim[1];
  d.sub_dim = static_cast<uint64_t>(std::ceil(div32(1.0f, res)));
  d.sub_dim_2 = d.sub_dim * d.sub_dim;
  d.grid_occ_size = d.up_dim[0] * d.up_dim[1] * d.up_dim[2];
  d.data_size = 0;
  if (d.grid_occ_size == 0 || d.grid_occ_size >= (1ull << 31)) { delete m; return TSDFLOC_E_BAD_ARG; }
  m->grid_occ.assign(d.grid_occ_size, -1);
  *out = m;
  return TSDFLOC_OK;
}

int tsdfloc_map_set_data(tsdfloc_host_map* m, const float* cells, uint64_t n)
{
  if (!m || (n && !cells)) return TSDFLOC_E_BAD_ARG;
  if (n == 0) return TSDFLOC_OK;
  tsdfloc_map_desc& d = m->desc;
  std::fill(m->grid_occ.begin(), m->grid_occ.end(), -1);
  std::vector<uint64_t> upper(n), inner(n);
  for (uint64_t i = 0; i < n; ++i)
  {
    uint64_t up[3], sub[3];
    for (int a = 0; a < 3; ++a)
    {
      volatile float off = cells[4 * i + a] - d.min[a];
      if (!(off >= 0.0f)) return TSDFLOC_E_BAD_ARG;
      split_axis(off, d.resolution, up[a], sub[a]);
      if (up[a] >= d.up_dim[a]) return TSDFLOC_E_BAD_ARG;
    }
    upper[i] = up[0] + up[1] * d.up_dim[0] + up[2] * d.up_dim_2;
    inner[i] = sub[0] + sub[1] * d.sub_dim + sub[2] * d.sub_dim_2;
    m->grid_occ[upper[i]] = 0;
  }
  const uint64_t brick = d.sub_dim * d.sub_dim * d.sub_dim;
  uint64_t offset = 0;
  for (uint64_t u = 0; u < d.grid_occ_size; ++u)
  {
    if (m->grid_occ[u] < 0) continue;
    if (offset + brick >= (1ull << 31)) return TSDFLOC_E_BAD_ARG;  // reference offsets are int
    m->grid_occ[u] = static_cast<int32_t>(offset);
    offset += brick;
  }
  d.data_size = offset;
  m->data.assign(offset, d.init_value);
  for (uint64_t i = 0; i < n; ++i)
  {
    const uint64_t idx = static_cast<uint64_t>(m->grid_occ[upper[i]]) + inner[i];
    if (idx < offset) m->data[idx] = cells[4 * i + 3];
  }
  return TSDFLOC_OK;
}

const tsdfloc_map_desc* tsdfloc_map_get_desc(const tsdfloc_host_map* m) { return m ? &m->desc : nullptr; }
const int32_t* tsdfloc_map_grid_occ(const tsdfloc_host_map* m) { return m ? m->grid_occ.data() : nullptr; }
const float* tsdfloc_map_data(const tsdfloc_host_map* m) { return m ? m->data.data() : nullptr; }
void tsdfloc_map_destroy(tsdfloc_host_map* m) { delete m; }

// Map ingest: createTSDFMap without the HDF5 layer (include/tsdf_localization/map/map_util.h:17-154). The mapping
// pipeline stores the TSDF as 64^3-voxel chunks of packed {int16 value_mm, int16 weight} words (util/tsdf.h:11-87,
// map/grid_map.h:18-24) in datasets named "<cx>_<cy>_<cz>"; the caller reads them with whatever HDF5 binding it has and
// hands over the raw words.
int tsdfloc_map_from_chunks(const int32_t* chunk_pos, const uint32_t* chunk_data, uint64_t n_chunks, float sigma, tsdfloc_host_map** out)
{
  if (!out) return TSDFLOC_E_BAD_ARG;
  *out = nullptr;
  if (n_chunks && (!chunk_pos || !chunk_data)) return TSDFLOC_E_BAD_ARG;
  if (!(sigma > 0.0f)) return TSDFLOC_E_BAD_ARG;
  constexpr int kChunk = 64;          // CHUNK_SIZE (grid_map.h:18-19)
  constexpr int kResMm = 64;          // MAP_RESOLUTION in millimetres (grid_map.h:21-22)
  constexpr float kTruncation = 600;  // grid_map.h:24

  tsdfloc_host_map* m = nullptr;
  std::vector<uint64_t> order;
  int rc = tsdfloc_host::begin_chunk_map(chunk_pos, n_chunks, sigma, &m, order);
  if (rc != TSDFLOC_OK) return rc;
  const std::vector<float> lut = tsdfloc_host::likelihood_lut(sigma);

  std::vector<float> cells;
  const size_t words = static_cast<size_t>(kChunk) * kChunk * kChunk;
  for (uint64_t k = 0; k < n_chunks; ++k)
  {
    const uint64_t c = order[k];
    const uint32_t* w = chunk_data + c * words;
    const int bx = kChunk * chunk_pos[3 * c], by = kChunk * chunk_pos[3 * c + 1], bz = kChunk * chunk_pos[3 * c + 2];
    for (int i = 0; i < kChunk; ++i)
      for (int j = 0; j < kChunk; ++j)
        for (int q = 0; q < kChunk; ++q)
        {
          const uint32_t raw = w[(static_cast<size_t>(i) * kChunk + j) * kChunk + q];  // :106
          const int16_t value_mm = static_cast<int16_t>(raw & 0xffffu);                // TSDFValueHW::value (tsdf.h:16)
          const int16_t weight = static_cast<int16_t>(raw >> 16);                      // TSDFValueHW::weight
          if (weight == 0) continue;
          const float tsdf = static_cast<float>(value_mm);
          // voxel CORNER position: float(index) * 64 in fp32, * 0.001 in double, stored as fp32 (:118-120, :129)
          volatile float fx = static_cast<float>(bx + i) * kResMm, fy = static_cast<float>(by + j) * kResMm, fz = static_cast<float>(bz + q) * kResMm;
          const float px = static_cast<float>(fx * 0.001), py = static_cast<float>(fy * 0.001), pz = static_cast<float>(fz * 0.001);
          if (std::fabs(tsdf) < kTruncation)
          {
            cells.push_back(px);
            cells.push_back(py);
            cells.push_back(pz);
            cells.push_back(lut[value_mm + 599]);
          }
          else
          {
            m->free_points.push_back(px);
            m->free_points.push_back(py);
            m->free_points.push_back(pz);
          }
        }
  }
  rc = tsdfloc_map_set_data(m, cells.data(), cells.size() / 4);
  if (rc != TSDFLOC_OK)
  {
    delete m;
    return rc;
  }
  *out = m;
  return TSDFLOC_OK;
}

const float* tsdfloc_map_free_points(const tsdfloc_host_map* m, uint64_t* n)
{
  if (n) *n = m ? m->free_points.size() / 3 : 0;
  return m && !m->free_points.empty() ? m->free_points.data() : nullptr;
}

float tsdfloc_likelihood_init(float sigma)
{
  // N(10 m; 0, sigma)^3 — the value of unmapped space (0 in fp32 for sigma = 0.1)
  const float s2 = sigma * sigma;
  const float e = std::exp(static_cast<float>(-(10.0 * 10.0) / s2 / 2));
  const float nrm = std::sqrt(static_cast<float>(2 * s2 * 3.14159265358979323846));
  const float v = e / nrm;
  return v * v * v;
}

float tsdfloc_likelihood_value(float tsdf_mm, float sigma)
{
  // metres in double, exp/sqrt in fp32, the cube in double, stored as fp32 — the reference's mixed precision
  const float s2 = sigma * sigma;
  const double metres = tsdf_mm * 0.001;
  const float e = std::exp(static_cast<float>(-(metres * metres) / s2 / 2));
  const float nrm = std::sqrt(static_cast<float>(2 * s2 * 3.14159265358979323846));
  const double v = e / nrm;
  return static_cast<float>(v * v * v);
}

}  // extern "C"
