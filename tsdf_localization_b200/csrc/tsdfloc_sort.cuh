// tsdfloc_sort.cuh — spatial ordering of the particles in front of the evaluation kernel (no reference counterpart).
//
// When the map is larger than L2 (BASELINE configs C4/C5: 1.4 GB) and the particles are spread over it (global
// localisation), the particles that happen to be evaluated concurrently — ~7,000 of them — touch bricks all over the map and
// every voxel gather goes to HBM. Evaluating them in SPATIAL order instead makes the concurrently running CTAs share the
// bricks of one neighbourhood, which fit in L2. Each particle's weight is computed independently and bit-exactly, so the
// order of evaluation cannot change any result: the kernel reads matrix j = particle perm[j] and stores its weight to slot
// perm[j]. The permutation is a counting sort on a cell key (2-D Morton code of the 1 m column the particle stands in, z as
// the minor digit): histogram by atomicAdd, single-CTA exclusive scan (k_red_scan), scatter through atomic cursors. The
// order INSIDE a cell depends on atomic timing and is irrelevant.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tsdfloc
{

struct SortArgs
{
  float min[3];
  float inv_cell;      // 1 / cell edge (cells of 2^k metres so that the key space stays small)
  uint32_t dim[3];     // cells per axis
  uint32_t z_bits;
  uint32_t n_keys;     // key space size
};

__device__ __forceinline__ uint32_t part1by1(uint32_t v)
{
  v &= 0x0000ffffu;
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}

__device__ __forceinline__ uint32_t sort_key(const SortArgs& A, float x, float y, float z)
{
  // position -> cell, clamped into the map (particles outside still get a valid key; NaN compares false -> cell 0)
  const float fx = (x - A.min[0]) * A.inv_cell, fy = (y - A.min[1]) * A.inv_cell, fz = (z - A.min[2]) * A.inv_cell;
  const uint32_t cx = fx > 0.0f ? min(static_cast<uint32_t>(fx), A.dim[0] - 1u) : 0u;
  const uint32_t cy = fy > 0.0f ? min(static_cast<uint32_t>(fy), A.dim[1] - 1u) : 0u;
  const uint32_t cz = fz > 0.0f ? min(static_cast<uint32_t>(fz), A.dim[2] - 1u) : 0u;
  const uint32_t key = ((part1by1(cx) | (part1by1(cy) << 1)) << A.z_bits) | cz;
  return min(key, A.n_keys - 1u);
}

__global__ void __launch_bounds__(256) k_sort_keys(const float* __restrict__ particles, uint32_t first, uint32_t count, const SortArgs A,
                                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ hist)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float* p = particles + 7ull * (first + i);
  const uint32_t k = sort_key(A, p[0], p[1], p[2]);
  keys[i] = k;
  atomicAdd(hist + k, 1u);
}

// hist holds the exclusive prefix sums (first output slot of every key) and is used up as the per-key cursor.
__global__ void __launch_bounds__(256) k_sort_scatter(const uint32_t* __restrict__ keys, uint32_t count, uint32_t* __restrict__ hist,
                                                     uint32_t* __restrict__ perm)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  perm[atomicAdd(hist + keys[i], 1u)] = i;
}

}  // namespace tsdfloc
