// tsdfloc_reduce.cuh — scan reduction on the device: the step immediately before the evaluation kernel.
//
// Replaces (does not port) the serial host code of TSDFEvaluator::evaluateParticles,
// src/evaluation/tsdf_evaluator.cpp:304-376 — an std::unordered_set<SortClass> with one std::shared_ptr per point,
// 64 per-ring std::vectors and a std::sort per ring — and its ring-agnostic cell-centre siblings
// (src/cuda/cuda_evaluator.cu:78-116, src/num_particles_eval.cpp:134-191):
//   1. drop points nearer than 1 m (:317-322) and, by policy, points with a non-finite coordinate;
//   2. key every surviving point by (ring, centre of its reduction cell) — fp32 `floor(x / res) * res + res/2`, every
//      operation rounded separately like the CPU build (:324-326) — and keep per key the FIRST point in cloud order
//      (unordered_set::insert keeps the element already present);
//   3. emit the ORIGINAL points ring by ring, inside a ring in cloud order (:340-376).
// The result is fully determined by the input (no dependence on hash-table iteration order), so it is reproduced
// bit for bit in parallel:
//   k_red_mark    survive flags, per-CTA counts                          }  the reference's running `index` and — in
//   k_red_scan    single-CTA exclusive scan (CTA counts / histogram)     }  RING_DESYNC mode — its ring-iterator bug
//   k_red_keys    running index, ring, cell centre -> key4[i]
//   k_red_insert  open-addressing table keyed on (ring, centre bits); slot value = atomicMin of the cloud positions
//   k_red_count   winner = table value == own position; stable in-CTA rank per ring (match.any + per-warp counts, no
//                 atomics -> deterministic); (ring x CTA) histogram
//   k_red_scan    ring-major exclusive scan of the histogram; total = number of output points
//   k_red_scatter out[offset(ring, CTA) + rank] = original point (or the cell centre in CENTRES mode)
// All kernels are bandwidth-trivial (16 B per point); the cost is launch latency, ~7 launches.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tsdfloc
{

constexpr int kRedThreads = 256;                 // points per CTA
constexpr int kRedWarps = kRedThreads / 32;
constexpr uint32_t kRedEmpty = 0xffffffffu;
constexpr uint32_t kRedMaxRings = 1024;

constexpr uint32_t kRedFlagDesync = 1u;          // TSDFLOC_REDUCE_RING_DESYNC_LIKE_REFERENCE
constexpr uint32_t kRedFlagCentres = 2u;         // TSDFLOC_REDUCE_EMIT_CENTRES

struct RedStatus
{
  uint32_t n_kept;    // points that survive the 1 m / finiteness test
  uint32_t n_out;     // points emitted
  uint32_t bad_ring;  // 1: a ring outside [0, n_rings) was seen
  uint32_t pad;
};

struct RedArgs
{
  const float* __restrict__ xyz;    // [n][3]
  const int32_t* __restrict__ ring; // [n] or nullptr (all ring 0)
  uint32_t n;
  uint32_t n_rings;
  uint32_t flags;
  uint32_t n_ctas;
  float res, half;                  // fp32 cell size and half of it (tsdf_evaluator.h:76-77)
  double res_d, half_d;             // CENTRES mode: the reference's double literals (cuda_evaluator.cu:100-102)
};

// Does point i survive? (tsdf_evaluator.cpp:311-322; the centre variants have no range test)
__device__ __forceinline__ bool red_survives(const RedArgs& A, uint32_t i, float& x, float& y, float& z)
{
  if (i >= A.n) return false;
  x = A.xyz[3ull * i];
  y = A.xyz[3ull * i + 1];
  z = A.xyz[3ull * i + 2];
  if (!(isfinite(x) && isfinite(y) && isfinite(z))) return false;
  if (A.flags & kRedFlagCentres) return true;
  const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  return !(static_cast<double>(dist) < 1.0);
}

__global__ void __launch_bounds__(kRedThreads) k_red_mark(const RedArgs A, uint32_t* __restrict__ cta_count)
{
  float x, y, z;
  const bool keep = red_survives(A, blockIdx.x * kRedThreads + threadIdx.x, x, y, z);
  const int c = __syncthreads_count(keep ? 1 : 0);
  if (threadIdx.x == 0) cta_count[blockIdx.x] = static_cast<uint32_t>(c);
}

// In-place exclusive scan of v[0..m) by ONE CTA of 1024 threads; *total = sum. Every thread owns a contiguous chunk of
// `per` elements (a multiple of 4, so chunks are 16 B aligned: v comes from cudaMalloc) and moves it with unrolled uint4
// accesses — all loads of a pass are in flight together instead of `per` dependent round trips to L2.
__global__ void __launch_bounds__(1024) k_red_scan(uint32_t* __restrict__ v, uint32_t m, uint32_t* __restrict__ total)
{
  __shared__ uint32_t warp_sum[32];
  const uint32_t t = threadIdx.x;
  const uint32_t per = ((m + 1023u) / 1024u + 3u) & ~3u;
  const uint32_t lo = min(t * per, m), hi = min(lo + per, m);
  const uint32_t full = lo + ((hi - lo) & ~3u);   // whole uint4 groups
  uint32_t s = 0;
#pragma unroll 8
  for (uint32_t i = lo; i < full; i += 4)
  {
    const uint4 q = *reinterpret_cast<const uint4*>(v + i);
    s += q.x + q.y + q.z + q.w;
  }
  for (uint32_t i = full; i < hi; ++i) s += v[i];
  // block-wide exclusive scan of s
  uint32_t incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if ((t & 31u) >= static_cast<uint32_t>(d)) incl += o;
  }
  if ((t & 31u) == 31u) warp_sum[t >> 5] = incl;
  __syncthreads();
  if (t < 32u)
  {
    uint32_t w = warp_sum[t];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const uint32_t o = __shfl_up_sync(0xffffffffu, w, d);
      if (t >= static_cast<uint32_t>(d)) w += o;
    }
    warp_sum[t] = w;  // inclusive over warps
  }
  __syncthreads();
  const uint32_t warp_off = (t >> 5) ? warp_sum[(t >> 5) - 1u] : 0u;
  uint32_t run = warp_off + incl - s;
#pragma unroll 8
  for (uint32_t i = lo; i < full; i += 4)
  {
    const uint4 q = *reinterpret_cast<const uint4*>(v + i);
    uint4 o;
    o.x = run;
    o.y = o.x + q.x;
    o.z = o.y + q.y;
    o.w = o.z + q.z;
    run = o.w + q.w;
    *reinterpret_cast<uint4*>(v + i) = o;
  }
  for (uint32_t i = full; i < hi; ++i)
  {
    const uint32_t x = v[i];
    v[i] = run;
    run += x;
  }
  if (t == 1023u && total) *total = warp_sum[31];
}

// key4[i] = (centre x, centre y, centre z, ring) as raw bits; ring = -1 marks a dropped point.
__global__ void __launch_bounds__(kRedThreads) k_red_keys(const RedArgs A, const uint32_t* __restrict__ cta_offset, int4* __restrict__ key4,
                                                         RedStatus* __restrict__ st)
{
  __shared__ uint32_t warp_cnt[kRedWarps];
  const uint32_t i = blockIdx.x * kRedThreads + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  float x = 0.f, y = 0.f, z = 0.f;
  const bool keep = red_survives(A, i, x, y, z);
  const uint32_t ball = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_cnt[warp] = __popc(ball);
  __syncthreads();
  uint32_t before = 0;
#pragma unroll
  for (int w = 0; w < kRedWarps; ++w) before += (static_cast<uint32_t>(w) < warp) ? warp_cnt[w] : 0u;
  if (i >= A.n) return;
  int4 k = make_int4(0, 0, 0, -1);
  if (keep)
  {
    // the reference's running `index` of surviving points (tsdf_evaluator.cpp:309,332)
    const uint32_t index = cta_offset[blockIdx.x] + before + __popc(ball & ((1u << lane) - 1u));
    // RING_DESYNC: `continue` skips `++iter_ring` (:319-322), so survivor #index reads the ring of cloud point #index
    int32_t r = 0;
    if (A.ring) r = A.ring[(A.flags & kRedFlagDesync) ? index : i];
    if (r < 0 || static_cast<uint32_t>(r) >= A.n_rings)
      atomicOr(&st->bad_ring, 1u);
    else
    {
      float cx, cy, cz;
      if (A.flags & kRedFlagCentres)
      {
        cx = static_cast<float>(__dadd_rn(__dmul_rn(floor(__ddiv_rn(static_cast<double>(x), A.res_d)), A.res_d), A.half_d));
        cy = static_cast<float>(__dadd_rn(__dmul_rn(floor(__ddiv_rn(static_cast<double>(y), A.res_d)), A.res_d), A.half_d));
        cz = static_cast<float>(__dadd_rn(__dmul_rn(floor(__ddiv_rn(static_cast<double>(z), A.res_d)), A.res_d), A.half_d));
      }
      else
      {
        cx = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(x, A.res)), A.res), A.half);
        cy = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(y, A.res)), A.res), A.half);
        cz = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(z, A.res)), A.res), A.half);
      }
      // SortClass::operator== compares floats (cuda_evaluator.h:56-59): +0 == -0, NaN != NaN. Canonicalise so that bit
      // equality means the same: -0 -> +0; a non-finite centre (overflow) drops the point like a non-finite coordinate.
      if (isfinite(cx) && isfinite(cy) && isfinite(cz))
        k = make_int4(__float_as_int(cx + 0.0f), __float_as_int(cy + 0.0f), __float_as_int(cz + 0.0f), r);
    }
  }
  key4[i] = k;
}

__device__ __forceinline__ uint32_t red_hash(const int4& k)
{
  uint32_t h = 0x9E3779B9u ^ static_cast<uint32_t>(k.w) * 0x85EBCA6Bu;
  h = (h ^ static_cast<uint32_t>(k.x)) * 0xC2B2AE35u;
  h ^= h >> 15;
  h = (h ^ static_cast<uint32_t>(k.y)) * 0x27D4EB2Fu;
  h ^= h >> 13;
  h = (h ^ static_cast<uint32_t>(k.z)) * 0x165667B1u;
  h ^= h >> 16;
  return h;
}

__device__ __forceinline__ bool red_same(const int4& a, const int4& b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }

// Slot value = smallest cloud position among the points sharing the slot's key. A slot's KEY never changes once claimed
// (atomicMin only swaps in positions with the same key), so comparing against whatever representative is read is sound.
__global__ void __launch_bounds__(kRedThreads) k_red_insert(const int4* __restrict__ key4, uint32_t n, uint32_t* __restrict__ table,
                                                           uint32_t mask)
{
  const uint32_t i = blockIdx.x * kRedThreads + threadIdx.x;
  if (i >= n) return;
  const int4 k = key4[i];
  if (k.w < 0) return;
  uint32_t slot = red_hash(k) & mask;
  for (;;)
  {
    uint32_t cur = *reinterpret_cast<volatile uint32_t*>(table + slot);
    if (cur == kRedEmpty)
    {
      cur = atomicCAS(table + slot, kRedEmpty, i);
      if (cur == kRedEmpty) return;
    }
    if (red_same(key4[cur], k))
    {
      atomicMin(table + slot, i);
      return;
    }
    slot = (slot + 1u) & mask;
  }
}

// Winners (first point of their key), their stable in-CTA rank among winners of the same ring, and the
// (ring x CTA) histogram. Dynamic shared memory: kRedWarps * n_rings uint16.
__global__ void __launch_bounds__(kRedThreads) k_red_count(const int4* __restrict__ key4, uint32_t n, const uint32_t* __restrict__ table,
                                                          uint32_t mask, uint32_t n_rings, uint32_t n_ctas, int32_t* __restrict__ win_ring,
                                                          uint32_t* __restrict__ local_rank, uint32_t* __restrict__ hist)
{
  extern __shared__ uint16_t cnt[];  // [kRedWarps][n_rings]
  for (uint32_t e = threadIdx.x; e < kRedWarps * n_rings; e += kRedThreads) cnt[e] = 0;
  __syncthreads();
  const uint32_t i = blockIdx.x * kRedThreads + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  int32_t ring = -1;
  if (i < n)
  {
    const int4 k = key4[i];
    if (k.w >= 0)
    {
      uint32_t slot = red_hash(k) & mask;
      for (;;)
      {
        const uint32_t cur = table[slot];
        if (cur == kRedEmpty) break;            // cannot happen: every surviving key was inserted
        if (red_same(key4[cur], k))
        {
          if (cur == i) ring = k.w;
          break;
        }
        slot = (slot + 1u) & mask;
      }
    }
  }
  const bool win = ring >= 0;
  const uint32_t act = __ballot_sync(0xffffffffu, win);
  uint32_t rank_in_warp = 0;
  if (win)
  {
    const uint32_t peers = __match_any_sync(act, ring);
    rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (rank_in_warp == 0) cnt[warp * n_rings + ring] = static_cast<uint16_t>(__popc(peers));
  }
  __syncthreads();
  if (i < n)
  {
    uint32_t before = 0;
    if (win)
      for (uint32_t w = 0; w < warp; ++w) before += cnt[w * n_rings + ring];
    win_ring[i] = ring;
    local_rank[i] = before + rank_in_warp;
  }
  for (uint32_t r = threadIdx.x; r < n_rings; r += kRedThreads)
  {
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < kRedWarps; ++w) s += cnt[w * n_rings + r];
    hist[static_cast<size_t>(r) * n_ctas + blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(kRedThreads) k_red_scatter(const RedArgs A, const int4* __restrict__ key4, const int32_t* __restrict__ win_ring,
                                                            const uint32_t* __restrict__ local_rank, const uint32_t* __restrict__ hist_offset,
                                                            float* __restrict__ out_xyz, uint32_t* __restrict__ out_src)
{
  const uint32_t i = blockIdx.x * kRedThreads + threadIdx.x;
  if (i >= A.n) return;
  const int32_t ring = win_ring[i];
  if (ring < 0) return;
  const uint32_t pos = hist_offset[static_cast<size_t>(ring) * A.n_ctas + blockIdx.x] + local_rank[i];
  float x, y, z;
  if (A.flags & kRedFlagCentres)
  {
    const int4 k = key4[i];
    x = __int_as_float(k.x);
    y = __int_as_float(k.y);
    z = __int_as_float(k.z);
  }
  else
  {
    x = A.xyz[3ull * i];
    y = A.xyz[3ull * i + 1];
    z = A.xyz[3ull * i + 2];
  }
  out_xyz[3ull * pos] = x;
  out_xyz[3ull * pos + 1] = y;
  out_xyz[3ull * pos + 2] = z;
  if (out_src) out_src[pos] = i;
}

}  // namespace tsdfloc
