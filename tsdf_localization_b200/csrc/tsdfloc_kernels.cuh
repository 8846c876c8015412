// tsdfloc_kernels.cuh — kernels of the B200 sensor update around the evaluation kernel (K1, tsdfloc_eval.cuh): scan
// preparation, K0 pose->matrix, K2 normalise + moments + CDF, K3 U-table, K4 draw. Included once by tsdfloc_api.cu. sm_100a only.
//
// Reference functions these replace (paths relative to the reference repo):
//   K0/K1  cudaEvaluateParticlesOrdered + cudaEvaluatePose   include/tsdf_localization/cuda/cuda_eval_particles.h:167-215, 273-333
//   K2     weightSum / chunk_sums_kernel / weight_particles   src/cuda/cuda_sum.cu:18-173, cuda_eval_particles.h:521-557
//   K3/K4  SystematicResampler::resample (CPU, serial)        include/tsdf_localization/resampling/novel_resampling.h:41-72
#pragma once
#include "tsdfloc_device.cuh"
#include "tsdfloc_eval.cuh"
#include <cstring>

namespace tsdfloc
{

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;  // particles per scan tile
constexpr int kMaxUSegs = 192;

// Device-side status block, read back by tsdfloc_check.
struct Status
{
  double weight_sum;      // sum of raw weights (fp64, fixed order)
  double s_last;          // last CDF entry (sum of normalised weights, fp64 running sum)
  float weight_sum_f;     // (float)weight_sum: the divisor of the normalisation
  uint32_t zero_sum;      // 1: weight_sum == 0 -> "No particle is valid!"
  uint32_t inexact;       // 1: a parallel fp64 add was inexact -> sequential fallback ran
  uint32_t n_segs;
  unsigned long long n_out;  // particles the reference recurrence emits
  uint32_t ticket;        // last-block-done counter of k_weight_sum (self-resetting)
  uint32_t table_overflow;   // build_u_table flags: 1 table full, 2 stalled recurrence, 4 max_j reached
  unsigned long long best_key;  // arg-max of the weights: (weight bits << 32) | ~index; 0 = no particle with weight > 0
  float best_pose[6];        // pose of that particle
  float best_weight;
  uint32_t pad;
};

// Arg-max key of mcl_3d.cpp:382-395 (`if (value > max_value)` from max_value = 0: the FIRST particle carrying the largest
// weight > 0): a larger weight wins, among equal weights the smaller index; weights <= 0 and NaN never win.
__device__ __forceinline__ unsigned long long best_key_of(float w, uint32_t i)
{
  if (!(w > 0.0f)) return 0ull;
  return (static_cast<unsigned long long>(__float_as_uint(w)) << 32) | static_cast<unsigned long long>(~i);
}

// One linear run of the reference's fp32 U recurrence: U_j = u_start + (j - j0) * step, exact, for j0 <= j < next j0.
struct USeg
{
  unsigned long long j0;
  float u_start;
  float step;
};

// ------------------------------------------------------------------------------------------------------------
// Measurement probe: the chip's ceiling for dependent-free 4 B gathers out of an L2-resident array, in the
// two access shapes that bracket the evaluation kernel — every lane its own random sector (32 sectors per request, what the
// reference's one-particle-per-thread kernel does) and the 32 lanes of a request spread over `spread` consecutive sectors
// around a random base (12.5 sectors per request is what k_eval2 measures). 8 independent gathers per thread per round,
// ld.global.nc like the kernel's table/voxel loads. Not part of the product path.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_probe_gather(const float* __restrict__ data, uint32_t n_words, uint32_t rounds, uint32_t spread_sectors,
                                                      float* __restrict__ sink)
{
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u, warp_id = tid >> 5;
  uint32_t state = (spread_sectors ? warp_id : tid) * 747796405u + 2891336453u;
  float acc = 0.0f;
  for (uint32_t r = 0; r < rounds; ++r)
  {
    uint32_t idx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
      state = state * 1664525u + 1013904223u;   // LCG; warp-uniform when the lanes share a neighbourhood
      const uint32_t h = (state >> 8) ^ (state << 7);
      if (spread_sectors)
      {
        // lanes land in `spread` consecutive sectors (8 words each) behind a random base
        const uint32_t base = (h % (n_words - spread_sectors * 8u)) & ~7u;
        idx[k] = base + ((lane * spread_sectors) >> 5) * 8u + (lane & 7u);
      }
      else
        idx[k] = h % n_words;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += __ldg(data + idx[k]);
  }
  if (acc == 123456.789f) sink[0] = acc;   // keep the loads alive
}

// ------------------------------------------------------------------------------------------------------------
// Scan preparation: xyz -> float4 with w = the point's range term
//   |p|^2 < max_range^2 ? a_range * (1/max_range) : a_max     (tsdf_evaluator.cpp:56-65, cuda_eval_particles.h:200-209)
// ------------------------------------------------------------------------------------------------------------
__global__ void k_prep_scan(const float* __restrict__ xyz, uint32_t p, float4* __restrict__ out, float a_range_term, float a_max,
                            float max_range_sq)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p; i += gridDim.x * blockDim.x)
  {
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const float sq = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    out[i] = make_float4(x, y, z, sq < max_range_sq ? a_range_term : a_max);
  }
}

// ------------------------------------------------------------------------------------------------------------
// K0: pose (x y z roll pitch yaw) o tf -> 3x4 sensor->map matrix, rounded exactly like the reference CPU build:
// sin/cos in fp64 rounded to fp32, every product/sum separately rounded (tsdf_evaluator.cpp:102-145).
// ------------------------------------------------------------------------------------------------------------
struct Tf12
{
  float m[12];
};

// perm (optional): matrix i belongs to particle first + perm[i] (spatial evaluation order, tsdfloc_sort.cuh).
__global__ void k_pose_matrices(const float* __restrict__ particles, uint32_t first, uint32_t count, Tf12 tf, float* __restrict__ mats,
                                const uint32_t* __restrict__ perm)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float* p = particles + 7ull * (first + (perm ? perm[i] : i));
  double sd, cd;
  sincos(static_cast<double>(p[3]), &sd, &cd);
  const float sa = static_cast<float>(sd), ca = static_cast<float>(cd);
  sincos(static_cast<double>(p[4]), &sd, &cd);
  const float sb = static_cast<float>(sd), cb = static_cast<float>(cd);
  sincos(static_cast<double>(p[5]), &sd, &cd);
  const float sg = static_cast<float>(sd), cg = static_cast<float>(cd);

  float r[3][4];
  r[0][0] = __fmul_rn(cb, cg);
  r[1][0] = __fmul_rn(cb, sg);
  r[2][0] = -sb;
  r[0][3] = p[0];
  r[0][1] = __fsub_rn(__fmul_rn(__fmul_rn(sa, sb), cg), __fmul_rn(ca, sg));
  r[1][1] = __fadd_rn(__fmul_rn(__fmul_rn(sa, sb), sg), __fmul_rn(ca, cg));
  r[2][1] = __fmul_rn(sa, cb);
  r[1][3] = p[1];
  r[0][2] = __fadd_rn(__fmul_rn(__fmul_rn(ca, sb), cg), __fmul_rn(sa, sg));
  r[1][2] = __fsub_rn(__fmul_rn(__fmul_rn(ca, sb), sg), __fmul_rn(sa, cg));
  r[2][2] = __fmul_rn(ca, cb);
  r[2][3] = p[2];

  float* o = mats + 12ull * i;
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    const float a = r[k][0], b = r[k][1], c = r[k][2], d = r[k][3];
    o[4 * k + 0] = __fadd_rn(__fadd_rn(__fmul_rn(a, tf.m[0]), __fmul_rn(b, tf.m[4])), __fmul_rn(c, tf.m[8]));
    o[4 * k + 1] = __fadd_rn(__fadd_rn(__fmul_rn(a, tf.m[1]), __fmul_rn(b, tf.m[5])), __fmul_rn(c, tf.m[9]));
    o[4 * k + 2] = __fadd_rn(__fadd_rn(__fmul_rn(a, tf.m[2]), __fmul_rn(b, tf.m[6])), __fmul_rn(c, tf.m[10]));
    o[4 * k + 3] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, tf.m[3]), __fmul_rn(b, tf.m[7])), __fmul_rn(c, tf.m[11])), d);
  }
}

// Exhaustive proof obligations of the sub-voxel quotient for ONE resolution, over every float a in [0, 1):
//   out[0]  a where the 3-instruction quotient's floor differs from floor(fl(a / res))            (0 = kDivThree is exact)
//   out[1]  a where floor(a * lo1) <= floor(fl(a / res)) <= floor(a * hi1) is VIOLATED             (0 = the bracket holds)
//   out[2]  a where that bracket is open (the two floors differ: the kernel redoes such a block exactly)
//   out[3], out[4]  the same for the wider pair (lo2, hi2)
__global__ void k_check_div(MapDev M, float lo1, float hi1, float lo2, float hi2, unsigned long long* __restrict__ out)
{
  unsigned long long bad3 = 0, bad1 = 0, open1 = 0, bad2 = 0, open2 = 0;
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < 0x3F800000u; b += gridDim.x * blockDim.x)
  {
    const float a = __uint_as_float(b);
    const uint32_t ieee = __float_as_uint(sub_coord<kDivIeee>(M, a));
    bad3 += (__float_as_uint(sub_coord<kDivThree>(M, a)) != ieee);
    const uint32_t l1 = __float_as_uint(__fmaf_rd(a, lo1, kMagic)), h1 = __float_as_uint(__fmaf_rd(a, hi1, kMagic));
    const uint32_t l2 = __float_as_uint(__fmaf_rd(a, lo2, kMagic)), h2 = __float_as_uint(__fmaf_rd(a, hi2, kMagic));
    bad1 += !(l1 <= ieee && ieee <= h1);
    open1 += (l1 != h1);
    bad2 += !(l2 <= ieee && ieee <= h2);
    open2 += (l2 != h2);
  }
  if (bad3) atomicAdd(out + 0, bad3);
  if (bad1) atomicAdd(out + 1, bad1);
  if (open1) atomicAdd(out + 2, open1);
  if (bad2) atomicAdd(out + 3, bad2);
  if (open2) atomicAdd(out + 4, open2);
}

// ------------------------------------------------------------------------------------------------------------
// K2a: sum of raw weights in fp64, fixed order (block tree, then the last block adds the block sums in index order).
// ------------------------------------------------------------------------------------------------------------
__global__ void k_weight_sum(const float* __restrict__ raw, uint32_t stride, uint32_t n, double* __restrict__ block_sums,
                             Status* __restrict__ st)
{
  __shared__ double s_part[kScanThreads / 32];
  __shared__ bool s_last;
  const uint32_t base = blockIdx.x * kScanTile;
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
  {
    const uint32_t i = base + threadIdx.x * kScanItems + k;
    if (i < n) acc += static_cast<double>(raw[static_cast<size_t>(i) * stride]);
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double t = 0.0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += s_part[w];
    block_sums[blockIdx.x] = t;
    __threadfence();
    const uint32_t ticket = atomicAdd(&st->ticket, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0)
  {
    __threadfence();
    double t = 0.0;
    for (uint32_t b = 0; b < gridDim.x; ++b) t += reinterpret_cast<volatile double*>(block_sums)[b];
    st->weight_sum = t;
    st->weight_sum_f = static_cast<float>(t);
    st->zero_sum = (t == 0.0) ? 1u : 0u;
    st->inexact = 0u;
    st->table_overflow = 0u;
    st->ticket = 0u;
  }
}

// exact-add check: returns a+b and sets `bad` if the fp64 addition rounded.
__device__ __forceinline__ double add_checked(double a, double b, bool& bad)
{
  const double s = a + b;
  const double bb = s - a;
  const double err = (a - (s - bb)) + (b - bb);
  bad |= (err != 0.0);
  return s;
}

// ------------------------------------------------------------------------------------------------------------
// K2b: normalise (w /= (float)sum, cuda_eval_particles.h:556), per-tile fp64 inclusive scan of the normalised
// weights, per-tile weighted moments (x y z, sin/cos of the three angles; tsdf_evaluator.cpp:203-217).
// raw == nullptr: the particles already carry their weights (Resampler::resample on a weighted cloud): they are
// scanned as they are and not rewritten.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads) k_normalise_scan(float* __restrict__ particles, const float* __restrict__ raw, uint32_t n,
                                                                 Status* __restrict__ st, double* __restrict__ cdf,
                                                                 double* __restrict__ tile_total, double* __restrict__ tile_moments,
                                                                 unsigned long long* __restrict__ tile_best)
{
  __shared__ double s_warp[kScanThreads / 32];
  __shared__ double s_mom[kScanThreads / 32][9];
  __shared__ unsigned long long s_best[kScanThreads / 32];
  unsigned long long best = 0ull;
  const float inv_den = st->weight_sum_f;
  const bool dead = st->zero_sum != 0u;
  const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  bool bad = false;

  double loc[kScanItems];
  double mom[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) mom[q] = 0.0;
  double run = 0.0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
  {
    const uint32_t i = base + k;
    float w = 0.0f;
    if (i < n)
    {
      float* p = particles + 7ull * i;
      if (raw)
      {
        w = dead ? 0.0f : __fdiv_rn(raw[i], inv_den);
        p[6] = w;
      }
      else
      {
        w = p[6];
      }
      const unsigned long long key = best_key_of(w, i);
      best = key > best ? key : best;
      const float x = p[0], y = p[1], z = p[2];
      mom[0] += static_cast<double>(__fmul_rn(x, w));
      mom[1] += static_cast<double>(__fmul_rn(y, w));
      mom[2] += static_cast<double>(__fmul_rn(z, w));
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        double sd, cd;
        sincos(static_cast<double>(p[3 + a]), &sd, &cd);
        mom[3 + 2 * a] += sd * static_cast<double>(w);
        mom[4 + 2 * a] += cd * static_cast<double>(w);
      }
    }
    run = add_checked(run, static_cast<double>(w), bad);
    loc[k] = run;
  }

  // inclusive scan of the per-thread totals across the warp, then across warps
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  double incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const double up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= static_cast<uint32_t>(o)) incl = add_checked(incl, up, bad);
  }
  double lane_excl = __shfl_up_sync(0xffffffffu, incl, 1);  // exclusive prefix over lanes: an already-checked sum
  if (lane == 0) lane_excl = 0.0;
  if (lane == 31) s_warp[warp] = incl;
#pragma unroll
  for (int q = 0; q < 9; ++q)
  {
    double v = mom[q];
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_mom[warp][q] = v;
  }
  for (int o = 16; o; o >>= 1)
  {
    const unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if (lane == 0) s_best[warp] = best;
  __syncthreads();
  double warp_off = 0.0;
  for (uint32_t w = 0; w < warp; ++w) warp_off = add_checked(warp_off, s_warp[w], bad);
  const double excl = add_checked(warp_off, lane_excl, bad);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
  {
    const uint32_t i = base + k;
    if (i < n) cdf[i] = add_checked(excl, loc[k], bad);
  }
  if (threadIdx.x == kScanThreads - 1) tile_total[blockIdx.x] = add_checked(excl, run, bad);
  if (threadIdx.x < 9)
  {
    double v = 0.0;
    for (int w = 0; w < kScanThreads / 32; ++w) v += s_mom[w][threadIdx.x];
    tile_moments[static_cast<size_t>(blockIdx.x) * 9 + threadIdx.x] = v;
  }
  if (threadIdx.x == 9)
  {
    unsigned long long b = 0ull;
    for (int w = 0; w < kScanThreads / 32; ++w) b = s_best[w] > b ? s_best[w] : b;
    tile_best[blockIdx.x] = b;
  }
  if (bad) atomicOr(&st->inexact, 1u);
}

// K2c: one warp. Exclusive scan of the tile totals (fp64, exactness-checked: when every addition is exact the result does
// not depend on the order, so the warp scans 32 tiles at a time instead of one thread walking them; an inexact addition
// raises st->inexact and K3 redoes the CDF sequentially), moments -> mean pose, per-tile arg-max -> best particle.
__global__ void __launch_bounds__(32) k_scan_tiles(const double* __restrict__ tile_total, double* __restrict__ tile_offset, uint32_t n_tiles,
                                                   const double* __restrict__ tile_moments, float* __restrict__ mean_pose,
                                                   Status* __restrict__ st, const unsigned long long* __restrict__ tile_best,
                                                   const float* __restrict__ particles)
{
  const uint32_t lane = threadIdx.x;
  bool bad = false;
  double carry = 0.0;
  double mom[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) mom[q] = 0.0;
  unsigned long long best = 0ull;
  for (uint32_t base = 0; base < n_tiles; base += 32u)
  {
    const uint32_t t = base + lane;
    const bool live = t < n_tiles;
    const double v = live ? tile_total[t] : 0.0;
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const double up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= static_cast<uint32_t>(o)) incl = add_checked(incl, up, bad);
    }
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0.0;
    if (live) tile_offset[t] = add_checked(carry, excl, bad);
    carry = add_checked(carry, __shfl_sync(0xffffffffu, incl, 31), bad);
    if (live)
    {
      // per-lane partial sums in a fixed (tile-index) order; combined below in a fixed tree order -> deterministic
#pragma unroll
      for (int q = 0; q < 9; ++q) mom[q] += tile_moments[static_cast<size_t>(t) * 9 + q];
      const unsigned long long b = tile_best[t];
      best = b > best ? b : best;
    }
  }
#pragma unroll
  for (int q = 0; q < 9; ++q)
    for (int o = 16; o; o >>= 1) mom[q] += __shfl_down_sync(0xffffffffu, mom[q], o);
  for (int o = 16; o; o >>= 1)
  {
    const unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&st->inexact, 1u);
  if (lane == 0)
  {
    if (mean_pose)
    {
      mean_pose[0] = static_cast<float>(mom[0]);
      mean_pose[1] = static_cast<float>(mom[1]);
      mean_pose[2] = static_cast<float>(mom[2]);
      mean_pose[3] = static_cast<float>(atan2(mom[3], mom[4]));
      mean_pose[4] = static_cast<float>(atan2(mom[5], mom[6]));
      mean_pose[5] = static_cast<float>(atan2(mom[7], mom[8]));
    }
    st->best_key = best;
    if (best)
    {
      const uint32_t i = ~static_cast<uint32_t>(best & 0xffffffffull);
      for (int k = 0; k < 6; ++k) st->best_pose[k] = particles[7ull * i + k];
      st->best_weight = particles[7ull * i + 6];
    }
  }
}

// K2d: add the tile offsets (exactness-checked) -> global fp64 CDF s_m = sum_{i<=m} w_i.
__global__ void k_cdf_finalize(double* __restrict__ cdf, const double* __restrict__ tile_offset, uint32_t n, Status* __restrict__ st)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool bad = false;
  const uint32_t tile = i / kScanTile;
  if (tile) cdf[i] = add_checked(tile_offset[tile], cdf[i], bad);
  if (bad) atomicOr(&st->inexact, 1u);
}

// ------------------------------------------------------------------------------------------------------------
// The reference's U recurrence: U_{j+1} = (float)((double)U_j + 1/N)  (novel_resampling.h:61-64, float += double).
// Inside one binade the rounded step is constant after the first element (DESIGN.md "U recurrence"), so the whole
// sequence is a short list of exact linear runs. Built by one thread; unit-tested on the host against the loop.
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t f32_bits(float f)
{
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, sizeof(u));
  return u;
#endif
}

__host__ __device__ inline float u_next(float u, double inv_m)
{
  return static_cast<float>(static_cast<double>(u) + inv_m);
}

// Walks the recurrence from U_0 = u0 and records it as segments until U_j >= limit (so every j with U_j < limit is
// covered), max_j elements were produced, the recurrence stalls (U stops growing), or max_segs is exhausted.
// Returns the number of segments. *n_below = #{j : U_j < limit} = the reference's output length for limit = s_last.
// *flags: bit0 = table full, bit1 = stalled recurrence, bit2 = max_j reached.
//
// Why linear runs are exact: write U = k*u with u = ulp of U's binade, k in [2^23, 2^24). While U + 1/N stays inside
// the binade, (double)U + 1/N rounds (fp64) to U + d' with d' independent of k, and the fp32 rounding of that adds
// q or q+1 ulps depending only on d' — except on an exact tie, where round-half-even makes k even after one step and
// the step constant from then on. So within a binade the step is constant from the second element on; a run is
// opened only where two consecutive measured steps agree, and it stops early enough (k + q <= 2^24 - 2) that every
// transition it covers stays inside the binade.
__host__ __device__ inline uint32_t build_u_table(float u0, double inv_m, double limit, USeg* segs, uint32_t max_segs,
                                                  unsigned long long* n_below, unsigned long long max_j, uint32_t* flags)
{
  uint32_t ns = 0;
  unsigned long long j = 0, below = 0;
  float u = u0;
  uint32_t fl = 0;
  while (static_cast<double>(u) < limit)
  {
    if (j >= max_j) { fl |= 4u; break; }
    if (ns >= max_segs) { fl |= 1u; break; }
    const float n1 = u_next(u, inv_m);
    if (!(n1 > u)) { fl |= 2u; ++below; break; }
    const uint32_t b0 = f32_bits(u);
    const uint32_t e0 = (b0 >> 23) & 0xffu;
    if (u > 0.0f && e0 != 0u && e0 != 0xffu)
    {
      const float n2 = u_next(n1, inv_m);
      const uint32_t b1 = f32_bits(n1), b2 = f32_bits(n2);
      if (((b1 >> 23) & 0xffu) == e0 && ((b2 >> 23) & 0xffu) == e0)
      {
        const uint32_t k0 = (b0 & 0x7fffffu) | 0x800000u;
        const uint32_t q1 = b1 - b0, q2 = b2 - b1;  // same binade: bit patterns differ by the ulp count
        if (q1 == q2 && q1 != 0u && k0 + q1 <= 0xfffffeu)
        {
          unsigned long long len = (0xfffffeull - k0) / q1;  // elements i = 0..len-1; element len is still in-binade
          if (len >= 2)
          {
            if (j + len > max_j) len = max_j - j;
            const double start = static_cast<double>(u);
            const double step = static_cast<double>(n1) - start;  // exact
            unsigned long long cnt = len;
            if (!(start + static_cast<double>(len - 1) * step < limit))
            {
              const double di = (limit - start) / step;
              unsigned long long i = di > 0.0 ? static_cast<unsigned long long>(di) : 0ull;
              if (i > len) i = len;
              while (i > 0 && !(start + static_cast<double>(i - 1) * step < limit)) --i;
              while (i < len && (start + static_cast<double>(i) * step < limit)) ++i;
              cnt = i;
            }
            segs[ns].j0 = j;
            segs[ns].u_start = u;
            segs[ns].step = static_cast<float>(step);
            ++ns;
            below += cnt;
            j += len;
            if (cnt < len) { *n_below = below; *flags = fl; return ns; }
            u = static_cast<float>(start + static_cast<double>(len) * step);
            continue;
          }
        }
      }
    }
    segs[ns].j0 = j;
    segs[ns].u_start = u;
    segs[ns].step = 0.0f;
    ++ns;
    ++below;
    ++j;
    u = n1;
  }
  *n_below = below;
  *flags = fl;
  return ns;
}

// U_j from the table (exact).
__host__ __device__ inline float u_at(const USeg* segs, uint32_t n_segs, unsigned long long j)
{
  uint32_t lo = 0, hi = n_segs;  // last segment with j0 <= j
  while (hi - lo > 1)
  {
    const uint32_t mid = (lo + hi) >> 1;
    if (segs[mid].j0 <= j) lo = mid; else hi = mid;
  }
  const USeg s = segs[lo];
  return static_cast<float>(static_cast<double>(s.u_start) + static_cast<double>(j - s.j0) * static_cast<double>(s.step));
}

// K3: one warp. If the parallel scan was inexact, thread 0 redoes the CDF sequentially (the reference's own
// order of fp64 additions); then builds the U table and n_out.
__global__ void k_finish_cdf_utable(const float* __restrict__ particles, double* __restrict__ cdf, uint32_t n, float u0,
                                    USeg* __restrict__ segs, Status* __restrict__ st)
{
  if (threadIdx.x != 0) return;
  if (st->inexact)
  {
    double s = 0.0;
    for (uint32_t i = 0; i < n; ++i)
    {
      s += static_cast<double>(particles[7ull * i + 6]);
      cdf[i] = s;
    }
  }
  const double s_last = n ? cdf[n - 1] : 0.0;
  st->s_last = s_last;
  unsigned long long n_below = 0;
  const double inv_m = 1.0 / static_cast<double>(n);
  // the recurrence cannot emit more than ~ s_last * N + 1 elements; 2N + 64 bounds any sane weight vector
  const unsigned long long max_j = 2ull * n + 64ull;
  uint32_t flags = 0;
  const uint32_t ns = build_u_table(u0, inv_m, s_last, segs, kMaxUSegs, &n_below, max_j, &flags);
  st->n_segs = ns;
  st->n_out = st->zero_sum ? 0ull : n_below;
  st->table_overflow = flags;
}

// Multi-GPU: the same output slice inside every other rank's particle buffer (peer-mapped); n == 0 on one GPU.
struct DrawPeers
{
  float* out[kMaxPeers];
  uint32_t n;
};

// K4: draw. Output slot j copies the first particle m with s_m > U_j (strict, novel_resampling.h:59).
__global__ void k_draw(const float* __restrict__ particles, const double* __restrict__ cdf, uint32_t n, const USeg* __restrict__ segs,
                       const Status* __restrict__ st, unsigned long long first_out, uint32_t count_out, float* __restrict__ out,
                       uint32_t* __restrict__ parents, const DrawPeers peers)
{
  __shared__ USeg s_segs[kMaxUSegs];
  const uint32_t ns = st->n_segs;
  for (uint32_t i = threadIdx.x; i < ns; i += blockDim.x) s_segs[i] = segs[i];
  __syncthreads();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count_out) return;
  const unsigned long long n_out = st->n_out;
  unsigned long long j = first_out + t;
  uint32_t parent = 0;
  if (n_out != 0ull && ns != 0u)
  {
    if (j >= n_out) j = n_out - 1;  // padding slots repeat the last valid draw
    const double u = static_cast<double>(u_at(s_segs, ns, j));
    uint32_t lo = 0, hi = n;  // first m in [0, n) with cdf[m] > u
    while (lo < hi)
    {
      const uint32_t mid = (lo + hi) >> 1;
      if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    parent = lo < n ? lo : n - 1;
  }
  const float* src = particles + 7ull * parent;
  float* dst = out + 7ull * t;
  float v[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) v[k] = src[k];
#pragma unroll
  for (int k = 0; k < 7; ++k) dst[k] = v[k];
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r)
    if (static_cast<uint32_t>(r) < peers.n)
    {
      float* pd = peers.out[r] + 7ull * t;
#pragma unroll
      for (int k = 0; k < 7; ++k) pd[k] = v[k];
    }
  if (parents) parents[t] = parent;
}

// ------------------------------------------------------------------------------------------------------------
// Run expansion — the device half of the Residual / ResidualSystematic resamplers (novel_resampling.h:27-30, 95-98: the
// `new_particles.push_back(particle)` loops). The host recurrence (host_resample.cpp) decides how many copies of which
// particle follow each other; run r covers output slots [run_off[r], run_off[r + 1]) and copies particle run_parent[r]
// (run_parent == nullptr: particle r itself). One thread per output slot: binary search of its run, 28 B copy.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_expand_runs(const float* __restrict__ particles, const uint32_t* __restrict__ run_off, const uint32_t* __restrict__ run_parent,
                              uint32_t n_runs, unsigned long long first_out, uint32_t count_out, float* __restrict__ out,
                              uint32_t* __restrict__ parents, const DrawPeers peers)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count_out) return;
  const unsigned long long total = run_off[n_runs];
  unsigned long long j = first_out + t;
  uint32_t parent = 0;
  if (total != 0ull)
  {
    if (j >= total) j = total - 1;  // padding slots repeat the last valid copy
    uint32_t lo = 0, hi = n_runs;   // last run r in [0, n_runs) with run_off[r] <= j (empty runs share an offset with their successor)
    while (hi - lo > 1)
    {
      const uint32_t mid = (lo + hi) >> 1;
      if (run_off[mid] <= j) lo = mid; else hi = mid;
    }
    parent = run_parent ? run_parent[lo] : lo;
  }
  const float* src = particles + 7ull * parent;
  float* dst = out + 7ull * t;
  float v[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) v[k] = src[k];
#pragma unroll
  for (int k = 0; k < 7; ++k) dst[k] = v[k];
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r)
    if (static_cast<uint32_t>(r) < peers.n)
    {
      float* pd = peers.out[r] + 7ull * t;
#pragma unroll
      for (int k = 0; k < 7; ++k) pd[k] = v[k];
    }
  if (parents) parents[t] = parent;
}

// weights (slot 6 of the AoS particles) -> a contiguous vector, for the host recurrences
__global__ void k_pack_weights(const float* __restrict__ particles, uint32_t n, float* __restrict__ w)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] = particles[7ull * i + 6];
}

}  // namespace tsdfloc
