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
#include <cooperative_groups.h>
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
  uint32_t ticket;        // (unused)
  uint32_t table_overflow;   // build_u_table flags: 1 table full, 2 stalled recurrence, 4 max_j reached
  unsigned long long best_key;  // arg-max of the weights: (weight bits << 32) | ~index; 0 = no particle with weight > 0
  float best_pose[6];        // pose of that particle
  float best_weight;
  uint32_t pad;
  float mean[8];             // weighted mean pose when the caller did not supply its own device buffer: one read-back gets all
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
struct PrepArgs
{
  const float* __restrict__ xyz;   // [p][3]
  float4* __restrict__ out;        // [padded]
  uint32_t p, padded;              // points beyond p are written as zero points (the evaluation reads whole blocks)
  float a_range_term, a_max, max_range_sq;
};

__device__ __forceinline__ void prep_scan_points(const PrepArgs& A, uint32_t first_thread, uint32_t n_threads)
{
  for (uint32_t i = first_thread; i < A.padded; i += n_threads)
  {
    if (i < A.p)
    {
      const float x = A.xyz[3 * i], y = A.xyz[3 * i + 1], z = A.xyz[3 * i + 2];
      const float sq = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
      A.out[i] = make_float4(x, y, z, sq < A.max_range_sq ? A.a_range_term : A.a_max);
    }
    else
      A.out[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
}

__global__ void __launch_bounds__(256) k_prep_scan(const PrepArgs A)
{
  prep_scan_points(A, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
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
__device__ __forceinline__ void pose_matrix(const float* __restrict__ particles, uint32_t first, uint32_t i, const Tf12& tf, float* __restrict__ mats,
                                            const uint32_t* __restrict__ perm)
{
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

__global__ void __launch_bounds__(256) k_pose_matrices(const float* __restrict__ particles, uint32_t first, uint32_t count, Tf12 tf,
                                                      float* __restrict__ mats, const uint32_t* __restrict__ perm)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) pose_matrix(particles, first, i, tf, mats, perm);
}

// Scan preparation and pose matrices in ONE launch (the host-buffer update issues them together): CTAs [0, scan_ctas) prepare
// the scan, the rest build matrices.
__global__ void __launch_bounds__(256) k_prepare(const PrepArgs A, uint32_t scan_ctas, const float* __restrict__ particles, uint32_t first,
                                                uint32_t count, Tf12 tf, float* __restrict__ mats, const uint32_t* __restrict__ perm)
{
  if (blockIdx.x < scan_ctas)
    prep_scan_points(A, blockIdx.x * blockDim.x + threadIdx.x, scan_ctas * blockDim.x);
  else
  {
    const uint32_t i = (blockIdx.x - scan_ctas) * blockDim.x + threadIdx.x;
    if (i < count) pose_matrix(particles, first, i, tf, mats, perm);
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
// K2: ONE cooperative kernel (grid-wide syncs instead of kernel boundaries) for everything between the evaluation and the
// draw. Replaces weightSum x10, weight_particles, the host atan2s (cuda_evaluator.cu:364-392, cuda_sum.cu,
// cuda_eval_particles.h:521-557) and the arg-max loop of the node (mcl_3d.cpp:382-399):
//   A  per-tile fp64 sums of the raw weights (fixed order)                                  | grid sync
//   B  sum of the tile sums (fixed order, every CTA redundantly) -> the fp32 divisor; normalise (w /= (float)sum,
//      cuda_eval_particles.h:556), tile-local fp64 inclusive scan, weighted moments (tsdf_evaluator.cpp:203-217), arg-max
//      (raw == nullptr: the particles already carry their weights — Resampler::resample on a weighted cloud — they are scanned
//      as they are and nothing is rewritten)                                                | grid sync
//   C  tile offsets (exactness-checked fp64: when every addition is exact the order is irrelevant; one that rounds raises
//      st->inexact and k_cdf_fallback redoes the CDF in the reference's own order) -> global CDF s_m = sum_{i<=m} w_i;
//      CTA 0: moments -> mean pose, best particle.
// A CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the grid is sized to be co-resident (cooperative launch).
// ------------------------------------------------------------------------------------------------------------
struct NormArgs
{
  float* particles;               // [n][7]; slot 6 receives the normalised weight (raw != nullptr)
  const float* raw;               // [n] un-normalised weights, or nullptr
  uint32_t n;
  Status* st;
  double* cdf;                    // [n]
  double* tile_sum;               // [tiles] phase A
  double* tile_total;             // [tiles] phase B: sum of the tile's normalised weights
  double* tile_moments;           // [tiles][9]
  unsigned long long* tile_best;  // [tiles]
  float* mean_pose;               // [6] or nullptr
  float* w_out;                   // [n] or nullptr: the normalised weights once more as a contiguous vector (4 B per particle
                                  //     for the host write-back instead of the 28 B particles)
};

__device__ __forceinline__ double block_sum_fixed(double v, double* s_part)
{
  // fixed order: shuffle tree inside each warp, then thread 0 adds the warp sums in warp order
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kScanThreads / 32; ++w) t += s_part[w];
  __syncthreads();
  return t;   // valid in thread 0
}

// Grid-wide barrier of the two kernels below. kCoop = false: the launch is ONE CTA (at most 1,024 particles: C1) and an
// ordinary launch — a block barrier orders its global writes, and the cooperative launch's extra latency is saved.
template <bool kCoop>
__device__ __forceinline__ void grid_barrier()
{
  if (kCoop)
    cooperative_groups::this_grid().sync();
  else
    __syncthreads();
}

template <bool kCoop>
__global__ void __launch_bounds__(kScanThreads) k_normalise_cdf(const NormArgs A)
{
  __shared__ double s_part[kScanThreads / 32];
  __shared__ double s_mom[kScanThreads / 32][9];
  __shared__ unsigned long long s_best[kScanThreads / 32];
  __shared__ double s_bcast;
  const uint32_t n = A.n;
  const uint32_t n_tiles = (n + kScanTile - 1) / kScanTile;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t stride = A.raw ? 1u : 7u;
  const float* __restrict__ src = A.raw ? A.raw : A.particles + 6;

  // ---- A: tile sums ------------------------------------------------------------------------------------------
  if (blockIdx.x == 0 && threadIdx.x == 0) A.st->inexact = 0u;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    const uint32_t base = tile * kScanTile + threadIdx.x * kScanItems;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (base + k < n) acc += static_cast<double>(src[static_cast<size_t>(base + k) * stride]);
    const double t = block_sum_fixed(acc, s_part);
    if (threadIdx.x == 0) A.tile_sum[tile] = t;
  }
  grid_barrier<kCoop>();

  // ---- B: total (every CTA, same order), normalise + tile-local scan + moments + arg-max ------------------------
  {
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < n_tiles; i += kScanThreads) acc += A.tile_sum[i];
    const double t = block_sum_fixed(acc, s_part);
    if (threadIdx.x == 0) s_bcast = t;
    __syncthreads();
  }
  const double total = s_bcast;
  const float den = static_cast<float>(total);
  const bool dead = (total == 0.0);
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    A.st->weight_sum = total;
    A.st->weight_sum_f = den;
    A.st->zero_sum = dead ? 1u : 0u;
  }
  bool bad = false;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    const uint32_t base = tile * kScanTile + threadIdx.x * kScanItems;
    unsigned long long best = 0ull;
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
        float* p = A.particles + 7ull * i;
        if (A.raw)
        {
          w = dead ? 0.0f : __fdiv_rn(A.raw[i], den);
          p[6] = w;
        }
        else
          w = p[6];
        if (A.w_out) A.w_out[i] = w;
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
    double incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const double up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= static_cast<uint32_t>(o)) incl = add_checked(incl, up, bad);
    }
    double lane_excl = __shfl_up_sync(0xffffffffu, incl, 1);  // exclusive prefix over lanes: an already-checked sum
    if (lane == 0) lane_excl = 0.0;
    __syncthreads();   // s_part / s_mom / s_best of the previous tile have been consumed
    if (lane == 31) s_part[warp] = incl;
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
    for (uint32_t w = 0; w < warp; ++w) warp_off = add_checked(warp_off, s_part[w], bad);
    const double excl = add_checked(warp_off, lane_excl, bad);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (base + k < n) A.cdf[base + k] = add_checked(excl, loc[k], bad);
    if (threadIdx.x == kScanThreads - 1) A.tile_total[tile] = add_checked(excl, run, bad);
    if (threadIdx.x < 9)
    {
      double v = 0.0;
      for (int w = 0; w < kScanThreads / 32; ++w) v += s_mom[w][threadIdx.x];
      A.tile_moments[static_cast<size_t>(tile) * 9 + threadIdx.x] = v;
    }
    if (threadIdx.x == 9)
    {
      unsigned long long b = 0ull;
      for (int w = 0; w < kScanThreads / 32; ++w) b = s_best[w] > b ? s_best[w] : b;
      A.tile_best[tile] = b;
    }
  }
  grid_barrier<kCoop>();

  // ---- C: tile offsets -> global CDF; CTA 0: mean pose + best particle ----------------------------------------------
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    if (tile == 0) continue;
    // offset = sum of the totals of all earlier tiles, every addition checked for exactness
    double part = 0.0;
    for (uint32_t i = threadIdx.x; i < tile; i += kScanThreads) part = add_checked(part, A.tile_total[i], bad);
    for (int o = 16; o; o >>= 1)
    {
      const double other = __shfl_down_sync(0xffffffffu, part, o);
      part = add_checked(part, other, bad);
    }
    __syncthreads();
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0)
    {
      double t = 0.0;
      for (int w = 0; w < kScanThreads / 32; ++w) t = add_checked(t, s_part[w], bad);
      s_bcast = t;
    }
    __syncthreads();
    const double off = s_bcast;
    const uint32_t base = tile * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (base + k < n) A.cdf[base + k] = add_checked(off, A.cdf[base + k], bad);
  }
  if (bad) atomicOr(&A.st->inexact, 1u);

  if (blockIdx.x == 0 && warp == 0)
  {
    // per-lane partial sums in a fixed (tile-index) order, combined in a fixed tree order -> deterministic
    double mom[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) mom[q] = 0.0;
    unsigned long long best = 0ull;
    for (uint32_t t = lane; t < n_tiles; t += 32u)
    {
#pragma unroll
      for (int q = 0; q < 9; ++q) mom[q] += A.tile_moments[static_cast<size_t>(t) * 9 + q];
      const unsigned long long b = A.tile_best[t];
      best = b > best ? b : best;
    }
#pragma unroll
    for (int q = 0; q < 9; ++q)
      for (int o = 16; o; o >>= 1) mom[q] += __shfl_down_sync(0xffffffffu, mom[q], o);
    for (int o = 16; o; o >>= 1)
    {
      const unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0)
    {
      if (A.mean_pose)
      {
        A.mean_pose[0] = static_cast<float>(mom[0]);
        A.mean_pose[1] = static_cast<float>(mom[1]);
        A.mean_pose[2] = static_cast<float>(mom[2]);
        A.mean_pose[3] = static_cast<float>(atan2(mom[3], mom[4]));
        A.mean_pose[4] = static_cast<float>(atan2(mom[5], mom[6]));
        A.mean_pose[5] = static_cast<float>(atan2(mom[7], mom[8]));
      }
      A.st->best_key = best;
      if (best)
      {
        const uint32_t i = ~static_cast<uint32_t>(best & 0xffffffffull);
        for (int k = 0; k < 6; ++k) A.st->best_pose[k] = A.particles[7ull * i + k];
        A.st->best_weight = A.particles[7ull * i + 6];
      }
    }
  }
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

// The U table of one resampling call, built on the HOST (it depends only on u0 and N: build_u_table with no limit) and handed
// to the draw kernel BY VALUE as a kernel parameter: no device-side table builder, no copy to order against the launch.
struct UTable
{
  USeg segs[kMaxUSegs];
  unsigned long long n_elems;   // U_j is covered for j < n_elems
  uint32_t n_segs;
  uint32_t flags;               // build_u_table flags (1 table full, 2 stalled recurrence, 4 max_j reached)
};

// #{j < n_elems : U_j < limit}: U_j is non-decreasing in j, so a binary search over j with the exact u_at().
__device__ inline unsigned long long u_count_below(const USeg* segs, uint32_t n_segs, unsigned long long n_elems, double limit)
{
  unsigned long long lo = 0, hi = n_elems;   // first j with U_j >= limit
  while (lo < hi)
  {
    const unsigned long long mid = (lo + hi) >> 1;
    if (static_cast<double>(u_at(segs, n_segs, mid)) < limit) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------------------
// K3: the CDF in the reference's own order of additions — s += w_m, m = 0 .. n-1, in fp64 (novel_resampling.h:57) — for
// weight vectors whose fp64 sum ROUNDS (st->inexact: a parallel addition of K2 was inexact, so the order matters). With the
// normalised weights of one sensor update the additions are exact and every CTA returns at once; clouds whose weights span
// dozens of binades take this path (a serial loop costs 38 ms at 2^20 particles).
// The serial sum is reproduced in parallel. While the running sum s = S * u stays inside one binade (ulp u, S an integer)
// and the addends are >= 0, RN(s + w) adds RN(w / u) units — except for an exact TIE (w / u = k + 1/2), which
// round-half-even resolves by the parity of S + k: it adds k + ((S + k) mod 2) and leaves S EVEN. So the only state a weight
// needs from its predecessors is the parity of S, and every weight is a map {parity in} -> {units added, parity out}; those
// maps compose associatively, which makes the whole thing a parallel scan (struct ParityMap). Per 1,024-particle tile:
//   1  (all tiles in parallel) guess the tile's binade from K2's approximate scan, compose the tile's maps: units added and
//      parity out for both start parities; a tile with a negative / non-finite / huge addend or a binade change is "complex"
//   2  (one thread, tile by tile, exact) take the units for the parity of the exact s; accept s + units * u iff the exact s
//      really lies in the guessed binade and the result still does; otherwise the tile's additions are done one after the other
//      (s doubles at most ~80 times between the smallest fp32 weight and 1: a few dozen serial tiles at worst)
//   3  (accepted tiles in parallel) cdf[m] = s_tile + u * units(prefix up to m), exact
// ------------------------------------------------------------------------------------------------------------
constexpr uint32_t kTileComplex = 0u, kTileSimple = 1u, kTileZero = 2u;

struct ParityMap
{
  long long d0, d1;   // units added when the running sum's parity on entry is 0 / 1
  uint32_t o;         // bit 0 / bit 1: parity on exit for parity 0 / 1 on entry
};

__device__ __forceinline__ ParityMap pm_identity() { return ParityMap{0ll, 0ll, 2u}; }

// f first, then g
__device__ __forceinline__ ParityMap pm_then(const ParityMap& f, const ParityMap& g)
{
  const uint32_t f0 = f.o & 1u, f1 = (f.o >> 1) & 1u;
  ParityMap r;
  r.d0 = f.d0 + (f0 ? g.d1 : g.d0);
  r.d1 = f.d1 + (f1 ? g.d1 : g.d0);
  r.o = ((g.o >> f0) & 1u) | (((g.o >> f1) & 1u) << 1);
  return r;
}

// the map of one weight at ulp 1 / inv_u; `bad` flags what the scheme cannot express
__device__ __forceinline__ ParityMap pm_of_weight(float wf, double inv_u, bool& bad, bool& nonzero)
{
  nonzero |= (wf != 0.0f);
  const double q = static_cast<double>(wf) * inv_u;     // exact: a power-of-two scaling
  const bool ok = (q < 4503599627370496.0) && (wf >= 0.0f);   // < 2^52, finite, not NaN; negative addends break monotonicity
  bad |= !ok;
  if (!ok) return pm_identity();
  const double fl = floor(q);
  const long long k = __double2ll_rd(q);
  const uint32_t kp = static_cast<uint32_t>(k) & 1u;
  if (q - fl == 0.5)     // exact tie: to even
    return ParityMap{k + static_cast<long long>(kp), k + static_cast<long long>(kp ^ 1u), 0u};
  const long long r = __double2ll_rn(q);
  const uint32_t x = static_cast<uint32_t>(r) & 1u;
  return ParityMap{r, r, x | ((x ^ 1u) << 1)};
}

__device__ __forceinline__ ParityMap pm_shfl_up(const ParityMap& m, int delta)
{
  ParityMap r;
  r.d0 = __shfl_up_sync(0xffffffffu, m.d0, delta);
  r.d1 = __shfl_up_sync(0xffffffffu, m.d1, delta);
  r.o = __shfl_up_sync(0xffffffffu, m.o, delta);
  return r;
}

struct ExactArgs
{
  const float* particles;
  uint32_t n;
  const Status* st;
  double* cdf;            // in: K2's scan (approximate where additions rounded); out: the serial-order CDF
  long long* tile_units;  // [tiles][2] units the tile adds for start parity 0 / 1, at the tile's guessed ulp
  uint32_t* tile_flag;    // [tiles] kind | exit parities << 4 | biased exponent of the guessed binade << 8
  double* tile_start;     // [tiles] exact running sum in front of the tile
};

__device__ __forceinline__ uint32_t f64_exp(double x) { return static_cast<uint32_t>(__double_as_longlong(x) >> 52) & 0x7ffu; }

// Inclusive scan of the tile's maps: returns the composition of all maps of the items BEFORE this thread's first item
// (exclusive prefix) and leaves the thread's own per-item inclusive compositions (relative to that prefix) in loc[].
__device__ __forceinline__ ParityMap tile_scan(const ExactArgs& A, uint32_t base, double inv_u, ParityMap (&loc)[kScanItems], bool& bad,
                                               bool& nonzero, ParityMap* s_warp, ParityMap& tile_total)
{
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  ParityMap run = pm_identity();
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
  {
    if (base + k < A.n) run = pm_then(run, pm_of_weight(A.particles[7ull * (base + k) + 6], inv_u, bad, nonzero));
    loc[k] = run;
  }
  ParityMap incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const ParityMap up = pm_shfl_up(incl, o);
    if (lane >= static_cast<uint32_t>(o)) incl = pm_then(up, incl);
  }
  ParityMap excl = pm_shfl_up(incl, 1);
  if (lane == 0) excl = pm_identity();
  __syncthreads();
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  ParityMap off = pm_identity();
  for (uint32_t w = 0; w < warp; ++w) off = pm_then(off, s_warp[w]);
  tile_total = off;
  for (uint32_t w = warp; w < kScanThreads / 32; ++w) tile_total = pm_then(tile_total, s_warp[w]);
  return pm_then(off, excl);
}

template <bool kCoop>
__global__ void __launch_bounds__(kScanThreads) k_cdf_exact(const ExactArgs A)
{
  if (!A.st->inexact) return;   // uniform over the grid: nobody reaches a grid sync
  __shared__ ParityMap s_warp[kScanThreads / 32];
  __shared__ uint32_t s_flags;
  const uint32_t n = A.n;
  const uint32_t n_tiles = (n + kScanTile - 1) / kScanTile;

  // ---- 1: the tile's map at the guessed ulp ---------------------------------------------------------------------------
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    const uint32_t first = tile * kScanTile, last = min(first + kScanTile, n) - 1u;
    const double s_lo = tile ? A.cdf[first - 1u] : 0.0, s_hi = A.cdf[last];
    const uint32_t e = f64_exp(s_lo);
    const bool cand = s_lo > 0.0 && e == f64_exp(s_hi) && e >= 64u && e < 0x7ffu;
    const double inv_u = __longlong_as_double(static_cast<long long>(2098u - min(max(e, 64u), 2046u)) << 52);   // 2^(52 - (e - 1023))
    bool bad = false, nonzero = false;
    ParityMap loc[kScanItems], total;
    if (threadIdx.x == 0) s_flags = 0u;
    tile_scan(A, first + threadIdx.x * kScanItems, inv_u, loc, bad, nonzero, s_warp, total);
    if (bad) atomicOr(&s_flags, 1u);
    if (nonzero) atomicOr(&s_flags, 2u);
    __syncthreads();
    if (threadIdx.x == 0)
    {
      const uint32_t kind = !(s_flags & 2u) ? kTileZero : (cand && !(s_flags & 1u)) ? kTileSimple : kTileComplex;
      A.tile_units[2ull * tile] = total.d0;
      A.tile_units[2ull * tile + 1] = total.d1;
      A.tile_flag[tile] = kind | ((total.o & 3u) << 4) | (e << 8);
    }
    __syncthreads();
  }
  grid_barrier<kCoop>();

  // ---- 2: the exact running sum in front of every tile (CTA 0) ---------------------------------------------------------
  // Thread 0 walks the tiles; their flags / unit counts are staged in shared memory a chunk at a time so that the walk is a
  // chain of arithmetic, not of L2 round trips. A tile that needs its additions done one after the other is staged through
  // shared memory by the whole CTA (coalesced loads and stores around thread 0's 1,024 dependent additions).
  if (blockIdx.x == 0)
  {
    constexpr uint32_t kChunk = 512;
    __shared__ uint32_t s_flag[kChunk];
    __shared__ long long s_u0[kChunk], s_u1[kChunk];
    __shared__ float s_w[kScanTile];
    __shared__ double s_c[kScanTile];
    __shared__ double s_run;
    __shared__ uint32_t s_stop;     // next tile to be done serially (or end of chunk)
    if (threadIdx.x == 0) s_run = 0.0;
    for (uint32_t chunk0 = 0; chunk0 < n_tiles; chunk0 += kChunk)
    {
      const uint32_t chunk_n = min(kChunk, n_tiles - chunk0);
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < chunk_n; i += kScanThreads)
      {
        s_flag[i] = A.tile_flag[chunk0 + i];
        s_u0[i] = A.tile_units[2ull * (chunk0 + i)];
        s_u1[i] = A.tile_units[2ull * (chunk0 + i) + 1];
      }
      __syncthreads();
      uint32_t pos = 0;   // uniform: every thread follows s_stop
      while (pos < chunk_n)
      {
        if (threadIdx.x == 0)
        {
          double s = s_run;
          uint32_t i = pos;
          for (; i < chunk_n; ++i)
          {
            A.tile_start[chunk0 + i] = s;
            const uint32_t flag = s_flag[i], kind = flag & 0xfu, e = flag >> 8;
            if (kind == kTileZero) continue;
            bool ok = false;
            if (kind == kTileSimple && f64_exp(s) == e)
            {
              const uint32_t parity = static_cast<uint32_t>(__double_as_longlong(s)) & 1u;   // S = s / u: the mantissa's last bit
              const long long T = parity ? s_u1[i] : s_u0[i];
              if (T < (1ll << 53))
              {
                const double u = __longlong_as_double(static_cast<long long>(e - 52u) << 52);
                const double next = s + static_cast<double>(T) * u;   // exact when it stays inside the binade
                if (f64_exp(next) == e)
                {
                  s = next;
                  ok = true;
                }
              }
            }
            if (!ok) break;
          }
          s_run = s;
          s_stop = i;
        }
        __syncthreads();
        const uint32_t stop = s_stop;
        if (stop >= chunk_n) break;
        // tile chunk0 + stop: serial additions
        const uint32_t tile = chunk0 + stop, first = tile * kScanTile, cnt = min(static_cast<uint32_t>(kScanTile), n - first);
        for (uint32_t i = threadIdx.x; i < cnt; i += kScanThreads) s_w[i] = A.particles[7ull * (first + i) + 6];
        __syncthreads();
        if (threadIdx.x == 0)
        {
          A.tile_flag[tile] = kTileComplex | (s_flag[stop] & ~0xffu);
          double s = s_run;
          for (uint32_t i = 0; i < cnt; ++i)
          {
            s += static_cast<double>(s_w[i]);
            s_c[i] = s;
          }
          s_run = s;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += kScanThreads) A.cdf[first + i] = s_c[i];
        pos = stop + 1;
        __syncthreads();
      }
    }
  }
  grid_barrier<kCoop>();

  // ---- 3: accepted tiles: start + u * units(prefix) -------------------------------------------------------------------
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    const uint32_t flag = A.tile_flag[tile], kind = flag & 0xfu, e = flag >> 8;
    if (kind == kTileComplex) continue;
    const double start = A.tile_start[tile];
    const uint32_t base = tile * kScanTile + threadIdx.x * kScanItems;
    if (kind == kTileZero)
    {
#pragma unroll
      for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) A.cdf[base + k] = start;
      continue;
    }
    const double inv_u = __longlong_as_double(static_cast<long long>(2098u - e) << 52);
    const double u = __longlong_as_double(static_cast<long long>(e - 52u) << 52);
    const uint32_t parity = static_cast<uint32_t>(__double_as_longlong(start)) & 1u;
    bool bad = false, nonzero = false;
    ParityMap loc[kScanItems], total;
    const ParityMap excl = tile_scan(A, base, inv_u, loc, bad, nonzero, s_warp, total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (base + k < n)
      {
        const ParityMap m = pm_then(excl, loc[k]);
        A.cdf[base + k] = start + static_cast<double>(parity ? m.d1 : m.d0) * u;
      }
    __syncthreads();
  }
}

// Multi-GPU: the same output slice inside every other rank's particle buffer (peer-mapped); n == 0 on one GPU.
struct DrawPeers
{
  float* out[kMaxPeers];
  uint32_t n;
};

// K4: draw. Output slot j copies the first particle m with s_m > U_j (strict, novel_resampling.h:59). Every CTA derives the
// output length n_out = #{j : U_j < s_last} itself (the `while (s > U)` loop of the reference ends there); CTA 0 publishes it.
__global__ void __launch_bounds__(256) k_draw(const float* __restrict__ particles, const double* __restrict__ cdf, uint32_t n,
                                              const __grid_constant__ UTable T, Status* __restrict__ st, unsigned long long first_out,
                                              uint32_t count_out, float* __restrict__ out, uint32_t* __restrict__ parents, const DrawPeers peers)
{
  __shared__ USeg s_segs[kMaxUSegs];
  __shared__ unsigned long long s_n_out;
  const uint32_t ns = T.n_segs;
  for (uint32_t i = threadIdx.x; i < ns; i += blockDim.x) s_segs[i] = T.segs[i];
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const double s_last = n ? cdf[n - 1] : 0.0;
    const unsigned long long below = (ns && !st->zero_sum) ? u_count_below(s_segs, ns, T.n_elems, s_last) : 0ull;
    s_n_out = below;
    if (blockIdx.x == 0)
    {
      st->s_last = s_last;
      st->n_segs = ns;
      st->n_out = below;
      // the table ended before U reached s_last: it was full, the recurrence stalled, or the weights sum to more than 2
      st->table_overflow = (below >= T.n_elems && ns && !st->zero_sum) ? (T.flags ? T.flags : 4u) : 0u;
    }
  }
  __syncthreads();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count_out) return;
  const unsigned long long n_out = s_n_out;
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
