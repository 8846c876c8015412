// tsdfloc_motion.cuh — motion update on the device: the per-particle half of ParticleCloud::motionUpdate.
//
// Replaces (does not port) apply_model and the inlined copy of it in motionUpdate(odom)
// (src/particle_cloud.cpp:496-617, :219-327): six samples (dx dy dz roll pitch yaw) per particle build the odometry
// transform, it is composed with the particle's own rotation, the translation is added to the position and the new
// Euler angles are read back from the composed matrix (getAngleFromMat, src/util/util.cpp:86-111). The reference walks
// the particles serially on the host, drawing from six std::normal_distribution<> objects on one std::mt19937; here
// every particle is one thread and the samples come
//   * either from a caller-supplied array (n x 6 doubles) — parity mode: with the reference's own draws the result is the
//     reference's, bit for bit (sin/cos/asin/atan2 in fp64 rounded to fp32, every fp32 product and sum rounded separately),
//   * or from a counter-based Philox4x32-10 generator + Box-Muller (mean + sigma * z), keyed by (seed, sequence, particle):
//     reproducible for any launch geometry, no generator state in memory.
// Keeping the particles on the device between resampling and the next evaluation removes the last host round trip of the
// filter loop (src/mcl_3d.cpp:331-351).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tsdfloc
{

struct MotionArgs
{
  double mean[6];
  double sigma[6];
  unsigned long long seed;
  unsigned long long sequence;
};

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
}

// Philox4x32-10: 128-bit counter, 64-bit key -> 4 x 32 random bits.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
  for (int r = 0; r < 10; ++r)
  {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// two standard normal samples from 128 random bits (Box-Muller on two 53-bit uniforms in (0, 1))
__device__ __forceinline__ void normal_pair(const uint32_t (&r)[4], double& z0, double& z1)
{
  const unsigned long long a = (static_cast<unsigned long long>(r[0]) << 32) | r[1];
  const unsigned long long b = (static_cast<unsigned long long>(r[2]) << 32) | r[3];
  const double u1 = (static_cast<double>(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = (static_cast<double>(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

// R = Rz(c) Ry(b) Rx(a) rows from fp32 sines / cosines, every product and sum rounded separately, left to right
// (particle_cloud.cpp:525-539, :557-571)
__device__ __forceinline__ void rot_rows(float sa, float ca, float sb, float cb, float sc, float cc, float (&m)[12])
{
  m[0] = __fmul_rn(cb, cc);
  m[4] = __fmul_rn(cb, sc);
  m[8] = -sb;
  m[1] = __fsub_rn(__fmul_rn(__fmul_rn(sa, sb), cc), __fmul_rn(ca, sc));
  m[5] = __fadd_rn(__fmul_rn(__fmul_rn(sa, sb), sc), __fmul_rn(ca, cc));
  m[9] = __fmul_rn(sa, cb);
  m[2] = __fadd_rn(__fmul_rn(__fmul_rn(ca, sb), cc), __fmul_rn(sa, sc));
  m[6] = __fsub_rn(__fmul_rn(__fmul_rn(ca, sb), sc), __fmul_rn(sa, cc));
  m[10] = __fmul_rn(ca, cb);
}

__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2)
{
  return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// particles: n x 7 fp32, updated in place (weight slot untouched, particle_cloud.cpp:604). draws: n x 6 doubles or nullptr.
__global__ void __launch_bounds__(128) k_motion_apply(float* __restrict__ particles, uint32_t n, const double* __restrict__ draws,
                                                      const MotionArgs A)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d[6];
  if (draws)
  {
#pragma unroll
    for (int k = 0; k < 6; ++k) d[k] = draws[6ull * i + k];
  }
  else
  {
#pragma unroll
    for (int pair = 0; pair < 3; ++pair)
    {
      uint32_t c[4] = {i, static_cast<uint32_t>(pair), static_cast<uint32_t>(A.sequence), static_cast<uint32_t>(A.sequence >> 32)};
      philox4x32_10(c, static_cast<uint32_t>(A.seed), static_cast<uint32_t>(A.seed >> 32));
      double z0, z1;
      normal_pair(c, z0, z1);
      // std::normal_distribution returns z * stddev + mean
      d[2 * pair] = __dadd_rn(__dmul_rn(z0, A.sigma[2 * pair]), A.mean[2 * pair]);
      d[2 * pair + 1] = __dadd_rn(__dmul_rn(z1, A.sigma[2 * pair + 1]), A.mean[2 * pair + 1]);
    }
  }
  float* p = particles + 7ull * i;
  float o[12], q[12];
  {
    double s0, c0, s1, c1, s2, c2;
    sincos(d[3], &s0, &c0);
    sincos(d[4], &s1, &c1);
    sincos(d[5], &s2, &c2);
    rot_rows(static_cast<float>(s0), static_cast<float>(c0), static_cast<float>(s1), static_cast<float>(c1), static_cast<float>(s2),
             static_cast<float>(c2), o);
    o[3] = static_cast<float>(d[0]);
    o[7] = static_cast<float>(d[1]);
    o[11] = static_cast<float>(d[2]);
    sincos(static_cast<double>(p[3]), &s0, &c0);
    sincos(static_cast<double>(p[4]), &s1, &c1);
    sincos(static_cast<double>(p[5]), &s2, &c2);
    rot_rows(static_cast<float>(s0), static_cast<float>(c0), static_cast<float>(s1), static_cast<float>(c1), static_cast<float>(s2),
             static_cast<float>(c2), q);
  }
  // tf = tf_particle * tf_odom (:575-588); only the entries the pose read-back needs. tf_particle's translation is 0 (:560,565,570)
  const float t_x = __fadd_rn(dot3(q[0], o[3], q[1], o[7], q[2], o[11]), 0.0f);
  const float t_y = __fadd_rn(dot3(q[4], o[3], q[5], o[7], q[6], o[11]), 0.0f);
  const float t_z = __fadd_rn(dot3(q[8], o[3], q[9], o[7], q[10], o[11]), 0.0f);
  const float m0 = dot3(q[0], o[0], q[1], o[4], q[2], o[8]);
  const float m4 = dot3(q[4], o[0], q[5], o[4], q[6], o[8]);
  const float m8 = dot3(q[8], o[0], q[9], o[4], q[10], o[8]);
  const float m9 = dot3(q[8], o[1], q[9], o[5], q[10], o[9]);
  const float m10 = dot3(q[8], o[2], q[9], o[6], q[10], o[10]);
  // getAngleFromMat (util.cpp:86-111): fp64 libm calls on fp32 operands, results stored as fp32
  float roll, pitch, yaw;
  if (fabs(static_cast<double>(m8)) >= 1.0)
  {
    yaw = 0.0f;
    pitch = static_cast<float>(m8 < 0.0f ? 1.5707963267948966 : -1.5707963267948966);
    roll = static_cast<float>(atan2(static_cast<double>(m9), static_cast<double>(m10)));
  }
  else
  {
    pitch = static_cast<float>(-asin(static_cast<double>(m8)));
    const double cp = cos(static_cast<double>(pitch));
    roll = static_cast<float>(atan2(__ddiv_rn(static_cast<double>(m9), cp), __ddiv_rn(static_cast<double>(m10), cp)));
    yaw = static_cast<float>(atan2(__ddiv_rn(static_cast<double>(m4), cp), __ddiv_rn(static_cast<double>(m0), cp)));
  }
  p[0] = __fadd_rn(p[0], t_x);
  p[1] = __fadd_rn(p[1], t_y);
  p[2] = __fadd_rn(p[2], t_z);
  p[3] = roll;
  p[4] = pitch;
  p[5] = yaw;
}

// ------------------------------------------------------------------------------------------------------------
// Particle initialisation (ParticleCloud::initialize x 3, src/particle_cloud.cpp:32-148): every particle gets weight 1/n and
//   kInitNormal   x y z roll pitch yaw ~ N(mean[k], spread[k])                                   (:32-62)
//   kInitUniform  x y z roll pitch yaw ~ U(mean[k] - spread[k], mean[k] + spread[k])             (:64-103)
//   kInitFreeMap  xyz = a uniformly drawn free-space voxel (z - 0.5), angles ~ U(mean +- spread)  (:105-148)
// The double samples are rounded to fp32 when stored, like the reference's `first[k] = distribution(gen)`.
// Divergence: the reference draws the free-map index from [0, size] INCLUSIVE (:109) and reads one element past the end
// with probability 1/(size+1); here the index is uniform over [0, size).
// ------------------------------------------------------------------------------------------------------------
constexpr int kInitNormal = 0, kInitUniform = 1, kInitFreeMap = 2;

struct InitArgs
{
  double mean[6];
  double spread[6];
  unsigned long long seed;
  unsigned long long sequence;
  const float* __restrict__ free_map;  // [n_free][3] or nullptr
  uint32_t n_free;
  int mode;
};

__device__ __forceinline__ double uniform53(uint32_t hi, uint32_t lo)
{
  const unsigned long long a = (static_cast<unsigned long long>(hi) << 32) | lo;
  return static_cast<double>(a >> 11) * (1.0 / 9007199254740992.0);  // [0, 1), like std::generate_canonical<double, 53>
}

__global__ void __launch_bounds__(128) k_init_particles(float* __restrict__ particles, uint32_t n, const InitArgs A)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[6];
#pragma unroll
  for (int pair = 0; pair < 3; ++pair)
  {
    uint32_t c[4] = {i, 16u + static_cast<uint32_t>(pair), static_cast<uint32_t>(A.sequence), static_cast<uint32_t>(A.sequence >> 32)};
    philox4x32_10(c, static_cast<uint32_t>(A.seed), static_cast<uint32_t>(A.seed >> 32));
    if (A.mode == kInitNormal)
    {
      double z0, z1;
      normal_pair(c, z0, z1);
      v[2 * pair] = __dadd_rn(__dmul_rn(z0, A.spread[2 * pair]), A.mean[2 * pair]);
      v[2 * pair + 1] = __dadd_rn(__dmul_rn(z1, A.spread[2 * pair + 1]), A.mean[2 * pair + 1]);
    }
    else
    {
      // std::uniform_real_distribution: (b - a) * u + a
#pragma unroll
      for (int h = 0; h < 2; ++h)
      {
        const int k = 2 * pair + h;
        const double lo = A.mean[k] - A.spread[k], hi = A.mean[k] + A.spread[k];
        v[k] = __dadd_rn(__dmul_rn(hi - lo, uniform53(c[2 * h], c[2 * h + 1])), lo);
      }
    }
  }
  float* p = particles + 7ull * i;
  if (A.mode == kInitFreeMap && A.n_free)
  {
    uint32_t c[4] = {i, 32u, static_cast<uint32_t>(A.sequence), static_cast<uint32_t>(A.sequence >> 32)};
    philox4x32_10(c, static_cast<uint32_t>(A.seed), static_cast<uint32_t>(A.seed >> 32));
    const uint32_t idx = static_cast<uint32_t>((static_cast<unsigned long long>(c[0]) * A.n_free) >> 32);  // uniform over [0, n_free)
    p[0] = A.free_map[3ull * idx];
    p[1] = A.free_map[3ull * idx + 1];
    p[2] = static_cast<float>(static_cast<double>(A.free_map[3ull * idx + 2]) - 0.5);  // :139, float - double literal
  }
  else
  {
    p[0] = static_cast<float>(v[0]);
    p[1] = static_cast<float>(v[1]);
    p[2] = static_cast<float>(v[2]);
  }
  p[3] = static_cast<float>(v[3]);
  p[4] = static_cast<float>(v[4]);
  p[5] = static_cast<float>(v[5]);
  p[6] = static_cast<float>(1.0 / static_cast<double>(n));  // :44,56
}

}  // namespace tsdfloc
