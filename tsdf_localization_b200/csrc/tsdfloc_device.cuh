// tsdfloc_device.cuh — device-side data layout and the voxel index function of the B200 sensor update.
//
// Replaces (does not port) include/tsdf_localization/cuda/cuda_eval_particles.h:12-164 (getIndex/getEntry) of the
// reference. The index ARITHMETIC must be bit-identical to the reference's for every non-negative offset
// (SURVEY §2.5(3)); the way it is evaluated is new:
//   * the metre-truncated bound test and the `up_index >= grid_occ_size` test are folded, at map-upload time,
//     into a PADDED BRICK TABLE with a one-cell border of "miss" entries, so the kernel clamps each axis
//     offset into the padded range instead of branching;
//   * misses (border, unallocated upper cells, negative/NaN offsets — policy NEG_AS_MISS) point at a
//     MISS BRICK filled with init_value appended behind the real bricks: the gather is branch-free;
//   * floor() is a round-down add of 2^23 (FADD.RM) instead of F2I/I2F;
//   * floor(fl(x / resolution)) — the reference divides in fp32 and truncates — is BRACKETED instead of computed: two
//     round-down FMAs  floor(x * inv_lo) <= floor(fl(x / res)) <= floor(x * inv_hi)  with inv_lo < 1/res < inv_hi one or two
//     ulps apart. Where both floors agree (all but ~4e-6 of the quotients) that IS the reference's value; where they differ
//     the evaluation kernel redoes the summation block with the exact division. tsdfloc_create proves the bracket for the
//     map's resolution EXHAUSTIVELY (every float in [0, 1), k_check_div) and otherwise falls back to a verified
//     3-instruction correctly-rounded sequence (FMUL, FFMA, FFMA) or to __fdiv_rn.
//   * negative offsets: policy MISS (default; clamp into the padded border) or SATURATE (clamp to 0: the reference CUDA
//     build's saturating f32->u32 conversions, cuda_eval_particles.h:14-16,34-36,62-64) — one kernel parameter, no branch.
// All fp32 arithmetic that feeds an index uses explicit _rn/_rd intrinsics: nvcc never contracts those into FMAs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tsdfloc
{

constexpr float kMagic = 8388608.0f;        // 2^23
constexpr float kMagicP1 = 8388609.0f;      // 2^23 + 1: floor(x)+1 lands in the mantissa
constexpr uint32_t kMagicBits = 0x4B000000u;  // bit pattern of 2^23

// Device view of the map. 96 bytes, passed by value as a kernel parameter (constant bank).
struct MapDev
{
  const int32_t* __restrict__ table;  // padded brick table [pz][py][px]: element offset of the brick in voxels
  const float* __restrict__ voxels;   // bricks in reference order + one miss brick at miss_offset
  float min[3];
  float clamp_hi[3];   // (float)thr[a]: offsets are clamped into [clamp_lo, thr]
  float clamp_lo;      // -1: negative offsets land in the padded border (miss); 0: they saturate onto the min face (reference CUDA build)
  float res;
  float inv_res;       // RN(1/res)
  float inv_lo, inv_hi;  // bracket of 1/res for the round-down quotient floors (div_mode == kDivBracket)
  uint32_t pad_x;      // x stride of the padded table: pow2 >= thr[0] + 2
  uint32_t pad_xy;     // z stride: pad_x * (pow2 >= thr[1] + 2)
  uint32_t shift_x;    // log2(pad_x)
  uint32_t shift_xy;   // log2(pad_xy)
  uint32_t sub_dim;
  uint32_t sub_dim_2;
  uint32_t data_size;  // reference data_size; every index >= data_size is a miss
  uint32_t table_bias; // kMagicBits * (1 + pad_x + pad_xy)   (mod 2^32)
  uint32_t sub_bias;   // kMagicBits * (1 + sub_dim + sub_dim_2) (mod 2^32)
  int32_t div_mode;    // kDivIeee / kDivThree / kDivBracket: what k_check_div proved for this resolution
};

constexpr int kDivIeee = 0;      // __fdiv_rn
constexpr int kDivThree = 1;     // verified 3-instruction correctly-rounded quotient
constexpr int kDivBracket = 2;   // two round-down FMAs; disagreement -> the caller redoes the block with an exact mode

// 2^23 + floor(fl(a / res)) for a in [0, 1), as a float whose low mantissa bits are the sub-voxel coordinate.
// kDiv = kDivThree is bit-identical to IEEE a / res when k_check_div verified it for this resolution.
template <int kDiv>
__device__ __forceinline__ float sub_coord(const MapDev& M, float a)
{
  static_assert(kDiv == kDivIeee || kDiv == kDivThree, "the scalar path has no bracket mode");
  if (kDiv == kDivThree)
  {
    const float q0 = __fmul_rn(a, M.inv_res);
    const float e = __fmaf_rn(-M.res, q0, a);
    return __fadd_rd(__fmaf_rn(e, M.inv_res, q0), kMagic);
  }
  return __fadd_rd(__fdiv_rn(a, M.res), kMagic);
}

// Flat voxel offset (reference numbering: brick offset + sx + sy*sub_dim + sz*sub_dim^2) of world point t.
// Returns a value >= M.data_size (inside the miss brick) for every miss.
template <int kDiv>
__device__ __forceinline__ uint32_t voxel_index(const MapDev& M, float tx, float ty, float tz)
{
  // offsets x - min (cuda_eval_particles.h:27-29), clamped into the padded table's range
  const float ox = fminf(fmaxf(__fsub_rn(tx, M.min[0]), M.clamp_lo), M.clamp_hi[0]);
  const float oy = fminf(fmaxf(__fsub_rn(ty, M.min[1]), M.clamp_lo), M.clamp_hi[1]);
  const float oz = fminf(fmaxf(__fsub_rn(tz, M.min[2]), M.clamp_lo), M.clamp_hi[2]);
  // 2^23 + 1 + floor(o): the low mantissa bits are the padded upper-cell coordinate (:34-36)
  const float bx = __fadd_rd(ox, kMagicP1);
  const float by = __fadd_rd(oy, kMagicP1);
  const float bz = __fadd_rd(oz, kMagicP1);
  // position inside the 1 m upper cell (:40-42); exact
  const float px = __fsub_rn(ox, __fsub_rn(bx, kMagicP1));
  const float py = __fsub_rn(oy, __fsub_rn(by, kMagicP1));
  const float pz = __fsub_rn(oz, __fsub_rn(bz, kMagicP1));
  // padded table index; the three 2^23 biases are removed by one pre-computed constant
  const uint32_t ti = __float_as_uint(bx) + __float_as_uint(by) * M.pad_x + __float_as_uint(bz) * M.pad_xy - M.table_bias;
  const uint32_t brick = static_cast<uint32_t>(__ldg(M.table + ti));
  // sub-voxel coordinates (:62-64)
  const float qx = sub_coord<kDiv>(M, px), qy = sub_coord<kDiv>(M, py), qz = sub_coord<kDiv>(M, pz);
  return brick + __float_as_uint(qx) + __float_as_uint(qy) * M.sub_dim + __float_as_uint(qz) * M.sub_dim_2 - M.sub_bias;
}

// p' = M * p with the reference's operation order, every product and sum rounded separately
// (cuda_eval_particles.h:182-184 as the CPU build evaluates it, tsdf_evaluator.cpp:44-46).
__device__ __forceinline__ float row_apply(float a, float b, float c, float d, float x, float y, float z)
{
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, x), __fmul_rn(b, y)), __fmul_rn(c, z)), d);
}

// ---- packed fp32x2 path (Blackwell FMUL2 / FADD2 / FFMA2): two evaluations per instruction --------------------------
//
// ptxas (12.9) contracts mul.rn.f32x2 feeding add.rn.f32x2 into one FFMA2 even though both carry .rn and even with
// --fmad=false, which would break bit-parity with the reference's separately rounded products and sums. Every
// "sum of a product" is therefore written as fma(product, one, addend) with `one` = 1.0f read from a kernel
// parameter the compiler cannot see through: the product feeds the MULTIPLICAND slot, where no contraction exists,
// and fma(p, 1, q) rounds p + q exactly once — the same value as an unfused add.
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }

// Rows {a, b, c, d} applied to the point pair {xx, yy, zz}: ((a*x + b*y) + c*z) + d per half, each product and sum rounded
// separately. The halves are two particles against one point, or one particle against two points.
__device__ __forceinline__ float2 row_apply2(float2 a, float2 b, float2 c, float2 d, float2 xx, float2 yy, float2 zz, float2 one)
{
  const float2 p1 = __fmul2_rn(a, xx);
  const float2 p2 = __fmul2_rn(b, yy);
  const float2 p3 = __fmul2_rn(c, zz);
  const float2 s1 = __ffma2_rn(p1, one, p2);
  const float2 s2 = __ffma2_rn(s1, one, p3);
  return __fadd2_rn(s2, d);
}

__device__ __forceinline__ float2 clamp2(float2 o, float lo, float hi)
{
  return make_float2(fminf(fmaxf(o.x, lo), hi), fminf(fmaxf(o.y, lo), hi));
}

// 2^23 + floor(fl(a / res)) per half. kDivBracket: the floor under inv_lo; `mism` collects the bits in which the floor
// under inv_hi differs (non-zero = this quotient sits within ~4 ulps of an integer: the caller must redo it exactly).
template <int kDiv>
__device__ __forceinline__ float2 sub_coord2(const MapDev& M, float2 a, uint32_t& mism)
{
  const float2 km = dup2(kMagic);
  if (kDiv == kDivBracket)
  {
    const float2 lo = __ffma2_rd(a, dup2(M.inv_lo), km);
    const float2 hi = __ffma2_rd(a, dup2(M.inv_hi), km);
    mism |= (__float_as_uint(lo.x) ^ __float_as_uint(hi.x)) | (__float_as_uint(lo.y) ^ __float_as_uint(hi.y));
    return lo;
  }
  if (kDiv == kDivThree)
  {
    const float2 ir = dup2(M.inv_res);
    const float2 q0 = __fmul2_rn(a, ir);
    const float2 e = __ffma2_rn(dup2(-M.res), q0, a);
    return __fadd2_rd(__ffma2_rn(e, ir, q0), km);
  }
  return __fadd2_rd(make_float2(__fdiv_rn(a.x, M.res), __fdiv_rn(a.y, M.res)), km);
}

// voxel_index for two evaluations at once; same arithmetic per half as voxel_index<>.
template <int kDiv>
__device__ __forceinline__ void voxel_index2(const MapDev& M, float2 tx, float2 ty, float2 tz, uint32_t& ia, uint32_t& ib, uint32_t& mism)
{
  const float2 ox = clamp2(__fadd2_rn(tx, dup2(-M.min[0])), M.clamp_lo, M.clamp_hi[0]);
  const float2 oy = clamp2(__fadd2_rn(ty, dup2(-M.min[1])), M.clamp_lo, M.clamp_hi[1]);
  const float2 oz = clamp2(__fadd2_rn(tz, dup2(-M.min[2])), M.clamp_lo, M.clamp_hi[2]);
  const float2 kp = dup2(kMagicP1), kn = dup2(-kMagicP1);
  const float2 bx = __fadd2_rd(ox, kp);
  const float2 by = __fadd2_rd(oy, kp);
  const float2 bz = __fadd2_rd(oz, kp);
  const float2 fx = __fadd2_rn(bx, kn);
  const float2 fy = __fadd2_rn(by, kn);
  const float2 fz = __fadd2_rn(bz, kn);
  const float2 px = __fadd2_rn(ox, make_float2(-fx.x, -fx.y));
  const float2 py = __fadd2_rn(oy, make_float2(-fy.x, -fy.y));
  const float2 pz = __fadd2_rn(oz, make_float2(-fz.x, -fz.y));
  // padded strides are powers of two: shifts on the ALU pipe instead of IMADs on the FMA pipe
  const uint32_t ta = __float_as_uint(bx.x) + (__float_as_uint(by.x) << M.shift_x) + (__float_as_uint(bz.x) << M.shift_xy) - M.table_bias;
  const uint32_t tb = __float_as_uint(bx.y) + (__float_as_uint(by.y) << M.shift_x) + (__float_as_uint(bz.y) << M.shift_xy) - M.table_bias;
  const uint32_t brick_a = static_cast<uint32_t>(__ldg(M.table + ta));
  const uint32_t brick_b = static_cast<uint32_t>(__ldg(M.table + tb));
  const float2 qx = sub_coord2<kDiv>(M, px, mism);
  const float2 qy = sub_coord2<kDiv>(M, py, mism);
  const float2 qz = sub_coord2<kDiv>(M, pz, mism);
  ia = brick_a + __float_as_uint(qx.x) + __float_as_uint(qy.x) * M.sub_dim + __float_as_uint(qz.x) * M.sub_dim_2 - M.sub_bias;
  ib = brick_b + __float_as_uint(qx.y) + __float_as_uint(qy.y) * M.sub_dim + __float_as_uint(qz.y) * M.sub_dim_2 - M.sub_bias;
}

}  // namespace tsdfloc
