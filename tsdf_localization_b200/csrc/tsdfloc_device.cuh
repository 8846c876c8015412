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
  // DENSE layout (tsdfloc_create builds it when the resolution is aligned — sub_dim * res = 1 — and the bounding box fits the
  // budget): one float per voxel of the bounding box plus a one-voxel border of init_value, x fastest. Coordinate G_a of a
  // lookup = floor(sub_dim * offset_a) + 1, 0 and n_a - 1 being the border (every miss outside the map); unallocated cells hold
  // init_value. No brick table, no dependent gather. The speculative index (voxel_index2_spec) exists for this layout only.
  const float* __restrict__ dense;
  uint32_t nx, nxy;      // strides of G_y and G_z: powers of two (shifts and adds on the ALU pipe, the FMA pipe is the busy one)
  uint32_t dshift_y, dshift_z;   // log2(nx), log2(nxy)
  uint32_t dense_bias;   // kMagicBits * (1 + nx + nxy)
  int32_t g_const;       // exact path: G = bits(floor cell + 2^23 + 1) * sub_dim + bits(2^23 + q) - g_const
  float kk[3];           // K_a = thr_a * sub_dim + 1 + sub_dim * 2^-j (exact): G = floor(K_a * u) for the normalised offset u in [0, 1]
  float gamma;           // how far the reference's voxel lattice (fl(p / res) per cell) can sit from the ideal one, metres (k_check_div)
  float delta_max;       // particles whose certified margin would exceed this are evaluated exactly
  uint32_t dense_ok;     // 1: dense layout built and the lattice proof passed
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

// Reference flat index (brick offset + in-brick offset; data_size for every miss) of the dense coordinates G — parity dumps only.
__device__ __forceinline__ uint32_t ref_index_of(const MapDev& M, uint32_t gx, uint32_t gy, uint32_t gz)
{
  if (gx == 0u || gy == 0u || gz == 0u) return M.data_size;   // the border below the map; the one above it is a border cell of the table
  const uint32_t vx = gx - 1u, vy = gy - 1u, vz = gz - 1u;
  const uint32_t cx = vx / M.sub_dim, cy = vy / M.sub_dim, cz = vz / M.sub_dim;
  if (cx + 1u >= M.pad_x || cy + 1u >= M.pad_xy / M.pad_x) return M.data_size;
  const uint32_t brick = static_cast<uint32_t>(__ldg(M.table + (cx + 1u) + (cy + 1u) * M.pad_x + (cz + 1u) * M.pad_xy));
  const uint32_t idx = brick + (vx - cx * M.sub_dim) + (vy - cy * M.sub_dim) * M.sub_dim + (vz - cz * M.sub_dim) * M.sub_dim_2;
  return idx < M.data_size ? idx : M.data_size;
}

// voxel_index for two evaluations at once; same arithmetic per half as voxel_index<>. ia / ib index the array the kernel
// gathers from: M.voxels (brick layout) or — kDense — M.dense. kRef (parity dumps): ra / rb = the reference's flat index.
template <int kDiv, bool kDense, bool kRef>
__device__ __forceinline__ void voxel_index2(const MapDev& M, float2 tx, float2 ty, float2 tz, uint32_t& ia, uint32_t& ib, uint32_t& mism,
                                             uint32_t& ra, uint32_t& rb)
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
  const float2 qx = sub_coord2<kDiv>(M, px, mism);
  const float2 qy = sub_coord2<kDiv>(M, py, mism);
  const float2 qz = sub_coord2<kDiv>(M, pz, mism);
  if (kDense)
  {
    // G = cell * sub_dim + q + 1, cell in [-1, thr]: the lower border cell collapses onto G = 0
    // (unsigned wrap-around arithmetic; the result is a small signed number)
    const uint32_t sd = M.sub_dim, gc = static_cast<uint32_t>(M.g_const);
    const int32_t gxa = max(static_cast<int32_t>(__float_as_uint(bx.x) * sd + __float_as_uint(qx.x) - gc), 0);
    const int32_t gya = max(static_cast<int32_t>(__float_as_uint(by.x) * sd + __float_as_uint(qy.x) - gc), 0);
    const int32_t gza = max(static_cast<int32_t>(__float_as_uint(bz.x) * sd + __float_as_uint(qz.x) - gc), 0);
    const int32_t gxb = max(static_cast<int32_t>(__float_as_uint(bx.y) * sd + __float_as_uint(qx.y) - gc), 0);
    const int32_t gyb = max(static_cast<int32_t>(__float_as_uint(by.y) * sd + __float_as_uint(qy.y) - gc), 0);
    const int32_t gzb = max(static_cast<int32_t>(__float_as_uint(bz.y) * sd + __float_as_uint(qz.y) - gc), 0);
    ia = static_cast<uint32_t>(gxa) + static_cast<uint32_t>(gya) * M.nx + static_cast<uint32_t>(gza) * M.nxy;
    ib = static_cast<uint32_t>(gxb) + static_cast<uint32_t>(gyb) * M.nx + static_cast<uint32_t>(gzb) * M.nxy;
    if (kRef)
    {
      ra = ref_index_of(M, gxa, gya, gza);
      rb = ref_index_of(M, gxb, gyb, gzb);
    }
    return;
  }
  // padded strides are powers of two: shifts on the ALU pipe instead of IMADs on the FMA pipe
  const uint32_t ta = __float_as_uint(bx.x) + (__float_as_uint(by.x) << M.shift_x) + (__float_as_uint(bz.x) << M.shift_xy) - M.table_bias;
  const uint32_t tb = __float_as_uint(bx.y) + (__float_as_uint(by.y) << M.shift_x) + (__float_as_uint(bz.y) << M.shift_xy) - M.table_bias;
  const uint32_t brick_a = static_cast<uint32_t>(__ldg(M.table + ta));
  const uint32_t brick_b = static_cast<uint32_t>(__ldg(M.table + tb));
  ia = brick_a + __float_as_uint(qx.x) + __float_as_uint(qy.x) * M.sub_dim + __float_as_uint(qz.x) * M.sub_dim_2 - M.sub_bias;
  ib = brick_b + __float_as_uint(qx.y) + __float_as_uint(qy.y) * M.sub_dim + __float_as_uint(qz.y) * M.sub_dim_2 - M.sub_bias;
  if (kRef)
  {
    ra = ia;
    rb = ib;
  }
}

// ---- speculative voxel index (certified), dense layout --------------------------------------------------------------
//
// The exact path above spends 36 of its 41 packed instructions per step reproducing the reference's separately rounded
// transform and its floor / remainder / quotient per axis, plus twelve clamps. The speculative path computes, per axis, a
// normalised offset
//     u_lo = sat( fma(a0, x, fma(a1, y, fma(a2, z, a3))) ),   a_k = RN(m_k / S'),  a3 = RN((m3 - min + 1/sub_dim) / S') - delta_u
//     u_hi = RU( u_lo + 2 delta_u )
// (three fused multiply-adds, the last one saturating: the clamp is free) and the two floors
//     G_lo = floor(K u_lo),  G_hi = floor(K u_hi),  K = S' * sub_dim   (one round-down FMA each: exact floors of exact products).
// delta_u is chosen per particle and axis (SpecPlan, tsdfloc_eval.cuh) so that the reference's own rounded offset o_ref —
// whatever its seven roundings did — satisfies  S' u_lo - 1/sub_dim <= o_ref - gamma  and  o_ref + gamma <= S' u_hi - 1/sub_dim.
// The reference's voxel coordinate is monotone in o_ref and within gamma of the ideal lattice (k_check_div proves that
// exhaustively for the map's resolution), so G_lo <= G_ref <= G_hi: where the floors agree that IS the reference's voxel,
// bit for bit; where they differ (the point lies within ~1e-5 m of a voxel face: ~0.1 % of the evaluations) the caller
// redoes the step with the exact path. Clamped coordinates need no compare: u = 0 is the border voxel below the map, u = 1
// a point 2^-j m inside the border voxel above it, NaN saturates to 0 exactly where the exact path's fmaxf(NaN, -1) lands.
__device__ __forceinline__ float2 spec_axis(const MapDev& M, int r, float2 a0, float2 a1, float2 a2, float2 a3, float2 two_delta, float2 xx,
                                            float2 yy, float2 zz, uint32_t& mism)
{
  const float2 v1 = __ffma2_rn(a2, zz, a3);
  const float2 v2 = __ffma2_rn(a1, yy, v1);
  const float2 u = make_float2(__saturatef(__fmaf_rn(a0.x, xx.x, v2.x)), __saturatef(__fmaf_rn(a0.y, xx.y, v2.y)));
  const float2 uh = __fadd2_ru(u, two_delta);
  const float2 km = dup2(kMagic);
  const float2 gl = __ffma2_rd(u, dup2(M.kk[r]), km);
  const float2 gh = __ffma2_rd(uh, dup2(M.kk[r]), km);
  mism |= (__float_as_uint(gl.x) ^ __float_as_uint(gh.x)) | (__float_as_uint(gl.y) ^ __float_as_uint(gh.y));
  return gl;
}

// aa = the particle pair's speculative coefficients (rows of 4, slot 3 already lowered by delta_u), dl = 2 delta_u per axis.
template <bool kRef>
__device__ __forceinline__ void voxel_index2_spec(const MapDev& M, const float2 (&aa)[12], const float2 (&dl)[3], float2 xx, float2 yy, float2 zz,
                                                  uint32_t& ia, uint32_t& ib, uint32_t& mism, uint32_t& ra, uint32_t& rb)
{
  const float2 gx = spec_axis(M, 0, aa[0], aa[1], aa[2], aa[3], dl[0], xx, yy, zz, mism);
  const float2 gy = spec_axis(M, 1, aa[4], aa[5], aa[6], aa[7], dl[1], xx, yy, zz, mism);
  const float2 gz = spec_axis(M, 2, aa[8], aa[9], aa[10], aa[11], dl[2], xx, yy, zz, mism);
  ia = __float_as_uint(gx.x) + __float_as_uint(gy.x) * M.nx + __float_as_uint(gz.x) * M.nxy - M.dense_bias;
  ib = __float_as_uint(gx.y) + __float_as_uint(gy.y) * M.nx + __float_as_uint(gz.y) * M.nxy - M.dense_bias;
  if (kRef)
  {
    ra = ref_index_of(M, __float_as_uint(gx.x) - kMagicBits, __float_as_uint(gy.x) - kMagicBits, __float_as_uint(gz.x) - kMagicBits);
    rb = ref_index_of(M, __float_as_uint(gx.y) - kMagicBits, __float_as_uint(gy.y) - kMagicBits, __float_as_uint(gz.y) - kMagicBits);
  }
}

}  // namespace tsdfloc
