// tsdfloc_eval2.cuh — K1: the evaluation kernel (register-held summation blocks), in two shapes.
//
// Same arithmetic as k_eval (tsdfloc_kernels.cuh) — the reference's per-point term and its SEQUENTIAL fp32 sum, bit for
// bit — with a different execution plan, driven by the round-1 ncu profile of k_eval (profiles/r01_*): 15 % of all warp
// stalls sat on the float4 point load, another 17 % on the brick-table load, and 4 KB of shared memory per warp (the x
// staging) came out of L1.
//   * The summation block is BS steps whose x values stay in REGISTERS; the rare sequential fold (binade crossing, tie,
//     early phase) walks them with warp shuffles. No shared-memory staging of x, and all 2*BS table loads / voxel gathers
//     of a block are independent and can be in flight together.
//   * Tie detection is a running max of |residue| (2 FMNMX per step) instead of predicate bookkeeping.
//   * kDirect = true (the DEFAULT since round 1c, W = 1): one-warp CTAs without any shared memory; points by LDG.128 from the
//     prepared scan (the L1 serves the other CTAs of the SM), so the whole 228 KB of the SM are L1 for voxel sectors and 32
//     independent CTAs fill every SM. Measured 5-10 % faster than the TMA shape at every particle count
//     (profiles/r01_eval2_sweep.md): on this gather, L1 capacity is worth more than taking the point stream out of L2.
//   * kDirect = false: a CTA of W warps (2 particles per warp, packed fp32x2 math) shares ONE copy of the scan: 4 KB tiles
//     (8 steps x 32 points x float4) are pulled into a 4-stage shared-memory ring by cp.async.bulk (TMA, one elected
//     thread) and handed over with mbarriers (full: complete_tx; empty: one arrive per warp). Warps are only loosely
//     coupled — any warp may run up to 3 tiles ahead of the slowest — so there is no per-tile __syncthreads. Points come
//     from LDS.128: the L2 point traffic drops by W. Kept selectable (TSDFLOC_EVAL=2) and as the record of the experiment.
// Reference functions replaced: cudaEvaluatePose / getIndex / getEntry, include/tsdf_localization/cuda/cuda_eval_particles.h:84-215.
#pragma once
#include "tsdfloc_kernels.cuh"

namespace tsdfloc
{

constexpr int kTileSteps = 8;                    // 32-point steps per scan tile
constexpr int kTilePoints = kTileSteps * 32;     // 256 points = 4 KB
constexpr int kTileBytes = kTilePoints * 16;
#ifndef TSDFLOC_STAGES
#define TSDFLOC_STAGES 4
#endif
constexpr int kStages = TSDFLOC_STAGES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("{\n.reg .b64 state;\nmbarrier.arrive.shared::cta.b64 state, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (UBLKCP in SASS)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Sequential fp32 fold of one particle's x values of a block, in scan order, out of the lanes' registers.
// nvalid = steps of the block that hold real points; the last of them contributes lanes [0, last_lanes).
template <int BS>
__device__ __forceinline__ float fold_block(float a, const float (&xv)[BS], int nvalid, int last_lanes)
{
#pragma unroll
  for (int b = 0; b < BS; ++b)
  {
    if (b < nvalid)
    {
      const int cnt = (b == nvalid - 1) ? last_lanes : 32;
#pragma unroll 8
      for (int l = 0; l < cnt; ++l) a = __fadd_rn(a, __shfl_sync(0xffffffffu, xv[b], l));
    }
  }
  return a;
}

#ifdef TSDFLOC_EXP_NOFOLD   // timing experiment only (wrong sums): never take the sequential fold
#define TSDFLOC_NOFOLD true
#else
#define TSDFLOC_NOFOLD false
#endif

// R = resident warps per SM the register allocation is budgeted for (65536 / (32 R) registers per thread):
// R = 24 (80 registers) lets a BS = 8 block keep all its gathers in flight; R = 32 (64 registers) fits one more wave slot.
// kDirect = true (W must be 1): no shared memory at all — the points are read straight from global memory (LDG.128, L1
// keeps the tile for the other one-warp CTAs of the SM) and the whole 228 KB stay L1. This is the shape for a partial wave
// of warps (e.g. the 8,192-particle slice of an 8-GPU run), where independent one-warp CTAs pack the SMs best.
template <int W, int BS, int R, bool kFastDiv, bool kDirect = false>
__global__ void __launch_bounds__(W * 32, (R / W) > 0 ? (R / W) : 1) k_eval2(const MapDev M, const EvalArgs A)
{
  static_assert(kTileSteps % BS == 0, "a summation block must not straddle two tiles");
  static_assert(!kDirect || W == 1, "the direct variant is for one-warp CTAs");
  __shared__ __align__(128) float4 tiles[kDirect ? 1 : kStages][kDirect ? 1 : kTilePoints];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t part0 = (blockIdx.x * W + warp) * 2u;

  const uint32_t n_full = A.n_points >> 5;
  const uint32_t rem = A.n_points & 31u;
  const uint32_t n_steps = n_full + (rem ? 1u : 0u);
  const uint32_t n_tiles = (n_steps + kTileSteps - 1) / kTileSteps;
  const char* __restrict__ gsrc = reinterpret_cast<const char*>(A.pts);

  if (!kDirect && threadIdx.x == 0)
  {
#pragma unroll
    for (int s = 0; s < kStages; ++s)
    {
      mbar_init(&full_bar[s], 1u);
      mbar_init(&empty_bar[s], W);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (!kDirect) __syncthreads();
  if (!kDirect && threadIdx.x == 0)
  {
    const uint32_t pre = n_tiles < static_cast<uint32_t>(kStages) ? n_tiles : static_cast<uint32_t>(kStages);
    for (uint32_t t = 0; t < pre; ++t)
    {
      mbar_expect_tx(&full_bar[t], kTileBytes);
      bulk_g2s(tiles[t], gsrc + static_cast<size_t>(t) * kTileBytes, kTileBytes, &full_bar[t]);
    }
  }

  // the two particles' matrices as fp32x2 pairs {A, B}; warps past the end re-do the last particle and store nothing
  float2 mm[12];
  {
    const uint32_t pa = min(part0, A.n_local - 1), pb = min(part0 + 1, A.n_local - 1);
#pragma unroll
    for (int e = 0; e < 12; ++e) mm[e] = make_float2(__ldg(A.mats + 12ull * pa + e), __ldg(A.mats + 12ull * pb + e));
  }
  float s[2] = {0.0f, 0.0f};
  uint32_t n_blocks = 0, n_fold = 0, n_tie = 0;
  const float2 one = dup2(A.one);
  const float2 ah = dup2(A.a_hit);

  for (uint32_t t = 0; t < n_tiles; ++t)
  {
    const uint32_t stage = t % kStages;
    const uint32_t par = (t / kStages) & 1u;
    if (!kDirect) mbar_wait(&full_bar[stage], par);
    const float4* __restrict__ tp = kDirect ? A.pts + static_cast<size_t>(t) * kTilePoints : tiles[stage];
    const uint32_t step0 = t * kTileSteps;

#pragma unroll 1
    for (uint32_t b0 = 0; b0 < static_cast<uint32_t>(kTileSteps) && step0 + b0 < n_steps; b0 += BS)
    {
      // plan: binade of the running sums -> ulp u, 1/u, and the largest sum that still lies safely inside the binade
      float u[2], inv_u[2], limit[2];
      bool fast[2];
#pragma unroll
      for (int k = 0; k < 2; ++k)
      {
        const uint32_t e = __float_as_uint(s[k]) >> 23;  // s >= 0
        fast[k] = !A.force_seq && s[k] >= A.s_min && e > 24u && e < 253u;
        const uint32_t ec = min(max(e, 25u), 252u);
        u[k] = __uint_as_float((ec - 23u) << 23);
        inv_u[k] = __uint_as_float((277u - ec) << 23);
        limit[k] = __fsub_rn(__uint_as_float((ec + 1u) << 23), u[k]);
      }
      const float2 iu = make_float2(inv_u[0], inv_u[1]);
      uint32_t acc0 = 0u, acc1 = 0u;
      float mr0 = 0.0f, mr1 = 0.0f;
      float xa[BS], xb[BS];
#pragma unroll
      for (int b = 0; b < BS; ++b)
      {
        const float4 p = kDirect ? __ldg(tp + (((b0 + b) << 5) + lane)) : tp[((b0 + b) << 5) + lane];
        const float2 xx = dup2(p.x), yy = dup2(p.y), zz = dup2(p.z);
        const float2 tx = row_apply2(mm[0], mm[1], mm[2], mm[3], xx, yy, zz, one);
        const float2 ty = row_apply2(mm[4], mm[5], mm[6], mm[7], xx, yy, zz, one);
        const float2 tz = row_apply2(mm[8], mm[9], mm[10], mm[11], xx, yy, zz, one);
        uint32_t ia, ib;
        voxel_index2<kFastDiv>(M, tx, ty, tz, ia, ib);
        const float2 v = make_float2(__ldg(M.voxels + ia), __ldg(M.voxels + ib));
        const float2 x = __ffma2_rn(__fmul2_rn(ah, v), one, dup2(p.w));  // fl(fl(a_hit*v) + term)
        xa[b] = x.x;
        xb[b] = x.y;
        const float2 tq = __ffma2_rn(x, iu, dup2(kRoundMagic));          // RN-even(x/u) in the low mantissa bits
        acc0 += __float_as_uint(tq.x) - kRoundMagicBits;
        acc1 += __float_as_uint(tq.y) - kRoundMagicBits;
        const float2 tm = __fadd2_rn(tq, dup2(-kRoundMagic));
        const float2 r = __ffma2_rn(x, iu, make_float2(-tm.x, -tm.y));   // exact rounding residue, |r| <= 0.5
        mr0 = fmaxf(mr0, fabsf(r.x));
        mr1 = fmaxf(mr1, fabsf(r.y));
      }
      // which steps of this block hold real points (only the scan's last block can be short)
      const uint32_t g0 = step0 + b0;
      const uint32_t nvalid = min(static_cast<uint32_t>(BS), n_steps - g0);
      const bool whole = (g0 + BS <= n_full);
      const int last_lanes = (g0 + nvalid > n_full) ? static_cast<int>(rem) : 32;
      ++n_blocks;
      {
        const uint32_t tot = __reduce_add_sync(0xffffffffu, acc0);
        const bool any_tie = __any_sync(0xffffffffu, mr0 == 0.5f);
        const float cand = __fadd_rn(s[0], __fmul_rn(static_cast<float>(tot), u[0]));
        if (TSDFLOC_NOFOLD || (whole && fast[0] && !any_tie && tot < (1u << 24) && cand <= limit[0]))
          s[0] = cand;
        else
        {
          s[0] = fold_block<BS>(s[0], xa, static_cast<int>(nvalid), last_lanes);
          ++n_fold;
          n_tie += (whole && fast[0] && any_tie) ? 1u : 0u;
        }
      }
      {
        const uint32_t tot = __reduce_add_sync(0xffffffffu, acc1);
        const bool any_tie = __any_sync(0xffffffffu, mr1 == 0.5f);
        const float cand = __fadd_rn(s[1], __fmul_rn(static_cast<float>(tot), u[1]));
        if (TSDFLOC_NOFOLD || (whole && fast[1] && !any_tie && tot < (1u << 24) && cand <= limit[1]))
          s[1] = cand;
        else
        {
          s[1] = fold_block<BS>(s[1], xb, static_cast<int>(nvalid), last_lanes);
          ++n_fold;
          n_tie += (whole && fast[1] && any_tie) ? 1u : 0u;
        }
      }
    }

    // hand the stage back; one thread refills the stage released ONE tile ago (every warp has had a whole tile to leave it)
    if (kDirect) continue;
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
    if (threadIdx.x == 0 && t >= 1u && (t - 1u) + kStages < n_tiles)
    {
      const uint32_t pt = t - 1u;
      const uint32_t ps = pt % kStages;
      mbar_wait(&empty_bar[ps], (pt / kStages) & 1u);
      mbar_expect_tx(&full_bar[ps], kTileBytes);
      bulk_g2s(tiles[ps], gsrc + static_cast<size_t>(pt + kStages) * kTileBytes, kTileBytes, &full_bar[ps]);
    }
    __syncwarp();
  }

  if (lane == 0)
  {
    if (part0 < A.n_local) store_weight(A, part0, s[0]);
    if (part0 + 1 < A.n_local) store_weight(A, part0 + 1, s[1]);
    if (A.stats && part0 < A.n_local)
    {
      atomicAdd(A.stats + 0, static_cast<unsigned long long>(n_blocks) * 2ull);
      atomicAdd(A.stats + 1, static_cast<unsigned long long>(n_fold));
      atomicAdd(A.stats + 2, static_cast<unsigned long long>(n_tie));
    }
  }
}

}  // namespace tsdfloc
