// tsdfloc_eval.cuh — K1: the evaluation kernel of the B200 sensor update (the ONE shape the product ships).
//
// Reference functions replaced: cudaEvaluatePose / getIndex / getEntry,
// include/tsdf_localization/cuda/cuda_eval_particles.h:84-215 (one thread per particle, P sequential iterations with two
// dependent loads each). Here:
//   * One-warp CTAs, no shared memory, 64 registers: 32 independent CTAs per SM, and all 228 KB of the SM's unified memory
//     stay L1 for voxel and brick-table sectors. Points come by LDG.128 from the prepared scan (float4 x y z range-term);
//     the L1 keeps them for the other CTAs of the SM. (Round 1 measured this shape 5-10 % ahead of a TMA-staged ring at
//     every particle count — the ring's shared memory came out of L1; profiles/r01_eval2_sweep.md.)
//   * The 32 lanes of a warp take 32 CONSECUTIVE scan points, so the 32 gathers of one warp instruction land on
//     neighbouring voxels of the same surface (13 sectors per request measured) instead of 32 unrelated particles' voxels.
//   * All fp32 arithmetic is PACKED (FMUL2 / FADD2 / FFMA2): every instruction does two evaluations. Two pairings:
//       kPP = false  two PARTICLES against one point per lane   (a warp owns 2 particles; 32 points per step)
//       kPP = true   one particle against two POINTS per lane   (a warp owns 1 particle; 64 points per step)
//     Same instructions per evaluation; kPP doubles the number of warps, which is what a small particle shard needs to fill
//     the machine (the 8,192-particle slice of an 8-GPU run is 4,096 pair-warps on 4,736 warp slots: one partial wave).
//   * The weight is the reference's SEQUENTIAL fp32 sum over the points in scan order, bit for bit (eval_sum += a_hit*v + term,
//     cuda_eval_particles.h:200-211 / tsdf_evaluator.cpp:56-67, unfused like the CPU build), reproduced in parallel:
//     while the running sum s stays inside one binade (ulp u) and the addends are >= 0,  RN(s + x) = s + RN_u(x)  unless x
//     lies exactly between two multiples of u. A block of BS steps is therefore summed as exact integers q = RN(x/u)
//     (magic-number rounding, per-lane int32 accumulators, one warp reduction per block) and accepted only if, checked
//     afterwards, no lane saw an exact tie and the sum stayed below the top of the binade; otherwise (binade crossing, tie, or
//     the first steps while s is small: ~2 % of the blocks) the block is folded truly sequentially with warp shuffles out of
//     the lanes' registers. A tree sum would differ from the reference by up to 1e-3 at P = 131k.
//   * Two REGISTER BUDGETS of the same code (kMinCtas): 64 registers = 32 one-warp CTAs per SM, the fastest once the grid is
//     several waves deep (>= 32,768 particles); 128 registers = 16 CTAs per SM, where ptxas keeps more of a block's gathers in
//     flight: a warp walks the scan 1.6x faster on its own, which is what a shard that cannot fill the machine needs
//     (profiles/r02_eval_registers.md: 500 particles 0.90 -> 0.49 ms, 8,192: 2.80 -> 2.69 ms, 65,536: 17.6 vs 18.3 ms). The host
//     picks the budget, the pairing and the chunking per launch (eval_shape, tsdfloc_api.cu).
//   * CHAINED SCAN CHUNKS for grids that end in a long, half-empty last wave (kChain): a warp's unit of work is one CHUNK of
//     the scan for its particle(s), not the whole scan. Units are handed out chunk-major through an atomic ticket (all
//     particles' chunk 0, then everybody's chunk 1, ...); chunk k of a particle starts from the running sum chunk k - 1 left in
//     global memory (it waits on a flag, but with tickets in chunk-major order the predecessor has long finished) — the sum is
//     still the one sequential fp32 sum, bit for bit, only carried from warp to warp. The tail of the grid is then one chunk
//     long instead of one scan long: 8,192 particles x 131,072 points are 1.73 waves of the 128-register budget and used to
//     cost two full waves (profiles/r02_eval_chain.md).
//   * The sub-voxel quotient floor(fl(p / res)) is bracketed by two round-down FMAs (tsdfloc_device.cuh); a block in which
//     any quotient's bracket is open (~0.5 % of the blocks) is redone with the exact division. Exactness never depends on
//     the bracket being tight — only speed does.
#pragma once
#include "tsdfloc_device.cuh"

namespace tsdfloc
{

#ifndef TSDFLOC_EVAL_BLOCK_STEPS
#define TSDFLOC_EVAL_BLOCK_STEPS 8   // tuning builds only (scripts/probes): 4 / 6 / 12 / 16 measured slower, profiles/r02_eval_registers.md
#endif
constexpr int kEvalBlockSteps = TSDFLOC_EVAL_BLOCK_STEPS;   // steps per summation block (all 16 table loads / 16 gathers of a block are independent)
constexpr int kEvalPadPoints = 64 * kEvalBlockSteps;   // the prepared scan is padded to whole blocks of the widest step

constexpr float kRoundMagic = 12582912.0f;       // 1.5 * 2^23: x + magic rounds x to an integer (RN-even) for 0 <= x < 2^22
constexpr uint32_t kRoundMagicBits = 0x4B400000u;

constexpr int kMaxPeers = 8;         // ranks of one NVSwitch domain

struct EvalArgs
{
  const float4* __restrict__ pts;   // x y z term, padded with zero points to a multiple of kEvalPadPoints (+ one block)
  const float* __restrict__ mats;   // [n_local][12]
  float* raw_out;                   // [n_local] un-normalised weights (this rank's slice of its own weight vector)
  const uint32_t* perm;             // evaluation order: slot j is particle perm[j] (nullptr = identity), tsdfloc_sort.cuh
  float* const* peer_out;           // multi-GPU: device table of n_peer_out pointers = the same slice inside every OTHER
  uint32_t n_peer_out;              //            rank's weight vector (peer-mapped, NVLink); 0 / nullptr on one GPU
  unsigned long long* __restrict__ stats;  // [4]: blocks, blocks folded sequentially, tie folds, blocks redone (open bracket)
  uint32_t* idx_out;                // kDump only: [n_local][n_points] flat voxel index (data_size = miss)
  uint32_t* hits_out;               // kDump only: [n_local] lookups that hit an allocated brick
  uint32_t n_points;
  uint32_t n_local;
  float a_hit;
  float one;                        // 1.0f, opaque to the compiler (see tsdfloc_device.cuh, packed path)
  float s_min;                      // integer-block summation is used once s >= s_min (= 32 * bound of one addend)
  uint32_t force_seq;               // 1: negative/non-finite addends possible -> always fold sequentially
  // kChain only: the scan is cut into n_chunks chunks of blocks_per_chunk summation blocks; the grid holds n_tasks * n_chunks
  // one-warp CTAs (n_tasks = particle pairs, or particles in the point-pair shape)
  uint32_t* chain;                  // [kChainHeader + kChainStride * n_tasks]: ticket, finished units; per task: flag + carried state
  uint32_t n_tasks;
  uint32_t n_chunks;
  uint32_t blocks_per_chunk;
};

constexpr uint32_t kChainHeader = 8;   // words: [0] ticket, [1] finished units (both back to 0 when the launch ends)
constexpr uint32_t kChainStride = 8;   // words per task: [0] chunks committed (0 again after the last), [1..5] s0 s1 n_fold n_tie n_redo

// Final store of a particle's weight: into this rank's vector and — fused all-gather — straight into every peer's
// (P2P stores over NVLink; the kernel boundary + the driver's signal barrier order them before the peers' reads).
__device__ __forceinline__ void store_weight(const EvalArgs& A, uint32_t slot, float w)
{
  const uint32_t part = A.perm ? A.perm[slot] : slot;
  A.raw_out[part] = w;
  for (uint32_t r = 0; r < A.n_peer_out; ++r) A.peer_out[r][part] = w;   // pointer table in global memory: no register cost in the loop
}

// The BS steps of one summation block: transform, index, gather, x = fl(fl(a_hit*v) + term), integer rounding at ulp 1/iu.
// Leaves the block's x values in xa / xb (first / second half of every pair), the per-lane integer sums in acc0 / acc1 and
// the largest |rounding residue| in mr0 / mr1 (0.5 = an exact tie). Returns the bracket-mismatch bits (kDiv == kDivBracket).
// kDump (parity tests): also records every flat voxel index and counts the lookups that hit an allocated brick.
template <int BS, int kDiv, bool kPP, bool kDump>
__device__ __forceinline__ uint32_t eval_block(const MapDev& M, const EvalArgs& A, const float2 (&mm)[12], const float4* __restrict__ bp,
                                               uint32_t lane, float2 iu, float2 one, float2 ah, float (&xa)[BS], float (&xb)[BS],
                                               uint32_t& acc0, uint32_t& acc1, float& mr0, float& mr1, uint32_t part0, uint32_t point0,
                                               uint32_t& hit0, uint32_t& hit1)
{
  uint32_t mism = 0u;
  hit0 = hit1 = 0u;
  acc0 = acc1 = 0u;
  mr0 = mr1 = 0.0f;
#pragma unroll(BS)
  for (int b = 0; b < BS; ++b)
  {
    float2 xx, yy, zz, term;
    if (kPP)
    {
      const float4 p = __ldg(bp + (b << 6) + lane), q = __ldg(bp + (b << 6) + 32 + lane);
      xx = make_float2(p.x, q.x);
      yy = make_float2(p.y, q.y);
      zz = make_float2(p.z, q.z);
      term = make_float2(p.w, q.w);
    }
    else
    {
      const float4 p = __ldg(bp + (b << 5) + lane);
      xx = dup2(p.x);
      yy = dup2(p.y);
      zz = dup2(p.z);
      term = dup2(p.w);
    }
    const float2 tx = row_apply2(mm[0], mm[1], mm[2], mm[3], xx, yy, zz, one);
    const float2 ty = row_apply2(mm[4], mm[5], mm[6], mm[7], xx, yy, zz, one);
    const float2 tz = row_apply2(mm[8], mm[9], mm[10], mm[11], xx, yy, zz, one);
    uint32_t ia, ib;
    voxel_index2<kDiv>(M, tx, ty, tz, ia, ib, mism);
    if (kDump)
    {
      // (particle, point) of the two halves; identity evaluation order
      const uint32_t pa = part0, pb = kPP ? part0 : part0 + 1;
      const uint32_t qa = point0 + (kPP ? (b << 6) : (b << 5)) + lane, qb = kPP ? qa + 32u : qa;
      const bool va = pa < A.n_local && qa < A.n_points, vb = pb < A.n_local && qb < A.n_points && (kPP || pb != pa);
      const bool ha = va && ia < M.data_size, hb = vb && ib < M.data_size;
      if (A.idx_out)
      {
        if (va) A.idx_out[static_cast<size_t>(pa) * A.n_points + qa] = ha ? ia : M.data_size;
        if (vb) A.idx_out[static_cast<size_t>(pb) * A.n_points + qb] = hb ? ib : M.data_size;
      }
      hit0 += __popc(__ballot_sync(0xffffffffu, ha));   // committed by the caller once the block is accepted
      hit1 += __popc(__ballot_sync(0xffffffffu, hb));
    }
    const float2 v = make_float2(__ldg(M.voxels + ia), __ldg(M.voxels + ib));
    const float2 x = __ffma2_rn(__fmul2_rn(ah, v), one, term);       // fl(fl(a_hit*v) + term)
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
  return mism;
}

// Sequential fp32 fold of one summation block in scan order, out of the lanes' registers.
// Particle-pair shape: one particle's values xv[b] (step b = 32 consecutive points). nvalid = steps that hold real points;
// the last of them contributes lanes [0, last_lanes).
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

// Point-pair shape: step b holds points [64 b, 64 b + 32) in xa[b] and [64 b + 32, 64 b + 64) in xb[b]; `left` = real points
// from the start of the block.
template <int BS>
__device__ __forceinline__ float fold_block_pp(float a, const float (&xa)[BS], const float (&xb)[BS], int left)
{
#pragma unroll
  for (int b = 0; b < BS; ++b)
  {
    const int ca = min(max(left - (b << 6), 0), 32), cb = min(max(left - (b << 6) - 32, 0), 32);
#pragma unroll 8
    for (int l = 0; l < ca; ++l) a = __fadd_rn(a, __shfl_sync(0xffffffffu, xa[b], l));
#pragma unroll 8
    for (int l = 0; l < cb; ++l) a = __fadd_rn(a, __shfl_sync(0xffffffffu, xb[b], l));
  }
  return a;
}

// Binade plan of a running sum s >= 0: ulp u, 1/u, the largest sum safely inside the binade, and whether integer-block
// summation may be used at all.
struct Plan
{
  float u, inv_u, limit;
  bool fast;
};
__device__ __forceinline__ Plan plan_block(const EvalArgs& A, float s)
{
  Plan p;
  const uint32_t e = __float_as_uint(s) >> 23;  // s >= 0
  p.fast = !A.force_seq && s >= A.s_min && e > 24u && e < 253u;
  const uint32_t ec = min(max(e, 25u), 252u);
  p.u = __uint_as_float((ec - 23u) << 23);
  p.inv_u = __uint_as_float((277u - ec) << 23);
  p.limit = __fsub_rn(__uint_as_float((ec + 1u) << 23), p.u);   // top of the binade minus one ulp (exact)
  return p;
}

// Running state of one warp: the sequential sums of its particle(s) and the block statistics.
struct WarpSums
{
  float s0, s1;
  uint32_t n_fold, n_tie;
};

// One summation block [point0, point0 + BS * step) against the warp's particle(s): evaluate, then commit the block to the
// running sums (integer-block shortcut where it provably equals the sequential fp32 sum, sequential fold otherwise).
// kDiv == kDivBracket: returns false WITHOUT committing anything when a sub-voxel quotient's bracket is open — the caller
// then runs the same block through an exact quotient mode.
template <int BS, int kDiv, bool kPP, bool kDump>
__device__ __forceinline__ bool eval_commit_block(const MapDev& M, const EvalArgs& A, const float2 (&mm)[12], uint32_t lane, float2 one,
                                                  float2 ah, uint32_t part0, uint32_t point0, WarpSums& W)
{
  constexpr uint32_t kBlockPoints = (kPP ? 64u : 32u) * BS;
  const float4* __restrict__ bp = A.pts + point0;
  const Plan p0 = plan_block(A, W.s0), p1 = kPP ? p0 : plan_block(A, W.s1);
  const float2 iu = make_float2(p0.inv_u, p1.inv_u);
  float xa[BS], xb[BS];
  uint32_t acc0, acc1;
  float mr0, mr1;
  uint32_t hit0, hit1;
  const uint32_t mism = eval_block<BS, kDiv, kPP, kDump>(M, A, mm, bp, lane, iu, one, ah, xa, xb, acc0, acc1, mr0, mr1, part0, point0, hit0, hit1);
  if (kDiv == kDivBracket && __any_sync(0xffffffffu, mism != 0u)) return false;
  if (kDump && lane == 0 && A.hits_out)
  {
    if (kPP) { if (part0 < A.n_local && hit0 + hit1) atomicAdd(A.hits_out + part0, hit0 + hit1); }
    else
    {
      if (part0 < A.n_local && hit0) atomicAdd(A.hits_out + part0, hit0);
      if (part0 + 1u < A.n_local && hit1) atomicAdd(A.hits_out + part0 + 1u, hit1);
    }
  }

  const bool whole = point0 + kBlockPoints <= A.n_points;   // only the scan's last block can be short
  const int left = static_cast<int>(min(kBlockPoints, A.n_points - point0));
  if (kPP)
  {
    const uint32_t tot = __reduce_add_sync(0xffffffffu, acc0 + acc1);
    const bool any_tie = __any_sync(0xffffffffu, fmaxf(mr0, mr1) == 0.5f);
    const float cand = __fadd_rn(W.s0, __fmul_rn(static_cast<float>(tot), p0.u));
    if (whole && p0.fast && !any_tie && tot < (1u << 24) && cand <= p0.limit)
      W.s0 = cand;
    else
    {
      W.s0 = fold_block_pp<BS>(W.s0, xa, xb, left);
      ++W.n_fold;
      W.n_tie += (whole && p0.fast && any_tie) ? 1u : 0u;
    }
  }
  else
  {
    const int nvalid = (left + 31) >> 5;
    const int last_lanes = left - ((nvalid - 1) << 5);
    {
      const uint32_t tot = __reduce_add_sync(0xffffffffu, acc0);
      const bool any_tie = __any_sync(0xffffffffu, mr0 == 0.5f);
      const float cand = __fadd_rn(W.s0, __fmul_rn(static_cast<float>(tot), p0.u));
      if (whole && p0.fast && !any_tie && tot < (1u << 24) && cand <= p0.limit)
        W.s0 = cand;
      else
      {
        W.s0 = fold_block<BS>(W.s0, xa, nvalid, last_lanes);
        ++W.n_fold;
        W.n_tie += (whole && p0.fast && any_tie) ? 1u : 0u;
      }
    }
    {
      const uint32_t tot = __reduce_add_sync(0xffffffffu, acc1);
      const bool any_tie = __any_sync(0xffffffffu, mr1 == 0.5f);
      const float cand = __fadd_rn(W.s1, __fmul_rn(static_cast<float>(tot), p1.u));
      if (whole && p1.fast && !any_tie && tot < (1u << 24) && cand <= p1.limit)
        W.s1 = cand;
      else
      {
        W.s1 = fold_block<BS>(W.s1, xb, nvalid, last_lanes);
        ++W.n_fold;
        W.n_tie += (whole && p1.fast && any_tie) ? 1u : 0u;
      }
    }
  }
  return true;
}

// kExact = kDivIeee or kDivThree (what k_check_div proved); kBracket: try the bracketed quotients first.
// kChain: chained scan chunks (see the header); false = one warp walks the whole scan, task = blockIdx.x.
template <int BS, int kExact, bool kBracket, bool kPP, bool kDump, int kMinCtas, bool kChain = false>
__global__ void __launch_bounds__(32, kMinCtas) k_eval(const MapDev M, const EvalArgs A)
{
  static_assert(!(kChain && kDump), "the parity dumps walk whole scans");
  constexpr uint32_t kBlockPoints = (kPP ? 64u : 32u) * BS;
  const uint32_t lane = threadIdx.x;
  const uint32_t n_blocks_total = (A.n_points + kBlockPoints - 1u) / kBlockPoints;

  uint32_t task = blockIdx.x, chunk = 0u, blk = 0u, blk_end = n_blocks_total;
  if (kChain)
  {
    // units in the order in which their CTAs actually start: whoever holds (chunk, task) knows (chunk - 1, task) is running
    // or done, so waiting for it cannot deadlock whatever order the hardware dispatches CTAs in
    uint32_t t = 0u;
    if (lane == 0) t = atomicAdd(A.chain, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    chunk = t / A.n_tasks;
    task = t - chunk * A.n_tasks;
    blk = chunk * A.blocks_per_chunk;
    blk_end = min(n_blocks_total, blk + A.blocks_per_chunk);
  }
  const uint32_t part0 = kPP ? task : task * 2u;

  // the matrices as fp32x2 pairs: {particle A, particle B}, or {particle, particle}; a warp past the end re-does the last
  // particle and stores nothing
  float2 mm[12];
  {
    const uint32_t pa = min(part0, A.n_local - 1u), pb = kPP ? pa : min(part0 + 1u, A.n_local - 1u);
#pragma unroll
    for (int e = 0; e < 12; ++e) mm[e] = make_float2(__ldg(A.mats + 12ull * pa + e), __ldg(A.mats + 12ull * pb + e));
  }
  WarpSums W{0.0f, 0.0f, 0u, 0u};
  uint32_t n_redo = 0;
  uint32_t* const link = kChain ? A.chain + kChainHeader + static_cast<size_t>(kChainStride) * task : nullptr;
  if (kChain && chunk > 0u)
  {
    // the running sums chunk - 1 left behind (L2 reads: the writer sits on another SM)
    if (lane == 0)
      while (*reinterpret_cast<volatile uint32_t*>(link) < chunk) __nanosleep(200);
    __syncwarp();
    __threadfence();
    W.s0 = __uint_as_float(__ldcg(link + 1));
    W.s1 = __uint_as_float(__ldcg(link + 2));
    W.n_fold = __ldcg(link + 3);
    W.n_tie = __ldcg(link + 4);
    n_redo = __ldcg(link + 5);
  }
  const float2 one = dup2(A.one);
  const float2 ah = dup2(A.a_hit);

  while (blk < blk_end)
  {
    // hot loop: bracketed quotients (or, where the bracket is unproven for this resolution, the exact mode directly)
#pragma unroll 1
    for (; blk < blk_end; ++blk)
      if (!eval_commit_block<BS, kBracket ? kDivBracket : kExact, kPP, kDump>(M, A, mm, lane, one, ah, part0, blk * kBlockPoints, W)) break;
    if (kBracket && blk < blk_end)
    {
      // cold: a quotient of this block sits within a few ulps of an integer — the same block with the exact division
      eval_commit_block<BS, kExact, kPP, kDump>(M, A, mm, lane, one, ah, part0, blk * kBlockPoints, W);
      ++n_redo;
      ++blk;
    }
  }

  if (lane == 0)
  {
    if (kChain && chunk + 1u < A.n_chunks)
    {
      // hand the running sums to the warp that takes this task's next chunk
      __stcg(link + 1, __float_as_uint(W.s0));
      __stcg(link + 2, __float_as_uint(W.s1));
      __stcg(link + 3, W.n_fold);
      __stcg(link + 4, W.n_tie);
      __stcg(link + 5, n_redo);
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(link) = chunk + 1u;
    }
    else
    {
      if (part0 < A.n_local) store_weight(A, part0, W.s0);
      if (!kPP && part0 + 1u < A.n_local) store_weight(A, part0 + 1u, W.s1);
      if (A.stats && part0 < A.n_local)
      {
        atomicAdd(A.stats + 0, static_cast<unsigned long long>(n_blocks_total) * (kPP ? 1ull : 2ull));
        atomicAdd(A.stats + 1, static_cast<unsigned long long>(W.n_fold));
        atomicAdd(A.stats + 2, static_cast<unsigned long long>(W.n_tie));
        atomicAdd(A.stats + 3, static_cast<unsigned long long>(n_redo));
      }
      if (kChain) *reinterpret_cast<volatile uint32_t*>(link) = 0u;   // the next launch finds the chain empty
    }
    if (kChain)
    {
      // the last unit to finish re-arms the ticket (every unit has drawn its ticket by then)
      const uint32_t done = atomicAdd(A.chain + 1, 1u);
      if (done + 1u == gridDim.x)
      {
        A.chain[1] = 0u;
        A.chain[0] = 0u;
      }
    }
  }
}

}  // namespace tsdfloc
