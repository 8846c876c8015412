// tsdfloc_api.cu — context, map upload and the C ABI of libtsdfloc.so (see include/tsdfloc.h).
//
// Host orchestration that replaces src/cuda/cuda_evaluator.cu:21-59 (constructor: map upload) and :118-428
// (evaluate) of the reference, and hosts the GPU systematic resampler that replaces
// include/tsdf_localization/resampling/novel_resampling.h:41-72. No file-scope device globals (the reference keeps
// the map pointers in cuda_data.h:25-26): everything lives in the ctx, one per device, any number per process.
// No CPU fallback anywhere: without a CUDA device every compute entry point fails with TSDFLOC_E_CUDA.
#include "../../include/tsdfloc.h"
#include "tsdfloc_kernels.cuh"
#include "tsdfloc_reduce.cuh"
#include "tsdfloc_motion.cuh"
#include "tsdfloc_sort.cuh"

#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <string>
#include <vector>

using namespace tsdfloc;

namespace
{

thread_local std::string g_create_error;

struct DevBuf
{
  void* p = nullptr;
  size_t bytes = 0;
};

// One cached CUDA graph of a fixed-shape update (tsdfloc_graph.inc): the whole launch chain replayed with one
// cudaGraphLaunch; only the sensor transform (k_prepare) and the U table (k_draw) change between replays and are patched into
// their kernel nodes.
struct GraphSlot
{
  std::vector<unsigned char> key;   // every pointer / size / mode the captured launches depend on
  uint32_t seen = 0;                // consecutive eager calls with this key (the second one is captured)
  bool failed = false;              // capture of this key failed once: stay eager
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphNode_t node_prepare = nullptr, node_draw = nullptr;
  cudaKernelNodeParams kp_prepare{}, kp_draw{};
  std::vector<void*> args_prepare, args_draw;
  Tf12 tf{};
  UTable ut{};
  uint64_t n_kernels = 0;
  uint64_t last_use = 0;            // recency (graph_tick) for recycling
};

}  // namespace

struct tsdfloc_ctx
{
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  tsdfloc_params prm{};
  tsdfloc_map_desc desc{};
  MapDev map{};
  int sm_count = 148;
  float x_bound = 0.0f;     // upper bound of one point's contribution a_hit*v + term (k_eval block planning)
  uint32_t force_seq = 0;   // 1: contributions may be negative / non-finite -> always fold sequentially
  uint64_t launches = 0;
  int tune_shape = 0;       // tsdfloc_tune(TSDFLOC_TUNE_EVAL_PAIRING): 0 automatic, 1 particle pairs, 2 point pairs
  int tune_regs = 0;        // tsdfloc_tune(TSDFLOC_TUNE_EVAL_REGISTERS): 0 automatic, 1 64 registers (32 CTAs/SM), 2 128 registers (16 CTAs/SM)
  int tune_chunks = 0;      // tsdfloc_tune(TSDFLOC_TUNE_EVAL_CHUNKS): 0 automatic, k >= 1 scan chunks per particle (1 = whole scans)
  int tune_div = -1;        // tsdfloc_tune(TSDFLOC_TUNE_DIVISION): -1 what k_check_div proved, else kDivIeee / kDivThree / kDivBracket
  bool three_ok = false, bracket_ok = false;   // what k_check_div proved for this resolution
  unsigned long long bracket_open = 0;         // floats in [0, 1) whose bracket is open (statistics)
  cudaEvent_t ev_eval0 = nullptr, ev_eval1 = nullptr;  // bracket the last k_eval launch (tsdfloc_last_eval_ms)
  bool eval_timed = false;
  // optional per-stage timing (tsdfloc_tune(TSDFLOC_TUNE_STAGE_TIMERS, 1)): the reference's RuntimeEvaluator tasks
  // init_kernel / exec_kernel / weight_update (src/cuda/cuda_evaluator.cu:127,299,362) + the resampling stage
  enum { kEvPrep0, kEvPrep1, kEvInit0, kEvNorm0, kEvNorm1, kEvDraw0, kEvDraw1, kEvCount };
  cudaEvent_t ev_stage[kEvCount] = {};
  bool stage_timers = false;
  uint32_t stage_seen = 0;     // bit k: ev_stage[k] was recorded since the timers were switched on
  int norm_max_ctas = 0, exact_max_ctas = 0;   // co-resident CTAs of the cooperative kernels k_normalise_cdf / k_cdf_exact

  // steady-state CUDA graphs (tsdfloc_graph.inc): one slot per fixed-shape entry point
  enum { kGraphUpdateDevice, kGraphCount, kGraphWays = 4 };
  GraphSlot graphs[kGraphCount][kGraphWays];
  uint64_t graph_tick = 0;
  bool graphs_on = true;       // tsdfloc_tune(TSDFLOC_TUNE_GRAPHS)
  bool capturing = false;      // the stages are being recorded into a graph: timing events become external event nodes
  uint64_t alloc_epoch = 0;    // bumped whenever a device / pinned buffer is (re)allocated: cached graphs hold raw pointers
  uint64_t graph_replays = 0, graph_captures = 0;
  std::string graph_note;      // why the last capture attempt was abandoned (diagnostics)

  // map
  int32_t* d_table = nullptr;
  float* d_voxels = nullptr;
  float* d_free_map = nullptr;   // free-space points of a map ingested on the device (tsdfloc_create_from_chunks)
  uint64_t n_free_map = 0;

  // scan
  DevBuf d_xyz_stage, d_pts;
  uint64_t n_points = 0;

  // particles / scratch
  DevBuf d_particles, d_particles_out, d_mats, d_raw, d_cdf, d_tile_total, d_tile_offset, d_tile_moments, d_tile_best, d_parents,
      d_idx, d_hits, d_chain;
  float* d_mean = nullptr;
  unsigned long long* d_eval_stats = nullptr;  // k_eval block statistics (cumulative)
  Status* d_status = nullptr;
  uint64_t n_resident = 0;  // particles left on the device by tsdfloc_sensor_update
  bool have_cdf = false;

  // scan reduction scratch
  DevBuf d_red_in_xyz, d_red_in_ring, d_red_key4, d_red_table, d_red_cta, d_red_hist, d_red_win, d_red_rank, d_red_out, d_red_src;
  RedStatus* d_red_status = nullptr;
  RedStatus* h_red_status = nullptr;
  bool have_reduce = false;

  // multi-GPU: device copies of the peer-pointer tables handed to tsdfloc_eval_device_peers (a few distinct sets, cached)
  struct PeerTable
  {
    float* host[8] = {};
    uint32_t n = 0;
    float** dev = nullptr;
  };
  PeerTable peer_tables[8];
  uint32_t peer_tables_used = 0, peer_tables_next = 0;

  // spatial evaluation order (tsdfloc_sort.cuh)
  DevBuf d_sort_keys, d_sort_hist, d_perm;
  int sort_mode = -1;       // -1: automatic (map larger than L2 and enough particles); 0 / 1: tsdfloc_tune(TSDFLOC_TUNE_SPATIAL_ORDER)
  size_t l2_bytes = 0;
  SortArgs sort_args{};

  // run expansion (Residual / ResidualSystematic resamplers)
  DevBuf d_run_off, d_run_parent, d_wpack;

  // motion update
  DevBuf d_draws;

  // pinned staging
  void* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  Status* h_status = nullptr;
  float* h_mean = nullptr;
};

namespace
{

int fail(tsdfloc_ctx* c, int code, const std::string& msg)
{
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}

#define CU_TRY(c, expr, what)                                                                                       \
  do                                                                                                                \
  {                                                                                                                 \
    cudaError_t _e = (expr);                                                                                        \
    if (_e != cudaSuccess)                                                                                          \
      return fail((c), TSDFLOC_E_CUDA, std::string(what) + ": " + cudaGetErrorName(_e) + " (" + cudaGetErrorString(_e) + ")"); \
  } while (0)

int ensure(tsdfloc_ctx* c, DevBuf& b, size_t bytes, const char* what)
{
  if (bytes <= b.bytes) return TSDFLOC_OK;
  if (b.p)
  {
    CU_TRY(c, cudaStreamSynchronize(c->stream), "sync before regrow");
    CU_TRY(c, cudaDeviceSynchronize(), "device sync before regrow");
    CU_TRY(c, cudaFree(b.p), what);
    b.p = nullptr;
    b.bytes = 0;
  }
  size_t want = bytes + bytes / 4 + 256;
  CU_TRY(c, cudaMalloc(&b.p, want), what);
  b.bytes = want;
  ++c->alloc_epoch;
  return TSDFLOC_OK;
}

int ensure_host(tsdfloc_ctx* c, size_t bytes)
{
  if (bytes <= c->h_stage_bytes) return TSDFLOC_OK;
  if (c->h_stage)
  {
    CU_TRY(c, cudaStreamSynchronize(c->stream), "sync before host regrow");
    CU_TRY(c, cudaFreeHost(c->h_stage), "cudaFreeHost");
    c->h_stage = nullptr;
    c->h_stage_bytes = 0;
  }
  size_t want = bytes + bytes / 4 + 4096;
  CU_TRY(c, cudaMallocHost(&c->h_stage, want), "cudaMallocHost(staging)");
  c->h_stage_bytes = want;
  ++c->alloc_epoch;
  return TSDFLOC_OK;
}

cudaStream_t pick(tsdfloc_ctx* c, void* stream) { return stream ? static_cast<cudaStream_t>(stream) : c->stream; }

// Page-locked caller memory (cudaMallocHost / cudaHostRegister) is copied from / to directly; pageable memory goes through
// the ctx's pinned staging buffer.
bool is_pinned(const void* p)
{
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
  {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

struct DeviceGuard
{
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev)
  {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
    if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard()
  {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Smallest whole-metre offset m for which the reference's bound test rejects the point:
//   (unsigned)((float)m / resolution) >= dim     (cuda_eval_particles.h:14-25, cuda_sub_voxel_map.tcc:53-65)
uint32_t bound_threshold(uint64_t dim, float res)
{
  uint32_t m = 0;
  for (;;)
  {
    const volatile float q = static_cast<float>(m) / res;
    const double qd = q;
    const uint64_t g = qd >= 18446744073709551615.0 ? ~0ull : static_cast<uint64_t>(qd);
    if (g >= dim) return m;
    ++m;
    if (m == (1u << 22)) return m;
  }
}

int launch_check(tsdfloc_ctx* c, const char* what)
{
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(c, TSDFLOC_E_CUDA, std::string("launch of ") + what + " failed: " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
  ++c->launches;
  return TSDFLOC_OK;
}

// ---- stages (device pointers) ----------------------------------------------------------------------------

// Optional stage timers: record event k on s.
int mark(tsdfloc_ctx* c, int k, cudaStream_t s)
{
  if (!c->stage_timers) return TSDFLOC_OK;
  CU_TRY(c, cudaEventRecord(c->ev_stage[k], s), "event record");
  c->stage_seen |= 1u << k;
  return TSDFLOC_OK;
}

// Timing events around k_eval: inside a graph capture they must be external event-record nodes.
int record(tsdfloc_ctx* c, cudaEvent_t ev, cudaStream_t s)
{
  CU_TRY(c, c->capturing ? cudaEventRecordWithFlags(ev, s, cudaEventRecordExternal) : cudaEventRecord(ev, s), "event record");
  return TSDFLOC_OK;
}

// NVTX range over a host-side stage (visible in Nsight Systems; a no-op without a profiler attached)
struct Range
{
  explicit Range(const char* name) { nvtxRangePushA(name); }
  ~Range() { nvtxRangePop(); }
};

// The prepared scan: padded to whole summation blocks (+ one block), the pad written as zero points by the kernel itself.
int scan_layout(tsdfloc_ctx* c, const float* d_xyz, uint64_t p, PrepArgs* a)
{
  if (p > 0x7fffffffull) return fail(c, TSDFLOC_E_BAD_ARG, "scan larger than 2^31 points");
  const uint64_t padded = (p + kEvalPadPoints - 1) / kEvalPadPoints * kEvalPadPoints + kEvalPadPoints;
  int rc;
  if ((rc = ensure(c, c->d_pts, sizeof(float4) * padded, "cudaMalloc(scan)"))) return rc;
  c->n_points = p;
  a->xyz = d_xyz;
  a->out = static_cast<float4*>(c->d_pts.p);
  a->p = static_cast<uint32_t>(p);
  a->padded = static_cast<uint32_t>(padded);
  a->a_range_term = c->prm.a_range * static_cast<float>(1.0 / c->prm.max_range);
  a->a_max = c->prm.a_max;
  a->max_range_sq = c->prm.max_range * c->prm.max_range;
  return TSDFLOC_OK;
}

unsigned scan_ctas(const PrepArgs& a) { return static_cast<unsigned>(std::min<uint64_t>((static_cast<uint64_t>(a.padded) + 255) / 256, 1024)); }

int stage_prep_scan(tsdfloc_ctx* c, const float* d_xyz, uint64_t p, cudaStream_t s)
{
  Range r("tsdfloc:prep_scan");
  PrepArgs a{};
  int rc;
  if ((rc = scan_layout(c, d_xyz, p, &a))) return rc;
  if (p == 0) return TSDFLOC_OK;
  if ((rc = mark(c, tsdfloc_ctx::kEvPrep0, s))) return rc;
  k_prep_scan<<<scan_ctas(a), 256, 0, s>>>(a);
  if ((rc = launch_check(c, "k_prep_scan"))) return rc;
  return mark(c, tsdfloc_ctx::kEvPrep1, s);
}

// Device copy of a set of peer pointers. The sets repeat from update to update (two per buffer with double buffering), so
// they are cached; a new set is uploaded once (pageable source: staged by the driver before the call returns).
int peer_table(tsdfloc_ctx* c, float* const* want, uint32_t n, cudaStream_t s, float*** out)
{
  for (uint32_t k = 0; k < c->peer_tables_used; ++k)
  {
    tsdfloc_ctx::PeerTable& t = c->peer_tables[k];
    if (t.n == n && std::memcmp(t.host, want, sizeof(float*) * n) == 0)
    {
      *out = t.dev;
      return TSDFLOC_OK;
    }
  }
  uint32_t slot;
  if (c->peer_tables_used < 8) slot = c->peer_tables_used++;
  else
  {
    slot = c->peer_tables_next++ % 8;
    CU_TRY(c, cudaDeviceSynchronize(), "sync before peer-table reuse");  // a launch in flight may still read the old entry
  }
  tsdfloc_ctx::PeerTable& t = c->peer_tables[slot];
  if (!t.dev) CU_TRY(c, cudaMalloc(&t.dev, sizeof(float*) * 8), "cudaMalloc(peer table)");
  std::memset(t.host, 0, sizeof(t.host));
  std::memcpy(t.host, want, sizeof(float*) * n);
  t.n = n;
  CU_TRY(c, cudaMemcpyAsync(t.dev, t.host, sizeof(float*) * 8, cudaMemcpyHostToDevice, s), "H2D peer table");
  *out = t.dev;
  return TSDFLOC_OK;
}

// Launches the evaluation kernel: pairing (two particles per warp, or two points per lane), quotient mode and — parity
// dumps only — the index-recording instantiation of the very same code.
template <bool kPP, bool kDump, int kMinCtas>
void launch_eval_mode(const tsdfloc_ctx* c, int div, const EvalArgs& a, cudaStream_t s)
{
  constexpr int BS = kEvalBlockSteps;
  if constexpr (!kDump)
  {
    if (a.n_chunks > 1u)
    {
      const uint32_t grid = a.n_tasks * a.n_chunks;
      if (div == kDivBracket)
        k_eval<BS, kDivThree, true, kPP, false, kMinCtas, true><<<grid, 32, 0, s>>>(c->map, a);
      else if (div == kDivThree)
        k_eval<BS, kDivThree, false, kPP, false, kMinCtas, true><<<grid, 32, 0, s>>>(c->map, a);
      else
        k_eval<BS, kDivIeee, false, kPP, false, kMinCtas, true><<<grid, 32, 0, s>>>(c->map, a);
      return;
    }
  }
  const uint32_t grid = a.n_tasks;
  if (div == kDivBracket)
    k_eval<BS, kDivThree, true, kPP, kDump, kMinCtas><<<grid, 32, 0, s>>>(c->map, a);
  else if (div == kDivThree)
    k_eval<BS, kDivThree, false, kPP, kDump, kMinCtas><<<grid, 32, 0, s>>>(c->map, a);
  else
    k_eval<BS, kDivIeee, false, kPP, kDump, kMinCtas><<<grid, 32, 0, s>>>(c->map, a);
}

#ifndef TSDFLOC_EVAL_CTAS_SHALLOW
#define TSDFLOC_EVAL_CTAS_SHALLOW 32   // one-warp CTAs per SM of the two register budgets (64 | 128 registers); tuning builds only
#endif
#ifndef TSDFLOC_EVAL_CTAS_DEEP
#define TSDFLOC_EVAL_CTAS_DEEP 16
#endif

template <int kMinCtas>
void launch_eval_budget(const tsdfloc_ctx* c, bool pp, bool dump, int div, const EvalArgs& a, cudaStream_t s)
{
  if (pp)
    dump ? launch_eval_mode<true, true, kMinCtas>(c, div, a, s) : launch_eval_mode<true, false, kMinCtas>(c, div, a, s);
  else
    dump ? launch_eval_mode<false, true, kMinCtas>(c, div, a, s) : launch_eval_mode<false, false, kMinCtas>(c, div, a, s);
}

// The shape of one evaluation launch: pairing, register budget, quotient mode, and how the scan is chunked.
struct EvalShape
{
  int div = kDivIeee;
  bool pp = false, deep = false;
  uint32_t n_tasks = 0, n_chunks = 1, blocks_per_chunk = 0;
};

// Point pairs double the number of warps but read the scan once per particle instead of once per pair (+15 % at full
// occupancy): they win only while particle pairs cannot even fill the machine's warp slots once. With chained chunks a grid
// of a little more than one wave no longer costs two, so pairs take over as soon as there is more than one wave of them.
// Measured on B200 (profiles/r02_eval_registers.md, r02_eval_chain.md): 500 particles x 131,072 points 0.79 (pairs) -> 0.50 ms
// (point pairs), 2,000: 1.07 -> 0.97; 5,000: 1.64 (point pairs, chunked) vs 1.48 (pairs, chunked); 5,000 x 30,000: 0.418 vs 0.366.
EvalShape eval_shape(const tsdfloc_ctx* c, uint32_t n_local, uint32_t n_points, bool dump)
{
  EvalShape e;
  e.div = c->map.div_mode;
  if (c->tune_div >= 0)
  {
    e.div = c->tune_div;
    if (e.div == kDivBracket && !c->bracket_ok) e.div = c->map.div_mode;   // never run an unproven mode
    if (e.div != kDivIeee && !c->three_ok) e.div = kDivIeee;
  }
  const uint64_t slots_deep = static_cast<uint64_t>(c->sm_count) * TSDFLOC_EVAL_CTAS_DEEP;
  const uint64_t slots_shallow = static_cast<uint64_t>(c->sm_count) * TSDFLOC_EVAL_CTAS_SHALLOW;
  e.pp = (n_local + 1u) / 2u <= slots_deep;   // <= 4,736 particles on 148 SMs
  if (c->tune_shape == 1) e.pp = false;
  if (c->tune_shape == 2) e.pp = true;
  e.n_tasks = e.pp ? n_local : (n_local + 1u) / 2u;
  // Register budget: 128 registers (16 CTAs per SM) while the warps are fewer than 1.5 waves of the 64-register budget (32
  // CTAs per SM) — a warp on its own runs 1.6x faster with the deeper budget —, 64 registers beyond. Measured on B200
  // (profiles/r02_eval_registers.md, r02_eval_chain.md; chunked where the rule below says so): 8,192 particles 2.80 (64) vs
  // 2.33 ms (128), 12,000: 3.38 (128), 16,384: 4.46 (64) vs 4.60 (128), 32,768: 8.84 (64) vs 9.36 (128).
  e.deep = 2u * static_cast<uint64_t>(e.n_tasks) < 3u * slots_shallow;
  if (c->tune_regs == 1) e.deep = false;
  if (c->tune_regs == 2) e.deep = true;
  // Chained scan chunks (tsdfloc_eval.cuh): worth it when the grid is a few waves deep and the last one is far from full —
  // 8,192 particles are 1.73 waves of the 128-register budget and cost 2 without chunks. At least 12 summation blocks per chunk
  // (C2, 59 blocks: 4 chunks 0.419 ms, 7 chunks 0.428 ms, profiles/r02_eval_chain.md).
  const uint32_t block_points = (e.pp ? 64u : 32u) * static_cast<uint32_t>(kEvalBlockSteps);
  const uint32_t n_blocks = (n_points + block_points - 1u) / block_points;
  const uint64_t slots = e.deep ? slots_deep : slots_shallow;
  uint32_t chunks = 1;
  if (c->tune_chunks > 0) chunks = static_cast<uint32_t>(c->tune_chunks);
  else if (e.n_tasks > slots && e.n_tasks < 12u * slots)
  {
    const uint64_t waves = (e.n_tasks + slots - 1u) / slots;
    if (static_cast<double>(e.n_tasks) < 0.93 * static_cast<double>(waves * slots)) chunks = 8;
  }
  chunks = std::min(chunks, std::max(1u, n_blocks / 12u));
  if (dump || static_cast<uint64_t>(e.n_tasks) * chunks >= (1ull << 31)) chunks = 1;
  e.blocks_per_chunk = (n_blocks + chunks - 1u) / chunks;
  e.n_chunks = e.blocks_per_chunk ? (n_blocks + e.blocks_per_chunk - 1u) / e.blocks_per_chunk : 1u;
  return e;
}

void launch_eval(const tsdfloc_ctx* c, const EvalShape& e, const EvalArgs& a, cudaStream_t s, bool dump)
{
  if (e.deep)
    launch_eval_budget<TSDFLOC_EVAL_CTAS_DEEP>(c, e.pp, dump, e.div, a, s);
  else
    launch_eval_budget<TSDFLOC_EVAL_CTAS_SHALLOW>(c, e.pp, dump, e.div, a, s);
}

// Spatial evaluation order of particles [first, first + count): *perm = device permutation, or nullptr when ordering is off
// (map fits in L2, few particles, or TSDFLOC_SORT=0).
int stage_sort(tsdfloc_ctx* c, const float* d_particles, uint64_t first, uint64_t count, cudaStream_t s, const uint32_t** perm)
{
  *perm = nullptr;
  bool on = c->sort_mode == 1;
  if (c->sort_mode < 0) on = static_cast<size_t>(c->map.data_size) * 4u > c->l2_bytes && count >= 16384;
  if (!on || count < 2) return TSDFLOC_OK;
  int rc;
  const SortArgs& a = c->sort_args;
  if ((rc = ensure(c, c->d_sort_keys, sizeof(uint32_t) * count, "cudaMalloc(sort keys)"))) return rc;
  if ((rc = ensure(c, c->d_perm, sizeof(uint32_t) * count, "cudaMalloc(permutation)"))) return rc;
  if ((rc = ensure(c, c->d_sort_hist, sizeof(uint32_t) * (a.n_keys + 4), "cudaMalloc(sort histogram)"))) return rc;
  uint32_t* hist = static_cast<uint32_t*>(c->d_sort_hist.p);
  CU_TRY(c, cudaMemsetAsync(hist, 0, sizeof(uint32_t) * a.n_keys, s), "memset(sort histogram)");
  const unsigned blocks = static_cast<unsigned>((count + 255) / 256);
  k_sort_keys<<<blocks, 256, 0, s>>>(d_particles, static_cast<uint32_t>(first), static_cast<uint32_t>(count), a,
                                    static_cast<uint32_t*>(c->d_sort_keys.p), hist);
  if ((rc = launch_check(c, "k_sort_keys"))) return rc;
  k_red_scan<<<1, 1024, 0, s>>>(hist, a.n_keys, nullptr);
  if ((rc = launch_check(c, "k_red_scan"))) return rc;
  k_sort_scatter<<<blocks, 256, 0, s>>>(static_cast<const uint32_t*>(c->d_sort_keys.p), static_cast<uint32_t>(count), hist,
                                       static_cast<uint32_t*>(c->d_perm.p));
  if ((rc = launch_check(c, "k_sort_scatter"))) return rc;
  *perm = static_cast<const uint32_t*>(c->d_perm.p);
  return TSDFLOC_OK;
}

// K0, or — fused_scan != nullptr — scan preparation + K0 in one launch (the host-buffer update issues both together).
int stage_matrices(tsdfloc_ctx* c, const float* d_particles, uint64_t first, uint64_t count, const float tf[16], cudaStream_t s,
                   const uint32_t* perm, const PrepArgs* fused_scan = nullptr)
{
  int rc;
  if ((rc = ensure(c, c->d_mats, sizeof(float) * 12 * count, "cudaMalloc(matrices)"))) return rc;
  Tf12 t;
  std::memcpy(t.m, tf, sizeof(t.m));
  const unsigned mat_ctas = static_cast<unsigned>((count + 255) / 256);
  if (fused_scan)
  {
    const unsigned sc = scan_ctas(*fused_scan);
    k_prepare<<<sc + mat_ctas, 256, 0, s>>>(*fused_scan, sc, d_particles, static_cast<uint32_t>(first), static_cast<uint32_t>(count), t,
                                            static_cast<float*>(c->d_mats.p), perm);
    return launch_check(c, "k_prepare");
  }
  k_pose_matrices<<<mat_ctas, 256, 0, s>>>(d_particles, static_cast<uint32_t>(first), static_cast<uint32_t>(count), t,
                                           static_cast<float*>(c->d_mats.p), perm);
  return launch_check(c, "k_pose_matrices");
}

int stage_eval(tsdfloc_ctx* c, const float* d_particles, uint64_t n_total, uint64_t first, uint64_t count, const float tf[16],
               float* d_raw, cudaStream_t s, float* const* d_raw_peers = nullptr, uint32_t n_peers = 0, bool dump = false,
               uint32_t* d_idx = nullptr, uint32_t* d_hits = nullptr, const PrepArgs* fused_scan = nullptr)
{
  Range r("tsdfloc:eval");
  if (n_peers > static_cast<uint32_t>(kMaxPeers)) return fail(c, TSDFLOC_E_BAD_ARG, "more than 8 peer buffers");
  if (first + count > n_total) return fail(c, TSDFLOC_E_BAD_ARG, "particle slice exceeds n_total");
  if (n_total > (1ull << 24)) return fail(c, TSDFLOC_E_BAD_ARG, "more than 2^24 particles: the reference's fp32 U recurrence stalls");
  if (c->n_points == 0) return fail(c, TSDFLOC_E_EMPTY_SCAN, "empty scan");
  int rc;
  if (count == 0)
  {
    if (fused_scan)   // this rank evaluates nothing but still needs the prepared scan
    {
      k_prep_scan<<<scan_ctas(*fused_scan), 256, 0, s>>>(*fused_scan);
      return launch_check(c, "k_prep_scan");
    }
    return TSDFLOC_OK;
  }
  if ((rc = mark(c, tsdfloc_ctx::kEvInit0, s))) return rc;
  const uint32_t* perm = nullptr;
  if ((rc = stage_sort(c, d_particles, first, count, s, &perm))) return rc;
  if ((rc = stage_matrices(c, d_particles, first, count, tf, s, perm, fused_scan))) return rc;
  EvalArgs a{};
  a.perm = perm;
  a.pts = static_cast<const float4*>(c->d_pts.p);
  a.mats = static_cast<const float*>(c->d_mats.p);
  a.raw_out = d_raw + first;
  a.n_peer_out = 0;
  a.peer_out = nullptr;
  if (n_peers)
  {
    float* want[8] = {};
    uint32_t nw = 0;
    for (uint32_t r = 0; r < n_peers; ++r)
      if (d_raw_peers[r] && d_raw_peers[r] != d_raw) want[nw++] = d_raw_peers[r] + first;
    if (nw)
    {
      float** table = nullptr;
      if ((rc = peer_table(c, want, nw, s, &table))) return rc;
      a.peer_out = table;
      a.n_peer_out = nw;
    }
  }
  a.n_points = static_cast<uint32_t>(c->n_points);
  a.n_local = static_cast<uint32_t>(count);
  a.a_hit = c->prm.a_hit;
  a.one = 1.0f;
  a.s_min = 32.0f * c->x_bound;
  a.stats = c->d_eval_stats;
  a.force_seq = c->force_seq;
  a.idx_out = d_idx;
  a.hits_out = d_hits;
  const EvalShape shape = eval_shape(c, a.n_local, a.n_points, dump);
  a.n_tasks = shape.n_tasks;
  a.n_chunks = shape.n_chunks;
  a.blocks_per_chunk = shape.blocks_per_chunk;
  a.chain = nullptr;
  if (shape.n_chunks > 1u)
  {
    // ticket + per-task links of the chained chunks; zero when allocated, and every launch leaves it zero again
    const size_t words = kChainHeader + static_cast<size_t>(kChainStride) * shape.n_tasks;
    if (sizeof(uint32_t) * words > c->d_chain.bytes)
    {
      if ((rc = ensure(c, c->d_chain, sizeof(uint32_t) * words, "cudaMalloc(chunk chain)"))) return rc;
      CU_TRY(c, cudaMemsetAsync(c->d_chain.p, 0, c->d_chain.bytes, s), "memset(chunk chain)");
    }
    a.chain = static_cast<uint32_t*>(c->d_chain.p);
  }
  if ((rc = record(c, c->ev_eval0, s))) return rc;
  launch_eval(c, shape, a, s, dump);
  if ((rc = launch_check(c, "k_eval"))) return rc;
  if ((rc = record(c, c->ev_eval1, s))) return rc;
  c->eval_timed = true;
  return TSDFLOC_OK;
}

int stage_normalize(tsdfloc_ctx* c, float* d_particles, uint64_t n, const float* d_raw, float* d_mean, cudaStream_t s, float* d_w_out = nullptr)
{
  Range r("tsdfloc:weight_update");
  if (n == 0) return fail(c, TSDFLOC_E_BAD_ARG, "no particles");
  if (n > (1ull << 24)) return fail(c, TSDFLOC_E_BAD_ARG, "more than 2^24 particles: the reference's fp32 U recurrence stalls");
  const uint32_t tiles = static_cast<uint32_t>((n + kScanTile - 1) / kScanTile);
  int rc;
  if ((rc = ensure(c, c->d_cdf, sizeof(double) * n, "cudaMalloc(cdf)"))) return rc;
  if ((rc = ensure(c, c->d_tile_total, sizeof(double) * tiles, "cudaMalloc(tile totals)"))) return rc;
  if ((rc = ensure(c, c->d_tile_offset, sizeof(double) * tiles, "cudaMalloc(tile sums)"))) return rc;
  if ((rc = ensure(c, c->d_tile_moments, sizeof(double) * 9 * tiles, "cudaMalloc(tile moments)"))) return rc;
  if ((rc = ensure(c, c->d_tile_best, sizeof(unsigned long long) * tiles, "cudaMalloc(tile arg-max)"))) return rc;
  if ((rc = mark(c, tsdfloc_ctx::kEvNorm0, s))) return rc;
  NormArgs a{};
  a.particles = d_particles;
  a.raw = d_raw;      // nullptr: scan the weights the particles already carry (slot 6) without normalising them
  a.n = static_cast<uint32_t>(n);
  a.st = c->d_status;
  a.cdf = static_cast<double*>(c->d_cdf.p);
  a.tile_sum = static_cast<double*>(c->d_tile_offset.p);
  a.tile_total = static_cast<double*>(c->d_tile_total.p);
  a.tile_moments = static_cast<double*>(c->d_tile_moments.p);
  a.tile_best = static_cast<unsigned long long*>(c->d_tile_best.p);
  a.mean_pose = d_mean;
  a.w_out = d_w_out;
  // cooperative launch: the grid must be co-resident; a CTA walks several tiles when there are more tiles than that
  const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(tiles, static_cast<uint64_t>(c->norm_max_ctas)));
  void* args[] = {&a};
  if (tiles == 1)   // one CTA: no grid barrier needed, an ordinary launch
    k_normalise_cdf<false><<<1, kScanThreads, 0, s>>>(a);
  else
    CU_TRY(c, cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&k_normalise_cdf<true>), dim3(grid), dim3(kScanThreads), args, 0, s),
           "launch of k_normalise_cdf");
  if ((rc = launch_check(c, "k_normalise_cdf"))) return rc;
  // the serial-order CDF, computed only when a parallel fp64 addition rounded (every CTA returns at once otherwise); K2's
  // per-tile scratch is free again and is reused
  ExactArgs x{};
  x.particles = d_particles;
  x.n = static_cast<uint32_t>(n);
  x.st = c->d_status;
  x.cdf = static_cast<double*>(c->d_cdf.p);
  x.tile_units = static_cast<long long*>(c->d_tile_moments.p);   // 9 doubles per tile: room for 2 x int64
  x.tile_flag = static_cast<uint32_t*>(c->d_tile_best.p);
  x.tile_start = static_cast<double*>(c->d_tile_total.p);
  const unsigned xgrid = static_cast<unsigned>(std::min<uint64_t>(tiles, static_cast<uint64_t>(c->exact_max_ctas)));
  void* xargs[] = {&x};
  if (tiles == 1)
    k_cdf_exact<false><<<1, kScanThreads, 0, s>>>(x);
  else
    CU_TRY(c, cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&k_cdf_exact<true>), dim3(xgrid), dim3(kScanThreads), xargs, 0, s),
           "launch of k_cdf_exact");
  if ((rc = launch_check(c, "k_cdf_exact"))) return rc;
  if ((rc = mark(c, tsdfloc_ctx::kEvNorm1, s))) return rc;
  c->have_cdf = true;
  return TSDFLOC_OK;
}

// The reference's fp32 U recurrence as a table of exact linear runs, built here on the host: it depends only on u0 and N.
void host_u_table(float u0, uint64_t n, UTable* t)
{
  unsigned long long below = 0;
  uint32_t flags = 0;
  // no limit: the table covers every U_j the recurrence can emit before 2N + 64 outputs; the draw kernel cuts it at s_last
  t->n_segs = build_u_table(u0, 1.0 / static_cast<double>(n), HUGE_VAL, t->segs, kMaxUSegs, &below, 2ull * n + 64ull, &flags);
  t->n_elems = below;
  t->flags = flags;
}

int stage_draw(tsdfloc_ctx* c, const float* d_particles, uint64_t n, float u0, uint64_t first_out, uint64_t count_out, float* d_out,
               uint32_t* d_parents, cudaStream_t s, float* const* d_out_peers = nullptr, uint32_t n_peers = 0,
               const UTable* prebuilt = nullptr)
{
  Range r("tsdfloc:resample");
  if (n_peers > static_cast<uint32_t>(kMaxPeers)) return fail(c, TSDFLOC_E_BAD_ARG, "more than 8 peer buffers");
  DrawPeers peers{};
  for (uint32_t r = 0; r < n_peers; ++r)
    if (d_out_peers[r] && d_out_peers[r] != d_out) peers.out[peers.n++] = d_out_peers[r];
  if (!c->have_cdf) return fail(c, TSDFLOC_E_STATE, "draw before normalize");
  if (!(u0 >= 0.0f)) return fail(c, TSDFLOC_E_BAD_ARG, "u0 must be >= 0");
  if (count_out >= (1ull << 32)) return fail(c, TSDFLOC_E_BAD_ARG, "more than 2^32 output slots");
  int rc;
  if ((rc = mark(c, tsdfloc_ctx::kEvDraw0, s))) return rc;
  UTable own;
  if (!prebuilt) host_u_table(u0, n, &own);
  const UTable& t = prebuilt ? *prebuilt : own;
  // count_out == 0 still publishes n_out (one CTA)
  const unsigned grid = static_cast<unsigned>(std::max<uint64_t>(1, (count_out + 255) / 256));
  k_draw<<<grid, 256, 0, s>>>(d_particles, static_cast<const double*>(c->d_cdf.p), static_cast<uint32_t>(n), t, c->d_status, first_out,
                              static_cast<uint32_t>(count_out), d_out, d_parents, peers);
  if ((rc = launch_check(c, "k_draw"))) return rc;
  return mark(c, tsdfloc_ctx::kEvDraw1, s);
}

#include "tsdfloc_graph.inc"

// Scan reduction (tsdfloc_reduce.cuh). All pointers are device pointers.
int stage_reduce(tsdfloc_ctx* c, const float* d_xyz, const int32_t* d_ring, uint64_t n, double cell, uint32_t n_rings, uint32_t flags,
                 float* d_out, uint32_t* d_src, cudaStream_t s)
{
  if (n > (1ull << 22)) return fail(c, TSDFLOC_E_BAD_ARG, "scan reduction is limited to 2^22 points");
  if (!(cell > 0.0) || !std::isfinite(cell) || !(static_cast<float>(cell) > 0.0f)) return fail(c, TSDFLOC_E_BAD_ARG, "cell_size must be positive");
  if (n_rings == 0 || n_rings > kRedMaxRings) return fail(c, TSDFLOC_E_BAD_ARG, "n_rings must be in [1, 1024]");
  if (flags & ~(kRedFlagDesync | kRedFlagCentres)) return fail(c, TSDFLOC_E_BAD_ARG, "unknown reduction flag");
  c->have_reduce = false;
  CU_TRY(c, cudaMemsetAsync(c->d_red_status, 0, sizeof(RedStatus), s), "memset(reduce status)");
  if (n == 0)
  {
    c->have_reduce = true;
    return TSDFLOC_OK;
  }
  const uint32_t n32 = static_cast<uint32_t>(n);
  const uint32_t n_ctas = (n32 + kRedThreads - 1) / kRedThreads;
  uint32_t table_n = 1024;
  while (table_n < 2u * n32) table_n <<= 1;
  const size_t hist_n = static_cast<size_t>(n_rings) * n_ctas;
  int rc;
  if ((rc = ensure(c, c->d_red_key4, sizeof(int4) * n, "cudaMalloc(reduce keys)"))) return rc;
  if ((rc = ensure(c, c->d_red_table, sizeof(uint32_t) * table_n, "cudaMalloc(reduce table)"))) return rc;
  if ((rc = ensure(c, c->d_red_cta, sizeof(uint32_t) * n_ctas, "cudaMalloc(reduce cta counts)"))) return rc;
  if ((rc = ensure(c, c->d_red_hist, sizeof(uint32_t) * hist_n, "cudaMalloc(reduce histogram)"))) return rc;
  if ((rc = ensure(c, c->d_red_win, sizeof(int32_t) * n, "cudaMalloc(reduce winners)"))) return rc;
  if ((rc = ensure(c, c->d_red_rank, sizeof(uint32_t) * n, "cudaMalloc(reduce ranks)"))) return rc;
  CU_TRY(c, cudaMemsetAsync(c->d_red_table.p, 0xff, sizeof(uint32_t) * table_n, s), "memset(reduce table)");

  RedArgs a{};
  a.xyz = d_xyz;
  a.ring = d_ring;
  a.n = n32;
  a.n_rings = n_rings;
  a.flags = flags;
  a.n_ctas = n_ctas;
  a.res = static_cast<float>(cell);       // FLOAT_T map_res_ = reduction_cell_size (tsdf_evaluator.h:76)
  a.half = a.res / 2;                     // map_res_half_ (:77)
  a.res_d = cell;
  a.half_d = cell / 2;
  uint32_t* cta = static_cast<uint32_t*>(c->d_red_cta.p);
  uint32_t* hist = static_cast<uint32_t*>(c->d_red_hist.p);
  uint32_t* table = static_cast<uint32_t*>(c->d_red_table.p);
  int4* key4 = static_cast<int4*>(c->d_red_key4.p);
  int32_t* win = static_cast<int32_t*>(c->d_red_win.p);
  uint32_t* rank = static_cast<uint32_t*>(c->d_red_rank.p);
  k_red_mark<<<n_ctas, kRedThreads, 0, s>>>(a, cta);
  if ((rc = launch_check(c, "k_red_mark"))) return rc;
  k_red_scan<<<1, 1024, 0, s>>>(cta, n_ctas, &c->d_red_status->n_kept);
  if ((rc = launch_check(c, "k_red_scan"))) return rc;
  k_red_keys<<<n_ctas, kRedThreads, 0, s>>>(a, cta, key4, c->d_red_status);
  if ((rc = launch_check(c, "k_red_keys"))) return rc;
  k_red_insert<<<n_ctas, kRedThreads, 0, s>>>(key4, n32, table, table_n - 1);
  if ((rc = launch_check(c, "k_red_insert"))) return rc;
  k_red_count<<<n_ctas, kRedThreads, sizeof(uint16_t) * kRedWarps * n_rings, s>>>(key4, n32, table, table_n - 1, n_rings, n_ctas, win, rank, hist);
  if ((rc = launch_check(c, "k_red_count"))) return rc;
  k_red_scan<<<1, 1024, 0, s>>>(hist, static_cast<uint32_t>(hist_n), &c->d_red_status->n_out);
  if ((rc = launch_check(c, "k_red_scan"))) return rc;
  k_red_scatter<<<n_ctas, kRedThreads, 0, s>>>(a, key4, win, rank, hist, d_out, d_src);
  if ((rc = launch_check(c, "k_red_scatter"))) return rc;
  c->have_reduce = true;
  return TSDFLOC_OK;
}

int reduce_result(tsdfloc_ctx* c, uint64_t* n_out, cudaStream_t s)
{
  if (!c->have_reduce) return fail(c, TSDFLOC_E_STATE, "reduce_result before reduce_scan");
  CU_TRY(c, cudaMemcpyAsync(c->h_red_status, c->d_red_status, sizeof(RedStatus), cudaMemcpyDeviceToHost, s), "reduce status readback");
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");
  if (c->h_red_status->bad_ring) return fail(c, TSDFLOC_E_BAD_ARG, "scan reduction: ring outside [0, n_rings)");
  if (n_out) *n_out = c->h_red_status->n_out;
  return TSDFLOC_OK;
}

// Packs a strided host cloud (PointCloud2-style) into the pinned staging buffer and uploads it: xyz -> d_red_in_xyz,
// rings (widened to int32) -> d_red_in_ring.
int upload_cloud(tsdfloc_ctx* c, const void* xyz_base, uint64_t xyz_stride, const void* ring_base, uint64_t ring_stride, int ring_bytes,
                 uint64_t n, cudaStream_t s)
{
  if (n && !xyz_base) return fail(c, TSDFLOC_E_BAD_ARG, "xyz_base is NULL");
  if (xyz_stride < 12) return fail(c, TSDFLOC_E_BAD_ARG, "xyz_stride must be >= 12");
  if (ring_base && ring_bytes != 2 && ring_bytes != 4) return fail(c, TSDFLOC_E_BAD_ARG, "ring_bytes must be 2 or 4");
  if (ring_base && ring_stride < static_cast<uint64_t>(ring_bytes)) return fail(c, TSDFLOC_E_BAD_ARG, "ring_stride smaller than ring_bytes");
  if (n > (1ull << 22)) return fail(c, TSDFLOC_E_BAD_ARG, "scan reduction is limited to 2^22 points");
  int rc;
  if ((rc = ensure_host(c, 16 * (n + 1)))) return rc;
  if ((rc = ensure(c, c->d_red_in_xyz, 12 * (n + 1), "cudaMalloc(cloud xyz)"))) return rc;
  if ((rc = ensure(c, c->d_red_in_ring, 4 * (n + 1), "cudaMalloc(cloud rings)"))) return rc;
  if (n == 0) return TSDFLOC_OK;
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");  // the pinned buffer may still feed an earlier copy
  float* hx = static_cast<float*>(c->h_stage);
  int32_t* hr = reinterpret_cast<int32_t*>(static_cast<char*>(c->h_stage) + 12 * n);
  const char* xb = static_cast<const char*>(xyz_base);
  for (uint64_t i = 0; i < n; ++i) std::memcpy(hx + 3 * i, xb + i * xyz_stride, 12);
  if (ring_base)
  {
    const char* rb = static_cast<const char*>(ring_base);
    if (ring_bytes == 2)
      for (uint64_t i = 0; i < n; ++i)
      {
        int16_t v;
        std::memcpy(&v, rb + i * ring_stride, 2);
        hr[i] = v;
      }
    else
      for (uint64_t i = 0; i < n; ++i) std::memcpy(hr + i, rb + i * ring_stride, 4);
  }
  CU_TRY(c, cudaMemcpyAsync(c->d_red_in_xyz.p, hx, 12 * n, cudaMemcpyHostToDevice, s), "H2D cloud xyz");
  if (ring_base) CU_TRY(c, cudaMemcpyAsync(c->d_red_in_ring.p, hr, 4 * n, cudaMemcpyHostToDevice, s), "H2D cloud rings");
  return TSDFLOC_OK;
}

const char* overflow_text(uint32_t flags)
{
  if (flags & 4u) return "the weights sum to more than 2: the resampling recurrence would emit more than 2N + 64 particles (normalise them first)";
  return "U recurrence table overflow or stalled recurrence";
}

int read_status(tsdfloc_ctx* c, cudaStream_t s)
{
  CU_TRY(c, cudaMemcpyAsync(c->h_status, c->d_status, sizeof(Status), cudaMemcpyDeviceToHost, s), "status readback");
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");
  return TSDFLOC_OK;
}

bool valid_desc(const tsdfloc_map_desc* m, std::string& why)
{
  if (!(m->resolution > 0.0f) || !std::isfinite(m->resolution)) { why = "resolution must be positive"; return false; }
  for (int a = 0; a < 3; ++a)
  {
    if (!std::isfinite(m->min[a]) || !std::isfinite(m->max[a])) { why = "non-finite bounds"; return false; }
    if (m->up_dim[a] == 0 || m->up_dim[a] >= (1u << 20)) { why = "up_dim out of range"; return false; }
  }
  if (m->up_dim_2 != m->up_dim[0] * m->up_dim[1]) { why = "up_dim_2 != up_dim[0]*up_dim[1]"; return false; }
  if (m->grid_occ_size != m->up_dim[0] * m->up_dim[1] * m->up_dim[2]) { why = "grid_occ_size mismatch"; return false; }
  if (m->grid_occ_size >= (1ull << 31)) { why = "brick table too large"; return false; }
  if (m->sub_dim == 0 || m->sub_dim > 4096) { why = "sub_dim out of range"; return false; }
  if (m->sub_dim_2 != m->sub_dim * m->sub_dim) { why = "sub_dim_2 != sub_dim^2"; return false; }
  if (m->data_size >= (1ull << 31)) { why = "data_size exceeds the reference's int offsets (OCC_T = int)"; return false; }
  return true;
}

}  // namespace

// ---- C ABI ---------------------------------------------------------------------------------------------------

extern "C"
{

void tsdfloc_default_params(tsdfloc_params* p)
{
  if (!p) return;
  p->a_hit = 0.9f;
  p->a_range = 0.1f;
  p->a_max = 0.0f;
  p->max_range = 100.0f;
  p->per_point = 0;
  p->neg_policy = TSDFLOC_NEG_MISS;
}

int tsdfloc_abi_version(void) { return TSDFLOC_ABI_VERSION; }

const char* tsdfloc_status_string(int status)
{
  switch (status)
  {
    case TSDFLOC_OK: return "ok";
    case TSDFLOC_E_BAD_ARG: return "bad argument";
    case TSDFLOC_E_CUDA: return "CUDA error";
    case TSDFLOC_E_NO_VALID_PARTICLE: return "No particle is valid!";
    case TSDFLOC_E_EMPTY_SCAN: return "empty scan";
    case TSDFLOC_E_CAPACITY: return "output capacity exceeded";
    case TSDFLOC_E_STATE: return "call out of order";
    default: return "unknown status";
  }
}

const char* tsdfloc_last_error(const tsdfloc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

uint64_t tsdfloc_kernel_launches(const tsdfloc_ctx* ctx) { return ctx ? ctx->launches : 0; }

// Voxels that are already on the device (map ingested there): the ctx adopts the buffer (allocated with room for the miss
// brick behind data_size) and the free-space points; value range of the payload for the summation planning.
struct AdoptedVoxels
{
  float* d_voxels = nullptr;
  float vmin = 0.0f, vmax = 0.0f;
  bool finite = true;
  float* d_free = nullptr;
  uint64_t n_free = 0;
};

// in-brick offsets the index arithmetic can produce past a brick's end (sub coordinate == sub_dim): size of the miss brick
static uint64_t miss_brick_size(const tsdfloc_map_desc* map) { return map->sub_dim * (1 + map->sub_dim + map->sub_dim_2) + 1; }

static int create_impl(const tsdfloc_map_desc* map, const int32_t* grid_occ, const float* data, const AdoptedVoxels* adopt,
                       const tsdfloc_params* params, int device, tsdfloc_ctx** out)
{
  if (!out) return fail(nullptr, TSDFLOC_E_BAD_ARG, "out is NULL");
  *out = nullptr;
  if (!map || !grid_occ || (!data && !adopt && map->data_size)) return fail(nullptr, TSDFLOC_E_BAD_ARG, "map, grid_occ and data are required");
  std::string why;
  if (!valid_desc(map, why)) return fail(nullptr, TSDFLOC_E_BAD_ARG, "invalid map description: " + why);

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, TSDFLOC_E_CUDA,
                std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                    "); libtsdfloc has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(nullptr, TSDFLOC_E_BAD_ARG, "device index out of range");

  tsdfloc_ctx* c = new (std::nothrow) tsdfloc_ctx;
  if (!c) return fail(nullptr, TSDFLOC_E_BAD_ARG, "out of host memory");
  c->device = device;
  DeviceGuard guard(device);
  auto bail = [&](int code, const std::string& msg) {
    g_create_error = msg;
    if (adopt)   // the caller still owns the adopted buffers when creation fails
    {
      c->d_voxels = nullptr;
      c->d_free_map = nullptr;
    }
    tsdfloc_destroy(c);
    return code;
  };
#define CU_CREATE(expr, what)                                                                             \
  do                                                                                                      \
  {                                                                                                       \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) return bail(TSDFLOC_E_CUDA, std::string(what) + ": " + cudaGetErrorString(_e)); \
  } while (0)

  if (!guard.ok) return bail(TSDFLOC_E_CUDA, "cudaSetDevice failed");
  cudaDeviceProp prop{};
  CU_CREATE(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
  if (prop.major != 10)
    return bail(TSDFLOC_E_CUDA, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                    "; libtsdfloc is built for sm_100a (B200) only");
  c->sm_count = prop.multiProcessorCount;
  c->l2_bytes = static_cast<size_t>(prop.l2CacheSize);
  if (params) c->prm = *params; else tsdfloc_default_params(&c->prm);
  if (!(c->prm.max_range > 0.0f)) return bail(TSDFLOC_E_BAD_ARG, "max_range must be positive");
  if (c->prm.neg_policy != TSDFLOC_NEG_MISS && c->prm.neg_policy != TSDFLOC_NEG_SATURATE_LIKE_REF_GPU)
    return bail(TSDFLOC_E_BAD_ARG, "unknown neg_policy");
  c->desc = *map;
  CU_CREATE(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate");
  CU_CREATE(cudaEventCreate(&c->ev_eval0), "cudaEventCreate");
  CU_CREATE(cudaEventCreate(&c->ev_eval1), "cudaEventCreate");
  for (cudaEvent_t& e : c->ev_stage) CU_CREATE(cudaEventCreate(&e), "cudaEventCreate");
  {
    int per_sm = 0;
    CU_CREATE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_normalise_cdf<true>, kScanThreads, 0), "occupancy(k_normalise_cdf)");
    int coop = 0;
    CU_CREATE(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device), "cudaDeviceGetAttribute");
    if (!coop || per_sm < 1) return bail(TSDFLOC_E_CUDA, "device does not support cooperative launches");
    c->norm_max_ctas = per_sm * c->sm_count;
    CU_CREATE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cdf_exact<true>, kScanThreads, 0), "occupancy(k_cdf_exact)");
    if (per_sm < 1) return bail(TSDFLOC_E_CUDA, "k_cdf_exact cannot be made resident");
    c->exact_max_ctas = per_sm * c->sm_count;
  }

  // ---- padded brick table ------------------------------------------------------------------------------------
  uint32_t thr[3];
  for (int a = 0; a < 3; ++a)
  {
    thr[a] = bound_threshold(map->dim[a], map->resolution);
    if (thr[a] >= (1u << 22)) return bail(TSDFLOC_E_BAD_ARG, "map extent too large");
  }
  // x and y strides of the padded table are rounded up to powers of two: the kernel can index with shifts (ALU pipe)
  auto pow2_at_least = [](uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return p; };
  const uint64_t px = pow2_at_least(thr[0] + 2ull), py = pow2_at_least(thr[1] + 2ull), pz = thr[2] + 2ull;
  const uint64_t table_n = px * py * pz;
  if (table_n >= (1ull << 31)) return bail(TSDFLOC_E_BAD_ARG, "padded brick table too large");
  const uint64_t sub_n = map->sub_dim * map->sub_dim * map->sub_dim;
  // miss brick: large enough for the largest in-brick offset the index arithmetic can produce (sub coordinate == sub_dim)
  const uint64_t miss_n = miss_brick_size(map);
  const uint64_t miss_offset = map->data_size;
  if (miss_offset + miss_n + sub_n >= (1ull << 32)) return bail(TSDFLOC_E_BAD_ARG, "voxel array too large");
  std::vector<int32_t> table(table_n, static_cast<int32_t>(miss_offset));
  uint64_t bricks = 0;
  for (uint64_t z = 0; z < thr[2]; ++z)
    for (uint64_t y = 0; y < thr[1]; ++y)
      for (uint64_t x = 0; x < thr[0]; ++x)
      {
        // the reference's own (aliasing) upper index, cuda_eval_particles.h:44-58
        const uint64_t up_index = x + y * map->up_dim[0] + z * map->up_dim_2;
        if (up_index >= map->grid_occ_size) continue;
        const int32_t off = grid_occ[up_index];
        if (off < 0) continue;
        if (static_cast<uint64_t>(off) + sub_n > map->data_size) return bail(TSDFLOC_E_BAD_ARG, "grid_occ entry points outside data");
        table[(x + 1) + (y + 1) * px + (z + 1) * px * py] = off;
        ++bricks;
      }
  (void)bricks;
  CU_CREATE(cudaMalloc(&c->d_table, sizeof(int32_t) * table_n), "cudaMalloc(brick table)");
  CU_CREATE(cudaMemcpy(c->d_table, table.data(), sizeof(int32_t) * table_n, cudaMemcpyHostToDevice), "upload brick table");
  if (adopt)
  {
    c->d_voxels = adopt->d_voxels;      // ingested on this device: the bricks never visited the host
    c->d_free_map = adopt->d_free;
    c->n_free_map = adopt->n_free;
  }
  else
  {
    CU_CREATE(cudaMalloc(&c->d_voxels, sizeof(float) * (miss_offset + miss_n)), "cudaMalloc(voxels)");
    if (map->data_size)
      CU_CREATE(cudaMemcpy(c->d_voxels, data, sizeof(float) * map->data_size, cudaMemcpyHostToDevice), "upload voxels");
  }
  {
    std::vector<float> miss(miss_n, map->init_value);
    CU_CREATE(cudaMemcpy(c->d_voxels + miss_offset, miss.data(), sizeof(float) * miss_n, cudaMemcpyHostToDevice), "upload miss brick");
  }

  MapDev& M = c->map;
  M.table = c->d_table;
  M.voxels = c->d_voxels;
  for (int a = 0; a < 3; ++a)
  {
    M.min[a] = map->min[a];
    M.clamp_hi[a] = static_cast<float>(thr[a]);
  }
  M.clamp_lo = c->prm.neg_policy == TSDFLOC_NEG_SATURATE_LIKE_REF_GPU ? 0.0f : -1.0f;
  M.res = map->resolution;
  {
    const volatile float inv = 1.0f / map->resolution;
    M.inv_res = inv;
    M.inv_lo = M.inv_hi = inv;
  }
  M.pad_x = static_cast<uint32_t>(px);
  M.pad_xy = static_cast<uint32_t>(px * py);
  M.shift_x = 0;
  while ((1ull << M.shift_x) < px) ++M.shift_x;
  M.shift_xy = M.shift_x;
  while ((1ull << M.shift_xy) < px * py) ++M.shift_xy;
  M.sub_dim = static_cast<uint32_t>(map->sub_dim);
  M.sub_dim_2 = static_cast<uint32_t>(map->sub_dim_2);
  M.data_size = static_cast<uint32_t>(map->data_size);
  M.table_bias = kMagicBits * (1u + M.pad_x + M.pad_xy);
  M.sub_bias = kMagicBits * (1u + M.sub_dim + M.sub_dim_2);

  // ---- cell keys of the spatial evaluation order: cells of 2^k metres so that the key space stays <= 2^20 ------------------
  {
    SortArgs& sa = c->sort_args;
    float cell = 1.0f;
    for (;;)
    {
      uint32_t d[3];
      for (int a = 0; a < 3; ++a) d[a] = std::max<uint32_t>(1u, static_cast<uint32_t>(std::ceil((map->max[a] - map->min[a]) / cell)));
      uint32_t xy_bits = 0;
      while ((1u << xy_bits) < std::max(d[0], d[1])) ++xy_bits;
      uint32_t z_bits = 0;
      while ((1u << z_bits) < d[2]) ++z_bits;
      if (2 * xy_bits + z_bits <= 20)
      {
        for (int a = 0; a < 3; ++a)
        {
          sa.min[a] = map->min[a];
          sa.dim[a] = d[a];
        }
        sa.inv_cell = 1.0f / cell;
        sa.z_bits = z_bits;
        sa.n_keys = 1u << (2 * xy_bits + z_bits);
        break;
      }
      cell *= 2.0f;
    }
  }

  // ---- bound of one point's contribution, for the evaluation kernel's binade planning ---------------------------
  {
    float vmax = map->init_value, vmin = map->init_value;
    bool finite = std::isfinite(map->init_value);
    if (adopt)
    {
      finite = finite && adopt->finite;
      vmax = std::max(vmax, adopt->vmax);
      vmin = std::min(vmin, adopt->vmin);
    }
    else
      for (uint64_t i = 0; i < map->data_size; ++i)
      {
        const float v = data[i];
        if (!std::isfinite(v)) { finite = false; break; }
        vmax = std::max(vmax, v);
        vmin = std::min(vmin, v);
      }
    const float c_in = c->prm.a_range * static_cast<float>(1.0 / c->prm.max_range);
    const float c_out = c->prm.a_max;
    const bool nonneg = finite && vmin >= 0.0f && c->prm.a_hit >= 0.0f && c_in >= 0.0f && c_out >= 0.0f &&
                        std::isfinite(c->prm.a_hit) && std::isfinite(c_in) && std::isfinite(c_out);
    c->force_seq = nonneg ? 0u : 1u;
    const float xmax = c->prm.a_hit * std::max(vmax, 0.0f) + std::max(c_in, c_out);
    c->x_bound = nonneg ? xmax * 1.0001f + 1e-30f : 0.0f;
  }

  // ---- small fixed buffers -------------------------------------------------------------------------------------
  CU_CREATE(cudaMalloc(&c->d_mean, sizeof(float) * 8), "cudaMalloc(mean pose)");
  CU_CREATE(cudaMalloc(&c->d_eval_stats, sizeof(unsigned long long) * 4), "cudaMalloc(eval stats)");
  CU_CREATE(cudaMemset(c->d_eval_stats, 0, sizeof(unsigned long long) * 4), "memset(eval stats)");
  CU_CREATE(cudaMalloc(&c->d_status, sizeof(Status)), "cudaMalloc(status)");
  CU_CREATE(cudaMemset(c->d_status, 0, sizeof(Status)), "memset(status)");
  CU_CREATE(cudaMallocHost(&c->h_status, sizeof(Status)), "cudaMallocHost(status)");
  CU_CREATE(cudaMallocHost(&c->h_mean, sizeof(float) * 8), "cudaMallocHost(mean)");
  CU_CREATE(cudaMalloc(&c->d_red_status, sizeof(RedStatus)), "cudaMalloc(reduce status)");
  CU_CREATE(cudaMemset(c->d_red_status, 0, sizeof(RedStatus)), "memset(reduce status)");
  CU_CREATE(cudaMallocHost(&c->h_red_status, sizeof(RedStatus)), "cudaMallocHost(reduce status)");

  // ---- prove the quotient shortcuts for this resolution (exhaustive over every float in [0, 1)) ------------------
  {
    unsigned long long* d_out = nullptr;
    CU_CREATE(cudaMalloc(&d_out, sizeof(unsigned long long) * 5), "cudaMalloc(div check)");
    CU_CREATE(cudaMemset(d_out, 0, sizeof(unsigned long long) * 5), "memset(div check)");
    const float inv = M.inv_res;
    const float lo1 = std::nextafterf(inv, 0.0f), hi1 = std::nextafterf(inv, INFINITY);
    const float lo2 = std::nextafterf(lo1, 0.0f), hi2 = std::nextafterf(hi1, INFINITY);
    k_check_div<<<c->sm_count * 8, 256, 0, c->stream>>>(M, lo1, hi1, lo2, hi2, d_out);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess)
    {
      cudaFree(d_out);
      return bail(TSDFLOC_E_CUDA, std::string("kernel image not loadable on this device (built for sm_100a): ") + cudaGetErrorString(le));
    }
    ++c->launches;
    unsigned long long h[5] = {};
    cudaError_t ce = cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
    cudaFree(d_out);
    if (ce != cudaSuccess) return bail(TSDFLOC_E_CUDA, std::string("division self-check failed to run: ") + cudaGetErrorString(ce));
    c->three_ok = (h[0] == 0);
    // the bracket needs an exact mode for the blocks it leaves open; the redo uses the 3-instruction quotient
    if (c->three_ok && h[1] == 0) { c->bracket_ok = true; M.inv_lo = lo1; M.inv_hi = hi1; c->bracket_open = h[2]; }
    else if (c->three_ok && h[3] == 0) { c->bracket_ok = true; M.inv_lo = lo2; M.inv_hi = hi2; c->bracket_open = h[4]; }
    M.div_mode = c->bracket_ok ? kDivBracket : c->three_ok ? kDivThree : kDivIeee;
  }
#undef CU_CREATE
  *out = c;
  return TSDFLOC_OK;
}

int tsdfloc_create(const tsdfloc_map_desc* map, const int32_t* grid_occ, const float* data, const tsdfloc_params* params, int device,
                   tsdfloc_ctx** out)
{
  return create_impl(map, grid_occ, data, nullptr, params, device, out);
}

int tsdfloc_map_desc_of(const tsdfloc_ctx* c, tsdfloc_map_desc* desc)
{
  if (!c || !desc) return TSDFLOC_E_BAD_ARG;
  *desc = c->desc;
  return TSDFLOC_OK;
}

int tsdfloc_free_map_device(const tsdfloc_ctx* c, const float** d_points, uint64_t* n_points)
{
  if (!c || !d_points || !n_points) return TSDFLOC_E_BAD_ARG;
  *d_points = c->d_free_map;
  *n_points = c->n_free_map;
  return TSDFLOC_OK;
}

void tsdfloc_destroy(tsdfloc_ctx* c)
{
  if (!c) return;
  DeviceGuard guard(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  graph_drop_all(c);
  DevBuf* bufs[] = {&c->d_xyz_stage, &c->d_pts, &c->d_particles, &c->d_particles_out, &c->d_mats,
                    &c->d_raw, &c->d_cdf, &c->d_tile_total, &c->d_tile_offset, &c->d_tile_moments, &c->d_tile_best, &c->d_parents, &c->d_idx, &c->d_hits, &c->d_chain,
                    &c->d_red_in_xyz, &c->d_red_in_ring, &c->d_red_key4, &c->d_red_table, &c->d_red_cta, &c->d_red_hist, &c->d_red_win,
                    &c->d_red_rank, &c->d_red_out, &c->d_red_src, &c->d_draws, &c->d_sort_keys, &c->d_sort_hist, &c->d_perm,
                    &c->d_run_off, &c->d_run_parent, &c->d_wpack};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (c->d_table) cudaFree(c->d_table);
  if (c->d_voxels) cudaFree(c->d_voxels);
  if (c->d_free_map) cudaFree(c->d_free_map);
  if (c->d_mean) cudaFree(c->d_mean);
  if (c->d_eval_stats) cudaFree(c->d_eval_stats);
  if (c->d_status) cudaFree(c->d_status);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->h_mean) cudaFreeHost(c->h_mean);
  for (auto& t : c->peer_tables)
    if (t.dev) cudaFree(t.dev);
  if (c->d_red_status) cudaFree(c->d_red_status);
  if (c->h_red_status) cudaFreeHost(c->h_red_status);
  for (cudaEvent_t e : c->ev_stage)
    if (e) cudaEventDestroy(e);
  if (c->ev_eval0) cudaEventDestroy(c->ev_eval0);
  if (c->ev_eval1) cudaEventDestroy(c->ev_eval1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// ---- (B) device-pointer stages ----------------------------------------------------------------------------------

int tsdfloc_set_scan_device(tsdfloc_ctx* c, const float* d_points_xyz, uint64_t p, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (p && !d_points_xyz) return fail(c, TSDFLOC_E_BAD_ARG, "points is NULL");
  DeviceGuard guard(c->device);
  return stage_prep_scan(c, d_points_xyz, p, pick(c, stream));
}

int tsdfloc_set_scan_host(tsdfloc_ctx* c, const float* points_xyz, uint64_t p, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (p && !points_xyz) return fail(c, TSDFLOC_E_BAD_ARG, "points is NULL");
  DeviceGuard guard(c->device);
  cudaStream_t s = pick(c, stream);
  int rc;
  if ((rc = ensure(c, c->d_xyz_stage, sizeof(float) * 3 * (p + 1), "cudaMalloc(scan staging)"))) return rc;
  if (p) CU_TRY(c, cudaMemcpyAsync(c->d_xyz_stage.p, points_xyz, sizeof(float) * 3 * p, cudaMemcpyHostToDevice, s), "H2D scan");
  return stage_prep_scan(c, static_cast<const float*>(c->d_xyz_stage.p), p, s);
}

int tsdfloc_eval_device(tsdfloc_ctx* c, const float* d_particles, uint64_t n_total, uint64_t first, uint64_t count, const float tf[16],
                        float* d_raw_weights, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!d_particles || !tf || !d_raw_weights) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  DeviceGuard guard(c->device);
  return stage_eval(c, d_particles, n_total, first, count, tf, d_raw_weights, pick(c, stream));
}

int tsdfloc_eval_device_peers(tsdfloc_ctx* c, const float* d_particles, uint64_t n_total, uint64_t first, uint64_t count, const float tf[16],
                              float* d_raw_weights, float* const* d_raw_peers, uint32_t n_peers, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!d_particles || !tf || !d_raw_weights || (n_peers && !d_raw_peers)) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  DeviceGuard guard(c->device);
  return stage_eval(c, d_particles, n_total, first, count, tf, d_raw_weights, pick(c, stream), d_raw_peers, n_peers);
}

int tsdfloc_draw_device_peers(tsdfloc_ctx* c, const float* d_particles, uint64_t n_total, float u0, uint64_t first_out, uint64_t count_out,
                              float* d_particles_out, float* const* d_out_peers, uint32_t n_peers, uint32_t* d_parents, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!d_particles || (count_out && !d_particles_out) || (n_peers && !d_out_peers)) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  DeviceGuard guard(c->device);
  return stage_draw(c, d_particles, n_total, u0, first_out, count_out, d_particles_out, d_parents, pick(c, stream), d_out_peers, n_peers);
}

int tsdfloc_normalize_device(tsdfloc_ctx* c, float* d_particles, uint64_t n_total, const float* d_raw_weights, float* d_mean_pose,
                             void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!d_particles || !d_raw_weights) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  DeviceGuard guard(c->device);
  return stage_normalize(c, d_particles, n_total, d_raw_weights, d_mean_pose ? d_mean_pose : c->d_mean, pick(c, stream));
}

int tsdfloc_cdf_device(tsdfloc_ctx* c, float* d_particles, uint64_t n_total, float* d_mean_pose, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!d_particles) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  DeviceGuard guard(c->device);
  return stage_normalize(c, d_particles, n_total, nullptr, d_mean_pose ? d_mean_pose : c->d_mean, pick(c, stream));
}

int tsdfloc_draw_device(tsdfloc_ctx* c, const float* d_particles, uint64_t n_total, float u0, uint64_t first_out, uint64_t count_out,
                        float* d_particles_out, uint32_t* d_parents, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!d_particles || (count_out && !d_particles_out)) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  DeviceGuard guard(c->device);
  return stage_draw(c, d_particles, n_total, u0, first_out, count_out, d_particles_out, d_parents, pick(c, stream));
}

int tsdfloc_update_device(tsdfloc_ctx* c, const float* d_points_xyz, uint64_t p, float* d_particles, uint64_t n, const float tf[16], float u0,
                          float* d_particles_out, uint64_t count_out, float* d_mean_pose, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!d_points_xyz || !d_particles || !tf || (count_out && !d_particles_out)) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (p == 0) return fail(c, TSDFLOC_E_EMPTY_SCAN, "empty scan");
  if (n == 0) return fail(c, TSDFLOC_E_BAD_ARG, "no particles");
  if (!(u0 >= 0.0f)) return fail(c, TSDFLOC_E_BAD_ARG, "u0 must be >= 0");
  DeviceGuard guard(c->device);
  cudaStream_t s = pick(c, stream);
  int rc;
  if ((rc = ensure(c, c->d_raw, sizeof(float) * n, "cudaMalloc(raw weights)"))) return rc;
  c->have_cdf = false;
  PrepArgs scan{};
  if ((rc = scan_layout(c, d_points_xyz, p, &scan))) return rc;
  UTable table;
  host_u_table(u0, n, &table);
  float* d_mean = d_mean_pose ? d_mean_pose : c->d_mean;
  auto issue = [&]() -> int {
    int r;
    if ((r = stage_eval(c, d_particles, n, 0, n, tf, static_cast<float*>(c->d_raw.p), s, nullptr, 0, false, nullptr, nullptr, &scan))) return r;
    if ((r = stage_normalize(c, d_particles, n, static_cast<const float*>(c->d_raw.p), d_mean, s))) return r;
    return stage_draw(c, d_particles, n, u0, 0, count_out, d_particles_out, nullptr, s, nullptr, 0, &table);
  };
  // steady state (same buffers and sizes as the previous call): one graph launch instead of the five kernel launches
  GraphKey key;
  key.add(d_points_xyz).add(p).add(d_particles).add(n).add(d_particles_out).add(count_out).add(d_mean);
  if ((rc = run_graphed(c, tsdfloc_ctx::kGraphUpdateDevice, key, s, tf, &table, issue))) return rc;
  c->have_cdf = true;
  c->eval_timed = true;
  return TSDFLOC_OK;
}

int tsdfloc_check(tsdfloc_ctx* c, uint64_t* n_out, double* weight_sum, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  DeviceGuard guard(c->device);
  int rc = read_status(c, pick(c, stream));
  if (rc) return rc;
  if (n_out) *n_out = c->h_status->n_out;
  if (weight_sum) *weight_sum = c->h_status->weight_sum;
  if (c->h_status->zero_sum) return fail(c, TSDFLOC_E_NO_VALID_PARTICLE, "No particle is valid!");
  if (c->h_status->table_overflow & 7u) return fail(c, TSDFLOC_E_CAPACITY, overflow_text(c->h_status->table_overflow));
  return TSDFLOC_OK;
}

// ---- motion update ------------------------------------------------------------------------------------------------

static int stage_motion(tsdfloc_ctx* c, float* d_particles, uint64_t n, const double mean[6], const double sigma[6], const double* d_draws,
                        uint64_t seed, uint64_t sequence, cudaStream_t s)
{
  if (n > (1ull << 31)) return fail(c, TSDFLOC_E_BAD_ARG, "too many particles");
  if (n == 0) return TSDFLOC_OK;
  MotionArgs a{};
  for (int k = 0; k < 6; ++k)
  {
    a.mean[k] = mean ? mean[k] : 0.0;
    a.sigma[k] = sigma ? sigma[k] : 0.0;
  }
  a.seed = seed;
  a.sequence = sequence;
  k_motion_apply<<<static_cast<unsigned>((n + 127) / 128), 128, 0, s>>>(d_particles, static_cast<uint32_t>(n), d_draws, a);
  return launch_check(c, "k_motion_apply");
}

int tsdfloc_motion_update_device(tsdfloc_ctx* c, float* d_particles, uint64_t n, const double mean[6], const double sigma[6],
                                 const double* d_draws, uint64_t seed, uint64_t sequence, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (n && !d_particles) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (!d_draws && (!mean || !sigma)) return fail(c, TSDFLOC_E_BAD_ARG, "mean and sigma are required without injected draws");
  DeviceGuard guard(c->device);
  return stage_motion(c, d_particles, n, mean, sigma, d_draws, seed, sequence, pick(c, stream));
}

int tsdfloc_motion_update(tsdfloc_ctx* c, float* particles, uint64_t n, const double mean[6], const double sigma[6], const double* draws,
                          uint64_t seed, uint64_t sequence)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (n && !particles) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (!draws && (!mean || !sigma)) return fail(c, TSDFLOC_E_BAD_ARG, "mean and sigma are required without injected draws");
  if (n == 0) return TSDFLOC_OK;
  DeviceGuard guard(c->device);
  cudaStream_t s = c->stream;
  int rc;
  const size_t pbytes = sizeof(float) * 7 * n, dbytes = draws ? sizeof(double) * 6 * n : 0;
  if ((rc = ensure(c, c->d_particles, pbytes, "cudaMalloc(particles)"))) return rc;
  if (draws && (rc = ensure(c, c->d_draws, dbytes, "cudaMalloc(motion draws)"))) return rc;
  if ((rc = ensure_host(c, pbytes + dbytes + 64))) return rc;
  c->have_cdf = false;
  c->n_resident = 0;
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");
  char* h = static_cast<char*>(c->h_stage);
  const size_t doff = (pbytes + 63) / 64 * 64;
  std::memcpy(h, particles, pbytes);
  float* d_p = static_cast<float*>(c->d_particles.p);
  CU_TRY(c, cudaMemcpyAsync(d_p, h, pbytes, cudaMemcpyHostToDevice, s), "H2D particles");
  if (draws)
  {
    std::memcpy(h + doff, draws, dbytes);
    CU_TRY(c, cudaMemcpyAsync(c->d_draws.p, h + doff, dbytes, cudaMemcpyHostToDevice, s), "H2D motion draws");
  }
  if ((rc = stage_motion(c, d_p, n, mean, sigma, draws ? static_cast<const double*>(c->d_draws.p) : nullptr, seed, sequence, s))) return rc;
  CU_TRY(c, cudaMemcpyAsync(h, d_p, pbytes, cudaMemcpyDeviceToHost, s), "D2H particles");
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");
  std::memcpy(particles, h, pbytes);
  return TSDFLOC_OK;
}

static int stage_init(tsdfloc_ctx* c, float* d_particles, uint64_t n, int mode, const double mean[6], const double spread[6],
                      const float* d_free_map, uint64_t n_free, uint64_t seed, uint64_t sequence, cudaStream_t s)
{
  if (mode < kInitNormal || mode > kInitFreeMap) return fail(c, TSDFLOC_E_BAD_ARG, "unknown initialisation mode");
  if (n == 0 || n > (1ull << 24)) return fail(c, TSDFLOC_E_BAD_ARG, "particle count must be in [1, 2^24]");
  if (mode == kInitFreeMap && (!d_free_map || n_free == 0 || n_free > 0xffffffffull)) return fail(c, TSDFLOC_E_BAD_ARG, "free map required");
  InitArgs a{};
  for (int k = 0; k < 6; ++k)
  {
    a.mean[k] = mean[k];
    a.spread[k] = spread[k];
  }
  a.seed = seed;
  a.sequence = sequence;
  a.free_map = d_free_map;
  a.n_free = static_cast<uint32_t>(n_free);
  a.mode = mode;
  k_init_particles<<<static_cast<unsigned>((n + 127) / 128), 128, 0, s>>>(d_particles, static_cast<uint32_t>(n), a);
  return launch_check(c, "k_init_particles");
}

int tsdfloc_init_particles_device(tsdfloc_ctx* c, float* d_particles, uint64_t n, int mode, const double mean[6], const double spread[6],
                                  const float* d_free_map, uint64_t n_free, uint64_t seed, uint64_t sequence, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!d_particles || !mean || !spread) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  DeviceGuard guard(c->device);
  return stage_init(c, d_particles, n, mode, mean, spread, d_free_map, n_free, seed, sequence, pick(c, stream));
}

int tsdfloc_init_particles(tsdfloc_ctx* c, float* particles, uint64_t n, int mode, const double mean[6], const double spread[6],
                           const float* free_map, uint64_t n_free, uint64_t seed, uint64_t sequence)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!particles || !mean || !spread) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (n == 0 || n > (1ull << 24)) return fail(c, TSDFLOC_E_BAD_ARG, "particle count must be in [1, 2^24]");
  DeviceGuard guard(c->device);
  cudaStream_t s = c->stream;
  int rc;
  const size_t pbytes = sizeof(float) * 7 * n;
  if ((rc = ensure(c, c->d_particles, pbytes, "cudaMalloc(particles)"))) return rc;
  if ((rc = ensure_host(c, pbytes))) return rc;
  c->have_cdf = false;
  c->n_resident = 0;
  const float* d_free = nullptr;
  if (mode == kInitFreeMap)
  {
    if (!free_map && c->d_free_map && c->n_free_map)
    {
      d_free = c->d_free_map;      // the map was ingested on this device: its free-space points never left it
      n_free = c->n_free_map;
    }
    else
    {
      if (!free_map || n_free == 0) return fail(c, TSDFLOC_E_BAD_ARG, "free map required");
      if ((rc = ensure(c, c->d_draws, sizeof(float) * 3 * n_free, "cudaMalloc(free map)"))) return rc;
      CU_TRY(c, cudaMemcpyAsync(c->d_draws.p, free_map, sizeof(float) * 3 * n_free, cudaMemcpyHostToDevice, s), "H2D free map");
      d_free = static_cast<const float*>(c->d_draws.p);
    }
  }
  float* d_p = static_cast<float*>(c->d_particles.p);
  if ((rc = stage_init(c, d_p, n, mode, mean, spread, d_free, n_free, seed, sequence, s))) return rc;
  CU_TRY(c, cudaMemcpyAsync(c->h_stage, d_p, pbytes, cudaMemcpyDeviceToHost, s), "D2H particles");
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");
  std::memcpy(particles, c->h_stage, pbytes);
  return TSDFLOC_OK;
}

int tsdfloc_best_particle(tsdfloc_ctx* c, int64_t* index, float pose[6], float* weight, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!c->have_cdf) return fail(c, TSDFLOC_E_STATE, "best_particle needs a preceding sensor update / normalisation");
  DeviceGuard guard(c->device);
  int rc = read_status(c, pick(c, stream));
  if (rc) return rc;
  const unsigned long long key = c->h_status->best_key;
  if (index) *index = key ? static_cast<int64_t>(~static_cast<uint32_t>(key & 0xffffffffull)) : -1;
  if (pose)
    for (int k = 0; k < 6; ++k) pose[k] = key ? c->h_status->best_pose[k] : 0.0f;
  if (weight) *weight = key ? c->h_status->best_weight : 0.0f;
  return TSDFLOC_OK;
}

// ---- (A) host-buffer calls ---------------------------------------------------------------------------------------

// Host-buffer sensor update in parts: prepare sizes the buffers and picks the upload sources (page-locked caller memory is
// copied from directly, pageable memory goes through the pinned staging buffer); upload: scan + particles; kernels: (scan
// preparation +) evaluation, normalisation; download: weights + status; finish waits, checks, scatters the weights into the
// caller's particles. Launched kernel by kernel on purpose: here the caller waits for the result, so the host-side cost of a
// graph launch sits on the critical path and outweighs the tighter kernel spacing — measured on B200 through this call, C1
// 0.115 ms kernel by kernel, 0.119 ms with the kernel chain as a graph, 0.124 ms with the copies recorded as well; C2 0.576 /
// 0.582 / 0.585 ms (profiles/r02_graphs.md). The device-resident update (tsdfloc_update_device) does replay a graph.
struct HostUpdate
{
  float* particles = nullptr;
  uint64_t n = 0;
  const void* src_particles = nullptr;   // caller memory (page-locked) or the staging region
  char* stage = nullptr;                  // staging region of this update: particles up (pageable callers), weights back
  const void* src_scan = nullptr;         // raw scan upload (tsdfloc_sensor_update), nullptr when the prepared scan is resident
  size_t scan_bytes = 0;
  PrepArgs scan{};                        // valid when src_scan != nullptr
};

static int host_update_prepare(tsdfloc_ctx* c, float* particles, uint64_t n, size_t stage_off, HostUpdate* u)
{
  int rc;
  const size_t pbytes = sizeof(float) * 7 * n;
  if ((rc = ensure(c, c->d_particles, pbytes, "cudaMalloc(particles)"))) return rc;
  if ((rc = ensure(c, c->d_raw, sizeof(float) * n, "cudaMalloc(raw weights)"))) return rc;
  if ((rc = ensure(c, c->d_wpack, sizeof(float) * n, "cudaMalloc(packed weights)"))) return rc;
  u->particles = particles;
  u->n = n;
  u->stage = static_cast<char*>(c->h_stage) + stage_off;   // [stage_off, stage_off + pbytes) of the pinned buffer is ours
  if (is_pinned(particles)) u->src_particles = particles;
  else
  {
    std::memcpy(u->stage, particles, pbytes);
    u->src_particles = u->stage;
  }
  return TSDFLOC_OK;
}

static int host_update_upload(tsdfloc_ctx* c, const HostUpdate& u, cudaStream_t s)
{
  Range r("tsdfloc:init");
  if (u.src_scan) CU_TRY(c, cudaMemcpyAsync(c->d_xyz_stage.p, u.src_scan, u.scan_bytes, cudaMemcpyHostToDevice, s), "H2D scan");
  CU_TRY(c, cudaMemcpyAsync(c->d_particles.p, u.src_particles, sizeof(float) * 7 * u.n, cudaMemcpyHostToDevice, s), "H2D particles");
  return TSDFLOC_OK;
}

static int host_update_kernels(tsdfloc_ctx* c, const HostUpdate& u, const float tf[16], cudaStream_t s)
{
  int rc;
  float* d_p = static_cast<float*>(c->d_particles.p);
  if ((rc = stage_eval(c, d_p, u.n, 0, u.n, tf, static_cast<float*>(c->d_raw.p), s, nullptr, 0, false, nullptr, nullptr,
                       u.src_scan ? &u.scan : nullptr)))
    return rc;
  return stage_normalize(c, d_p, u.n, static_cast<const float*>(c->d_raw.p), c->d_status->mean, s, static_cast<float*>(c->d_wpack.p));
}

static int host_update_download(tsdfloc_ctx* c, const HostUpdate& u, cudaStream_t s)
{
  // 4 B per particle come back (the caller's poses are untouched, cuda_evaluator.cu:405-408) + the status block with the mean
  CU_TRY(c, cudaMemcpyAsync(u.stage, c->d_wpack.p, sizeof(float) * u.n, cudaMemcpyDeviceToHost, s), "D2H weights");
  CU_TRY(c, cudaMemcpyAsync(c->h_status, c->d_status, sizeof(Status), cudaMemcpyDeviceToHost, s), "status readback");
  return TSDFLOC_OK;
}

static int host_update_finish(tsdfloc_ctx* c, const HostUpdate& u, float mean_pose[6], cudaStream_t s)
{
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");
  c->have_cdf = true;
  c->eval_timed = true;
  if (c->h_status->zero_sum) return fail(c, TSDFLOC_E_NO_VALID_PARTICLE, "No particle is valid!");
  const float* h = reinterpret_cast<const float*>(u.stage);
  for (uint64_t i = 0; i < u.n; ++i) u.particles[7 * i + 6] = h[i];
  if (mean_pose) std::memcpy(mean_pose, c->h_status->mean, sizeof(float) * 6);
  c->n_resident = u.n;
  return TSDFLOC_OK;
}

// The prepared scan is resident (tsdfloc_sensor_update_cloud): the scan length differs from cloud to cloud, so no graph.
static int update_with_scan(tsdfloc_ctx* c, float* particles, uint64_t n, const float tf[16], float mean_pose[6], cudaStream_t s, size_t stage_off)
{
  int rc;
  HostUpdate u;
  if ((rc = host_update_prepare(c, particles, n, stage_off, &u))) return rc;
  if ((rc = host_update_upload(c, u, s))) return rc;
  if ((rc = host_update_kernels(c, u, tf, s))) return rc;
  if ((rc = host_update_download(c, u, s))) return rc;
  return host_update_finish(c, u, mean_pose, s);
}

int tsdfloc_sensor_update(tsdfloc_ctx* c, float* particles, uint64_t n, const float* points, uint64_t p, const float tf[16],
                          float mean_pose[6])
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!particles || !tf || (p && !points)) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (p == 0) return fail(c, TSDFLOC_E_EMPTY_SCAN, "empty scan: weights left untouched");
  if (n == 0) return fail(c, TSDFLOC_E_BAD_ARG, "no particles");
  DeviceGuard guard(c->device);
  cudaStream_t s = c->stream;
  int rc;
  // one pinned buffer, two regions: [particles / weights | scan] — no host sync between the two uploads
  const size_t scan_off = (sizeof(float) * 7 * n + 255) / 256 * 256;
  if ((rc = ensure_host(c, scan_off + sizeof(float) * 3 * p))) return rc;
  c->have_cdf = false;
  c->n_resident = 0;
  if ((rc = ensure(c, c->d_xyz_stage, sizeof(float) * 3 * (p + 1), "cudaMalloc(scan staging)"))) return rc;
  HostUpdate u;
  u.scan_bytes = sizeof(float) * 3 * p;
  if (is_pinned(points)) u.src_scan = points;
  else
  {
    char* h_scan = static_cast<char*>(c->h_stage) + scan_off;
    std::memcpy(h_scan, points, u.scan_bytes);
    u.src_scan = h_scan;
  }
  if ((rc = scan_layout(c, static_cast<const float*>(c->d_xyz_stage.p), p, &u.scan))) return rc;
  if ((rc = host_update_prepare(c, particles, n, 0, &u))) return rc;
  if ((rc = host_update_upload(c, u, s))) return rc;
  if ((rc = host_update_kernels(c, u, tf, s))) return rc;
  if ((rc = host_update_download(c, u, s))) return rc;
  return host_update_finish(c, u, mean_pose, s);
}

// ---- scan reduction -----------------------------------------------------------------------------------------------

int tsdfloc_reduce_scan_device(tsdfloc_ctx* c, const float* d_points_xyz, const int32_t* d_ring, uint64_t n_points, double cell_size,
                               uint32_t n_rings, uint32_t flags, float* d_points_out, uint32_t* d_src_index, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (n_points && (!d_points_xyz || !d_points_out)) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  DeviceGuard guard(c->device);
  return stage_reduce(c, d_points_xyz, d_ring, n_points, cell_size, n_rings, flags, d_points_out, d_src_index, pick(c, stream));
}

int tsdfloc_reduce_result(tsdfloc_ctx* c, uint64_t* n_out, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  DeviceGuard guard(c->device);
  return reduce_result(c, n_out, pick(c, stream));
}

// host cloud -> device -> reduced scan left in c->d_red_out (+ c->d_red_src); *m = its size
static int reduce_host_cloud(tsdfloc_ctx* c, const void* xyz_base, uint64_t xyz_stride, const void* ring_base, uint64_t ring_stride,
                             int ring_bytes, uint64_t n, double cell, uint32_t n_rings, uint32_t flags, uint64_t* m, cudaStream_t s)
{
  int rc;
  if ((rc = upload_cloud(c, xyz_base, xyz_stride, ring_base, ring_stride, ring_bytes, n, s))) return rc;
  if ((rc = ensure(c, c->d_red_out, 12 * (n + 1), "cudaMalloc(reduced scan)"))) return rc;
  if ((rc = ensure(c, c->d_red_src, 4 * (n + 1), "cudaMalloc(reduced scan sources)"))) return rc;
  if ((rc = stage_reduce(c, static_cast<const float*>(c->d_red_in_xyz.p), ring_base ? static_cast<const int32_t*>(c->d_red_in_ring.p) : nullptr,
                         n, cell, n_rings, flags, static_cast<float*>(c->d_red_out.p), static_cast<uint32_t*>(c->d_red_src.p), s)))
    return rc;
  return reduce_result(c, m, s);
}

int tsdfloc_reduce_scan(tsdfloc_ctx* c, const void* xyz_base, uint64_t xyz_stride, const void* ring_base, uint64_t ring_stride,
                        int ring_bytes, uint64_t n_points, double cell_size, uint32_t n_rings, uint32_t flags, float* points_out,
                        uint32_t* src_index, uint64_t cap, uint64_t* n_out)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!n_out || (cap && !points_out)) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  DeviceGuard guard(c->device);
  cudaStream_t s = c->stream;
  uint64_t m = 0;
  int rc;
  if ((rc = reduce_host_cloud(c, xyz_base, xyz_stride, ring_base, ring_stride, ring_bytes, n_points, cell_size, n_rings, flags, &m, s))) return rc;
  *n_out = m;
  if (m > cap) return fail(c, TSDFLOC_E_CAPACITY, "reduced scan has " + std::to_string(m) + " points, capacity is " + std::to_string(cap));
  if (m == 0) return TSDFLOC_OK;
  // the pinned buffer holds 16 B per input point: room for 12 + 4 B per output point
  float* hx = static_cast<float*>(c->h_stage);
  uint32_t* hs = reinterpret_cast<uint32_t*>(static_cast<char*>(c->h_stage) + 12 * m);
  CU_TRY(c, cudaMemcpyAsync(hx, c->d_red_out.p, 12 * m, cudaMemcpyDeviceToHost, s), "D2H reduced scan");
  if (src_index) CU_TRY(c, cudaMemcpyAsync(hs, c->d_red_src.p, 4 * m, cudaMemcpyDeviceToHost, s), "D2H reduced scan sources");
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");
  std::memcpy(points_out, hx, 12 * m);
  if (src_index) std::memcpy(src_index, hs, 4 * m);
  return TSDFLOC_OK;
}

int tsdfloc_sensor_update_cloud(tsdfloc_ctx* c, float* particles, uint64_t n, const void* xyz_base, uint64_t xyz_stride,
                                const void* ring_base, uint64_t ring_stride, int ring_bytes, uint64_t n_points, double cell_size,
                                uint32_t n_rings, uint32_t flags, const float tf[16], float mean_pose[6], uint64_t* n_points_used)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!particles || !tf) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (n == 0) return fail(c, TSDFLOC_E_BAD_ARG, "no particles");
  DeviceGuard guard(c->device);
  cudaStream_t s = c->stream;
  c->have_cdf = false;
  c->n_resident = 0;
  uint64_t m = 0;
  int rc;
  if ((rc = reduce_host_cloud(c, xyz_base, xyz_stride, ring_base, ring_stride, ring_bytes, n_points, cell_size, n_rings, flags, &m, s))) return rc;
  if (n_points_used) *n_points_used = m;
  if (m == 0) return fail(c, TSDFLOC_E_EMPTY_SCAN, "empty scan after reduction: weights left untouched");
  if ((rc = stage_prep_scan(c, static_cast<const float*>(c->d_red_out.p), m, s))) return rc;
  if ((rc = ensure_host(c, sizeof(float) * 7 * n))) return rc;
  return update_with_scan(c, particles, n, tf, mean_pose, s, 0);
}

int tsdfloc_resample_systematic(tsdfloc_ctx* c, float u0, float* particles_out, uint64_t cap, uint64_t* n_out, uint32_t* parents)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!particles_out || !n_out) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (c->n_resident == 0 || !c->have_cdf) return fail(c, TSDFLOC_E_STATE, "resample_systematic needs a preceding successful sensor_update");
  DeviceGuard guard(c->device);
  cudaStream_t s = c->stream;
  const uint64_t n = c->n_resident;
  int rc;
  if ((rc = ensure(c, c->d_particles_out, sizeof(float) * 7 * cap, "cudaMalloc(resampled particles)"))) return rc;
  if (parents && (rc = ensure(c, c->d_parents, sizeof(uint32_t) * cap, "cudaMalloc(parents)"))) return rc;
  float* d_p = static_cast<float*>(c->d_particles.p);
  if ((rc = stage_draw(c, d_p, n, u0, 0, cap, static_cast<float*>(c->d_particles_out.p), parents ? static_cast<uint32_t*>(c->d_parents.p) : nullptr, s)))
    return rc;
  // Page-locked output: the first min(cap, n) particles (the usual output length) go straight into the caller's buffer, in
  // flight together with the status read-back; whatever the recurrence emitted beyond n follows once n_out is known.
  // Pageable output: everything through the staging buffer.
  const bool direct = is_pinned(particles_out) && (!parents || is_pinned(parents));
  const uint64_t first = std::min<uint64_t>(cap, n);
  uint32_t* h_par = nullptr;
  if (direct)
  {
    CU_TRY(c, cudaMemcpyAsync(particles_out, c->d_particles_out.p, sizeof(float) * 7 * first, cudaMemcpyDeviceToHost, s), "D2H resampled particles");
    if (parents) CU_TRY(c, cudaMemcpyAsync(parents, c->d_parents.p, sizeof(uint32_t) * first, cudaMemcpyDeviceToHost, s), "D2H parents");
  }
  else
  {
    if ((rc = ensure_host(c, sizeof(float) * 7 * cap + sizeof(uint32_t) * cap))) return rc;
    h_par = reinterpret_cast<uint32_t*>(static_cast<char*>(c->h_stage) + sizeof(float) * 7 * cap);
    CU_TRY(c, cudaMemcpyAsync(c->h_stage, c->d_particles_out.p, sizeof(float) * 7 * cap, cudaMemcpyDeviceToHost, s), "D2H resampled particles");
    if (parents) CU_TRY(c, cudaMemcpyAsync(h_par, c->d_parents.p, sizeof(uint32_t) * cap, cudaMemcpyDeviceToHost, s), "D2H parents");
  }
  if ((rc = read_status(c, s))) return rc;
  if (c->h_status->table_overflow & 7u) return fail(c, TSDFLOC_E_CAPACITY, overflow_text(c->h_status->table_overflow));
  const uint64_t m = c->h_status->n_out;
  *n_out = m;
  if (m > cap) return fail(c, TSDFLOC_E_CAPACITY, "resampling emits " + std::to_string(m) + " particles, capacity is " + std::to_string(cap));
  if (direct)
  {
    if (m > first)
    {
      CU_TRY(c, cudaMemcpyAsync(particles_out + 7 * first, static_cast<const float*>(c->d_particles_out.p) + 7 * first,
                                sizeof(float) * 7 * (m - first), cudaMemcpyDeviceToHost, s), "D2H resampled particles (tail)");
      if (parents)
        CU_TRY(c, cudaMemcpyAsync(parents + first, static_cast<const uint32_t*>(c->d_parents.p) + first, sizeof(uint32_t) * (m - first),
                                  cudaMemcpyDeviceToHost, s), "D2H parents (tail)");
      CU_TRY(c, cudaStreamSynchronize(s), "stream sync");
    }
    return TSDFLOC_OK;
  }
  std::memcpy(particles_out, c->h_stage, sizeof(float) * 7 * m);
  if (parents) std::memcpy(parents, h_par, sizeof(uint32_t) * m);
  return TSDFLOC_OK;
}

int tsdfloc_resample_particles(tsdfloc_ctx* c, const float* particles, uint64_t n, float u0, float* particles_out, uint64_t cap,
                               uint64_t* n_out, uint32_t* parents)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!particles || !particles_out || !n_out) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (n == 0) return fail(c, TSDFLOC_E_BAD_ARG, "no particles");
  DeviceGuard guard(c->device);
  cudaStream_t s = c->stream;
  int rc;
  const size_t pbytes = sizeof(float) * 7 * n;
  if ((rc = ensure(c, c->d_particles, pbytes, "cudaMalloc(particles)"))) return rc;
  if ((rc = ensure_host(c, pbytes))) return rc;
  c->have_cdf = false;
  c->n_resident = 0;
  std::memcpy(c->h_stage, particles, pbytes);
  float* d_p = static_cast<float*>(c->d_particles.p);
  CU_TRY(c, cudaMemcpyAsync(d_p, c->h_stage, pbytes, cudaMemcpyHostToDevice, s), "H2D particles");
  if ((rc = stage_normalize(c, d_p, n, nullptr, c->d_mean, s))) return rc;
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");  // the pinned buffer is reused for the output
  c->n_resident = n;
  return tsdfloc_resample_systematic(c, u0, particles_out, cap, n_out, parents);
}

// ---- Residual / ResidualSystematic resampling: host recurrence + device expansion --------------------------------

int tsdfloc_resample_expand_device(tsdfloc_ctx* c, const float* d_particles, const uint32_t* d_run_off, const uint32_t* d_run_parent,
                                   uint64_t n_runs, uint64_t first_out, uint64_t count_out, float* d_particles_out,
                                   float* const* d_out_peers, uint32_t n_peers, uint32_t* d_parents, void* stream)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!d_particles || !d_run_off || (count_out && !d_particles_out) || (n_peers && !d_out_peers)) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (n_runs == 0 || n_runs > (1ull << 24)) return fail(c, TSDFLOC_E_BAD_ARG, "n_runs must be in [1, 2^24]");
  if (n_peers > static_cast<uint32_t>(kMaxPeers)) return fail(c, TSDFLOC_E_BAD_ARG, "more than 8 peer buffers");
  if (count_out == 0) return TSDFLOC_OK;
  DeviceGuard guard(c->device);
  DrawPeers peers{};
  for (uint32_t r = 0; r < n_peers; ++r)
    if (d_out_peers[r] && d_out_peers[r] != d_particles_out) peers.out[peers.n++] = d_out_peers[r];
  k_expand_runs<<<static_cast<unsigned>((count_out + 255) / 256), 256, 0, pick(c, stream)>>>(
      d_particles, d_run_off, d_run_parent, static_cast<uint32_t>(n_runs), first_out, static_cast<uint32_t>(count_out), d_particles_out,
      d_parents, peers);
  return launch_check(c, "k_expand_runs");
}

int tsdfloc_resample_expand(tsdfloc_ctx* c, const uint32_t* run_parent, const uint32_t* run_count, uint64_t n_runs, float* particles_out,
                            uint64_t cap, uint64_t* n_out, uint32_t* parents)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!run_count || !particles_out || !n_out) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (c->n_resident == 0) return fail(c, TSDFLOC_E_STATE, "resample_expand needs a particle set on the device (sensor_update / resample_particles)");
  if (n_runs == 0 || n_runs > (1ull << 24)) return fail(c, TSDFLOC_E_BAD_ARG, "n_runs must be in [1, 2^24]");
  const uint64_t n = c->n_resident;
  uint64_t total = 0;
  for (uint64_t r = 0; r < n_runs; ++r)
  {
    total += run_count[r];
    if ((run_parent ? run_parent[r] : r) >= n) return fail(c, TSDFLOC_E_BAD_ARG, "run names a particle outside the resident set");
  }
  *n_out = total;
  if (total > cap) return fail(c, TSDFLOC_E_CAPACITY, "resampling emits " + std::to_string(total) + " particles, capacity is " + std::to_string(cap));
  if (total == 0) return TSDFLOC_OK;
  if (total >= (1ull << 32)) return fail(c, TSDFLOC_E_BAD_ARG, "more than 2^32 output particles");
  DeviceGuard guard(c->device);
  cudaStream_t s = c->stream;
  int rc;
  const size_t off_bytes = sizeof(uint32_t) * (n_runs + 1), par_bytes = run_parent ? sizeof(uint32_t) * n_runs : 0;
  const size_t out_bytes = sizeof(float) * 7 * total, pidx_bytes = parents ? sizeof(uint32_t) * total : 0;
  if ((rc = ensure(c, c->d_run_off, off_bytes, "cudaMalloc(run offsets)"))) return rc;
  if (run_parent && (rc = ensure(c, c->d_run_parent, par_bytes, "cudaMalloc(run parents)"))) return rc;
  if ((rc = ensure(c, c->d_particles_out, out_bytes, "cudaMalloc(resampled particles)"))) return rc;
  if (parents && (rc = ensure(c, c->d_parents, pidx_bytes, "cudaMalloc(parents)"))) return rc;
  if ((rc = ensure_host(c, std::max(off_bytes + par_bytes, out_bytes + pidx_bytes)))) return rc;
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");   // the pinned buffer may still feed an earlier copy
  uint32_t* h_off = static_cast<uint32_t*>(c->h_stage);
  uint32_t acc = 0;
  for (uint64_t r = 0; r < n_runs; ++r)
  {
    h_off[r] = acc;
    acc += run_count[r];
  }
  h_off[n_runs] = acc;
  CU_TRY(c, cudaMemcpyAsync(c->d_run_off.p, h_off, off_bytes, cudaMemcpyHostToDevice, s), "H2D run offsets");
  if (run_parent)
  {
    uint32_t* h_par = h_off + n_runs + 1;
    std::memcpy(h_par, run_parent, par_bytes);
    CU_TRY(c, cudaMemcpyAsync(c->d_run_parent.p, h_par, par_bytes, cudaMemcpyHostToDevice, s), "H2D run parents");
  }
  if ((rc = tsdfloc_resample_expand_device(c, static_cast<const float*>(c->d_particles.p), static_cast<const uint32_t*>(c->d_run_off.p),
                                           run_parent ? static_cast<const uint32_t*>(c->d_run_parent.p) : nullptr, n_runs, 0, total,
                                           static_cast<float*>(c->d_particles_out.p), nullptr, 0,
                                           parents ? static_cast<uint32_t*>(c->d_parents.p) : nullptr, s)))
    return rc;
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");   // the offsets have left the pinned buffer; it now receives the output
  CU_TRY(c, cudaMemcpyAsync(c->h_stage, c->d_particles_out.p, out_bytes, cudaMemcpyDeviceToHost, s), "D2H resampled particles");
  uint32_t* h_pidx = reinterpret_cast<uint32_t*>(static_cast<char*>(c->h_stage) + out_bytes);
  if (parents) CU_TRY(c, cudaMemcpyAsync(h_pidx, c->d_parents.p, pidx_bytes, cudaMemcpyDeviceToHost, s), "D2H parents");
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");
  std::memcpy(particles_out, c->h_stage, out_bytes);
  if (parents) std::memcpy(parents, h_pidx, pidx_bytes);
  return TSDFLOC_OK;
}

// Weights of the resident particle set -> host (4 B per particle).
static int resident_weights_to_host(tsdfloc_ctx* c, std::vector<float>& w)
{
  const uint64_t n = c->n_resident;
  cudaStream_t s = c->stream;
  int rc;
  if ((rc = ensure(c, c->d_wpack, sizeof(float) * n, "cudaMalloc(packed weights)"))) return rc;
  if ((rc = ensure_host(c, sizeof(float) * n))) return rc;
  k_pack_weights<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(static_cast<const float*>(c->d_particles.p), static_cast<uint32_t>(n),
                                                                     static_cast<float*>(c->d_wpack.p));
  if ((rc = launch_check(c, "k_pack_weights"))) return rc;
  CU_TRY(c, cudaMemcpyAsync(c->h_stage, c->d_wpack.p, sizeof(float) * n, cudaMemcpyDeviceToHost, s), "D2H weights");
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");
  w.assign(static_cast<const float*>(c->h_stage), static_cast<const float*>(c->h_stage) + n);
  return TSDFLOC_OK;
}

// The weights a host-half resampler works on: of a weighted cloud in host memory (uploaded here so that the device half finds it
// resident), or of the set the last sensor update left on the device (4 B per particle come back).
static int weights_for_host_half(tsdfloc_ctx* c, const float* particles, uint64_t& n, std::vector<float>& w)
{
  int rc;
  if (particles)
  {
    // Resampler::resample(ParticleCloud&) on a weighted host cloud: upload it, keep the weights here
    if (n == 0 || n > (1ull << 24)) return fail(c, TSDFLOC_E_BAD_ARG, "particle count must be in [1, 2^24]");
    const size_t pbytes = sizeof(float) * 7 * n;
    if ((rc = ensure(c, c->d_particles, pbytes, "cudaMalloc(particles)"))) return rc;
    if ((rc = ensure_host(c, pbytes))) return rc;
    c->have_cdf = false;
    c->n_resident = 0;
    CU_TRY(c, cudaStreamSynchronize(c->stream), "stream sync");
    std::memcpy(c->h_stage, particles, pbytes);
    CU_TRY(c, cudaMemcpyAsync(c->d_particles.p, c->h_stage, pbytes, cudaMemcpyHostToDevice, c->stream), "H2D particles");
    w.resize(n);
    for (uint64_t i = 0; i < n; ++i) w[i] = particles[7 * i + 6];
    c->n_resident = n;
    return TSDFLOC_OK;
  }
  if (c->n_resident == 0) return fail(c, TSDFLOC_E_STATE, "resample needs a preceding successful sensor_update");
  n = c->n_resident;
  return resident_weights_to_host(c, w);
}

int tsdfloc_resample(tsdfloc_ctx* c, int method, const float* particles, uint64_t n, float u, tsdfloc_index_draw_fn draw, void* user,
                     float* particles_out, uint64_t cap, uint64_t* n_out, uint32_t* parents)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!particles_out || !n_out) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (method == TSDFLOC_RESAMPLE_SYSTEMATIC)
    return particles ? tsdfloc_resample_particles(c, particles, n, u, particles_out, cap, n_out, parents)
                     : tsdfloc_resample_systematic(c, u, particles_out, cap, n_out, parents);
  if (method != TSDFLOC_RESAMPLE_RESIDUAL && method != TSDFLOC_RESAMPLE_RESIDUAL_SYSTEMATIC) return fail(c, TSDFLOC_E_BAD_ARG, "unknown resampling method");
  if (method == TSDFLOC_RESAMPLE_RESIDUAL && !draw) return fail(c, TSDFLOC_E_BAD_ARG, "the residual resampler needs an index draw callback");
  DeviceGuard guard(c->device);
  int rc;
  std::vector<float> w;
  if ((rc = weights_for_host_half(c, particles, n, w))) return rc;
  std::vector<uint32_t> counts, run_parent;
  uint64_t n_runs = 0;
  if (method == TSDFLOC_RESAMPLE_RESIDUAL_SYSTEMATIC)
  {
    counts.resize(n);
    uint64_t total = 0;
    if (tsdfloc_residual_systematic_counts(w.data(), 1, n, u, counts.data(), &total) != TSDFLOC_OK)
      return fail(c, TSDFLOC_E_BAD_ARG, "residual-systematic resampling: negative or non-finite weight");
    n_runs = n;
    return tsdfloc_resample_expand(c, nullptr, counts.data(), n_runs, particles_out, cap, n_out, parents);
  }
  counts.resize(n);
  run_parent.resize(n);   // every run emits >= 1 of the n output particles
  rc = tsdfloc_residual_runs(w.data(), 1, n, draw, user, 64ull * n + 1024ull, run_parent.data(), counts.data(), n, &n_runs, nullptr);
  if (rc == TSDFLOC_E_CAPACITY) return fail(c, TSDFLOC_E_NO_VALID_PARTICLE, "residual resampling: the weights do not fill the output (all zero?)");
  if (rc != TSDFLOC_OK) return fail(c, rc, "residual resampling: bad index draw");
  return tsdfloc_resample_expand(c, run_parent.data(), counts.data(), n_runs, particles_out, cap, n_out, parents);
}

// Wheel / Metropolis / Rejection (src/mcl_3d.cpp:243-263 cases 0, 4, default): the host half picks one parent per output slot
// from the caller's draws (host_resample.cpp), the device copies the particles — runs of length one through k_expand_runs.
int tsdfloc_resample_drawn(tsdfloc_ctx* c, int method, const float* particles, uint64_t n, const tsdfloc_draws* draws, float* particles_out,
                           uint64_t cap, uint64_t* n_out, uint32_t* parents)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!particles_out || !n_out || !draws) return fail(c, TSDFLOC_E_BAD_ARG, "NULL argument");
  if (method != TSDFLOC_RESAMPLE_WHEEL && method != TSDFLOC_RESAMPLE_METROPOLIS && method != TSDFLOC_RESAMPLE_REJECTION)
    return fail(c, TSDFLOC_E_BAD_ARG, "tsdfloc_resample_drawn serves the Wheel, Metropolis and Rejection resamplers (use tsdfloc_resample for the others)");
  if (!draws->real) return fail(c, TSDFLOC_E_BAD_ARG, "the resampler needs a uniform_real draw callback");
  if (method != TSDFLOC_RESAMPLE_WHEEL && !draws->index) return fail(c, TSDFLOC_E_BAD_ARG, "the resampler needs an index draw callback");
  DeviceGuard guard(c->device);
  int rc;
  std::vector<float> w;
  if ((rc = weights_for_host_half(c, particles, n, w))) return rc;
  std::vector<uint32_t> parent(n), ones(n, 1u);
  if (method == TSDFLOC_RESAMPLE_WHEEL)
    rc = tsdfloc_wheel_parents(w.data(), 1, n, draws->real, draws->user, parent.data());
  else if (method == TSDFLOC_RESAMPLE_METROPOLIS)
    rc = tsdfloc_metropolis_parents(w.data(), 1, n, draws->metropolis_steps, draws->real, draws->index, draws->user, parent.data());
  else
    rc = tsdfloc_rejection_parents(w.data(), 1, n, draws->real, draws->index, draws->user, draws->max_draws, parent.data(), nullptr);
  if (rc == TSDFLOC_E_CAPACITY) return fail(c, TSDFLOC_E_CAPACITY, "rejection resampling: max_draws index draws did not fill the output");
  if (rc != TSDFLOC_OK) return fail(c, rc, "resampling: bad index draw");
  return tsdfloc_resample_expand(c, parent.data(), ones.data(), n, particles_out, cap, n_out, parents);
}

int tsdfloc_debug_eval(tsdfloc_ctx* c, const float* particles, uint64_t n, const float* points, uint64_t p, const float tf[16],
                       uint32_t* idx, uint32_t* hits, float* raw_weights)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  if (!particles || !points || !tf || n == 0 || p == 0) return fail(c, TSDFLOC_E_BAD_ARG, "NULL or empty argument");
  if (idx && n * p >= (1ull << 32)) return fail(c, TSDFLOC_E_BAD_ARG, "index dump limited to n*p < 2^32 pairs");
  DeviceGuard guard(c->device);
  cudaStream_t s = c->stream;
  int rc;
  if ((rc = ensure(c, c->d_particles, sizeof(float) * 7 * n, "cudaMalloc(particles)"))) return rc;
  if ((rc = ensure(c, c->d_raw, sizeof(float) * n, "cudaMalloc(raw weights)"))) return rc;
  if ((rc = ensure(c, c->d_xyz_stage, sizeof(float) * 3 * (p + 1), "cudaMalloc(scan staging)"))) return rc;
  if (idx && (rc = ensure(c, c->d_idx, sizeof(uint32_t) * n * p, "cudaMalloc(index dump)"))) return rc;
  if ((rc = ensure(c, c->d_hits, sizeof(uint32_t) * n, "cudaMalloc(hit counts)"))) return rc;
  c->have_cdf = false;
  c->n_resident = 0;
  CU_TRY(c, cudaMemcpyAsync(c->d_xyz_stage.p, points, sizeof(float) * 3 * p, cudaMemcpyHostToDevice, s), "H2D scan");
  if ((rc = stage_prep_scan(c, static_cast<const float*>(c->d_xyz_stage.p), p, s))) return rc;
  float* d_p = static_cast<float*>(c->d_particles.p);
  CU_TRY(c, cudaMemcpyAsync(d_p, particles, sizeof(float) * 7 * n, cudaMemcpyHostToDevice, s), "H2D particles");
  CU_TRY(c, cudaMemsetAsync(c->d_hits.p, 0, sizeof(uint32_t) * n, s), "memset hits");
  {
    // the index-recording instantiation of the production kernel itself (same pairing, same quotient mode), identity order
    const int saved = c->sort_mode;
    c->sort_mode = 0;
    rc = stage_eval(c, d_p, n, 0, n, tf, static_cast<float*>(c->d_raw.p), s, nullptr, 0, true,
                    idx ? static_cast<uint32_t*>(c->d_idx.p) : nullptr, static_cast<uint32_t*>(c->d_hits.p));
    c->sort_mode = saved;
    if (rc) return rc;
  }
  if (idx) CU_TRY(c, cudaMemcpyAsync(idx, c->d_idx.p, sizeof(uint32_t) * n * p, cudaMemcpyDeviceToHost, s), "D2H index dump");
  if (hits) CU_TRY(c, cudaMemcpyAsync(hits, c->d_hits.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s), "D2H hits");
  if (raw_weights) CU_TRY(c, cudaMemcpyAsync(raw_weights, c->d_raw.p, sizeof(float) * n, cudaMemcpyDeviceToHost, s), "D2H raw weights");
  CU_TRY(c, cudaStreamSynchronize(s), "stream sync");
  return TSDFLOC_OK;
}

// Measurement probe: random 4 B gathers over the first `bytes` of the voxel array (L2-resident when bytes <= ~100 MB).
// spread_sectors == 0: every lane its own random word; otherwise the lanes of a request share `spread_sectors` consecutive
// 32 B sectors. Reports the average kernel time over `reps` launches after one warm-up and the gathers per launch.
int tsdfloc_probe_gather(tsdfloc_ctx* c, uint64_t bytes, uint32_t spread_sectors, uint32_t reps, float* ms_per_launch, uint64_t* gathers_per_launch)
{
  if (!c || !ms_per_launch || !gathers_per_launch || reps == 0) return TSDFLOC_E_BAD_ARG;
  DeviceGuard guard(c->device);
  const uint64_t have = static_cast<uint64_t>(c->map.data_size) * 4ull;
  if (bytes == 0 || bytes > have) bytes = have;
  const uint32_t n_words = static_cast<uint32_t>(bytes / 4);
  if (n_words < 4096u || spread_sectors > 32u) return fail(c, TSDFLOC_E_BAD_ARG, "probe needs >= 16 KB of map data and spread <= 32");
  const uint32_t rounds = 64, grid = static_cast<uint32_t>(c->sm_count) * 8u * 16u;
  cudaStream_t s = c->stream;
  cudaEvent_t e0, e1;
  CU_TRY(c, cudaEventCreate(&e0), "event create");
  CU_TRY(c, cudaEventCreate(&e1), "event create");
  k_probe_gather<<<grid, 256, 0, s>>>(c->d_voxels, n_words, rounds, spread_sectors, c->d_mean);
  int rc = launch_check(c, "k_probe_gather");
  if (rc == TSDFLOC_OK)
  {
    cudaEventRecord(e0, s);
    for (uint32_t r = 0; r < reps; ++r)
    {
      k_probe_gather<<<grid, 256, 0, s>>>(c->d_voxels, n_words, rounds, spread_sectors, c->d_mean);
      ++c->launches;
    }
    cudaEventRecord(e1, s);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e != cudaSuccess) rc = fail(c, TSDFLOC_E_CUDA, std::string("probe: ") + cudaGetErrorString(e));
    *ms_per_launch = ms / static_cast<float>(reps);
    *gathers_per_launch = static_cast<uint64_t>(grid) * 256ull * rounds * 8ull;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

int tsdfloc_tune(tsdfloc_ctx* c, int knob, int value)
{
  if (!c) return TSDFLOC_E_BAD_ARG;
  switch (knob)
  {
    case TSDFLOC_TUNE_SPATIAL_ORDER:
      if (value < -1 || value > 1) return fail(c, TSDFLOC_E_BAD_ARG, "spatial order: -1 automatic, 0 off, 1 on");
      c->sort_mode = value;
      return TSDFLOC_OK;
    case TSDFLOC_TUNE_EVAL_PAIRING:
      if (value < 0 || value > 2) return fail(c, TSDFLOC_E_BAD_ARG, "pairing: 0 automatic, 1 particle pairs, 2 point pairs");
      c->tune_shape = value;
      return TSDFLOC_OK;
    case TSDFLOC_TUNE_DIVISION:
      if (value < -1 || value > kDivBracket) return fail(c, TSDFLOC_E_BAD_ARG, "division: -1 automatic, 0 IEEE, 1 three-instruction, 2 bracket");
      c->tune_div = value;
      return TSDFLOC_OK;
    case TSDFLOC_TUNE_EVAL_REGISTERS:
      if (value < 0 || value > 2) return fail(c, TSDFLOC_E_BAD_ARG, "registers: 0 automatic, 1 64 per thread, 2 128 per thread");
      c->tune_regs = value;
      return TSDFLOC_OK;
    case TSDFLOC_TUNE_EVAL_CHUNKS:
      if (value < 0 || value > 64) return fail(c, TSDFLOC_E_BAD_ARG, "chunks: 0 automatic, 1..64 scan chunks per particle");
      c->tune_chunks = value;
      return TSDFLOC_OK;
    case TSDFLOC_TUNE_STAGE_TIMERS:
      c->stage_timers = value != 0;
      c->stage_seen = 0;
      return TSDFLOC_OK;
    case TSDFLOC_TUNE_GRAPHS:
      c->graphs_on = value != 0;
      if (!c->graphs_on)
      {
        DeviceGuard guard(c->device);
        cudaStreamSynchronize(c->stream);
        graph_drop_all(c);
      }
      return TSDFLOC_OK;
    default: return fail(c, TSDFLOC_E_BAD_ARG, "unknown tuning knob");
  }
}

int tsdfloc_stage_times(tsdfloc_ctx* c, float ms[4])
{
  if (!c || !ms) return TSDFLOC_E_BAD_ARG;
  if (!c->stage_timers) return fail(c, TSDFLOC_E_STATE, "stage timers are off: tsdfloc_tune(ctx, TSDFLOC_TUNE_STAGE_TIMERS, 1)");
  DeviceGuard guard(c->device);
  CU_TRY(c, cudaStreamSynchronize(c->stream), "stream sync");
  CU_TRY(c, cudaDeviceSynchronize(), "device sync");
  auto span = [&](int a, int b, float* out) {
    *out = 0.0f;
    if ((c->stage_seen >> a & 1u) && (c->stage_seen >> b & 1u)) cudaEventElapsedTime(out, c->ev_stage[a], c->ev_stage[b]);
  };
  float prep = 0.0f, mats = 0.0f;
  span(tsdfloc_ctx::kEvPrep0, tsdfloc_ctx::kEvPrep1, &prep);
  if ((c->stage_seen >> tsdfloc_ctx::kEvInit0 & 1u) && c->eval_timed) cudaEventElapsedTime(&mats, c->ev_stage[tsdfloc_ctx::kEvInit0], c->ev_eval0);
  ms[0] = prep + mats;                                                       // init_kernel: scan preparation + spatial order + matrices
  ms[1] = 0.0f;
  if (c->eval_timed) cudaEventElapsedTime(&ms[1], c->ev_eval0, c->ev_eval1);  // exec_kernel: k_eval
  span(tsdfloc_ctx::kEvNorm0, tsdfloc_ctx::kEvNorm1, &ms[2]);                  // weight_update: K2 (+ K3)
  span(tsdfloc_ctx::kEvDraw0, tsdfloc_ctx::kEvDraw1, &ms[3]);                  // resampling: K4
  cudaGetLastError();
  return TSDFLOC_OK;
}

int tsdfloc_graph_stats(const tsdfloc_ctx* c, uint64_t out[2], const char** note)
{
  if (!c || !out) return TSDFLOC_E_BAD_ARG;
  out[0] = c->graph_captures;
  out[1] = c->graph_replays;
  if (note) *note = c->graph_note.c_str();
  return TSDFLOC_OK;
}

int tsdfloc_last_cdf_was_exact(const tsdfloc_ctx* c) { return (c && c->h_status) ? (c->h_status->inexact ? 0 : 1) : -1; }

int tsdfloc_division_mode(tsdfloc_ctx* c, uint64_t* open_brackets)
{
  if (!c) return -1;
  if (open_brackets) *open_brackets = c->bracket_open;
  return c->map.div_mode;
}

int tsdfloc_eval_stats(tsdfloc_ctx* c, uint64_t out[4])
{
  if (!c || !out) return TSDFLOC_E_BAD_ARG;
  DeviceGuard guard(c->device);
  CU_TRY(c, cudaDeviceSynchronize(), "device sync");
  unsigned long long h[4];
  CU_TRY(c, cudaMemcpy(h, c->d_eval_stats, sizeof(h), cudaMemcpyDeviceToHost), "D2H eval stats");
  for (int i = 0; i < 4; ++i) out[i] = h[i];
  return TSDFLOC_OK;
}

int tsdfloc_last_eval_ms(tsdfloc_ctx* c, float* ms)
{
  if (!c || !ms) return TSDFLOC_E_BAD_ARG;
  if (!c->eval_timed) return fail(c, TSDFLOC_E_STATE, "no evaluation kernel has been launched yet");
  DeviceGuard guard(c->device);
  CU_TRY(c, cudaEventSynchronize(c->ev_eval1), "event sync");
  CU_TRY(c, cudaEventElapsedTime(ms, c->ev_eval0, c->ev_eval1), "event elapsed");
  return TSDFLOC_OK;
}

// Host-only test hook (no GPU needed): the U recurrence table evaluated on the host, for unit tests of the
// __host__ __device__ table builder against the plain loop. Writes min(count, cap) values; returns count.
uint64_t tsdfloc_host_u_sequence(float u0, uint64_t n, double limit, float* out, uint64_t cap, uint32_t* n_segs, uint32_t* flags)
{
  std::vector<USeg> segs(kMaxUSegs);
  unsigned long long below = 0;
  uint32_t fl = 0;
  const uint32_t ns = build_u_table(u0, 1.0 / static_cast<double>(n), limit, segs.data(), kMaxUSegs, &below, 2ull * n + 64ull, &fl);
  if (n_segs) *n_segs = ns;
  if (flags) *flags = fl;
  for (uint64_t j = 0; j < below && j < cap; ++j) out[j] = u_at(segs.data(), ns, j);
  return below;
}

}  // extern "C"

#include "tsdfloc_multi.inc"
#include "tsdfloc_host_map.h"
#include "tsdfloc_ingest.inc"
#include "tsdfloc_mcl.inc"
