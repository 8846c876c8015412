/* include/tsdfloc.h — C ABI of the B200-native MCL sensor update (libtsdfloc.so).
 *
 * This is the drop-in boundary for ONE path of uos/tsdf_localization: particle poses + a LiDAR scan in,
 * per-particle TSDF-likelihood weights, the weighted mean pose and systematically resampled particles out.
 * Plain pointers and sizes only; no C++/torch types. Every entry point names the reference interface it
 * replaces (paths relative to the reference repo).
 *
 * Two layers:
 *   (A) host-buffer calls — what the reference's own call sites bind (INTEGRATION.md shows the C++ shim):
 *         tsdfloc_sensor_update        <- CudaEvaluator::evaluate(vector<Particle>&, vector<CudaPoint>&, tf[16])
 *                                         include/tsdf_localization/cuda/cuda_evaluator.h:119, src/cuda/cuda_evaluator.cu:118-428
 *         tsdfloc_resample_systematic  <- SystematicResampler::resample(ParticleCloud&)
 *                                         include/tsdf_localization/resampling/novel_resampling.h:38-74
 *   (B) device-pointer stage calls — the same kernels on caller-owned device memory and a caller stream, so
 *       a multi-GPU driver (one process per GPU) can put NCCL all-gathers between the stages.
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device and returns TSDFLOC_E_CUDA with a
 * message if the device or the kernel image (sm_100a) is unavailable.
 *
 * Threading: calls on one ctx are not re-entrant; use one ctx per device / per caller thread.
 */
#ifndef TSDFLOC_H
#define TSDFLOC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSDFLOC_ABI_VERSION 3

typedef struct tsdfloc_ctx tsdfloc_ctx;

/* Status codes. The reference signals the same conditions with std::runtime_error texts
 * (src/cuda/cuda_util.cu:8-17; "No particle is valid!" src/cuda/cuda_evaluator.cu:366-369). */
enum tsdfloc_status
{
  TSDFLOC_OK = 0,
  TSDFLOC_E_BAD_ARG = 1,
  TSDFLOC_E_CUDA = 2,
  TSDFLOC_E_NO_VALID_PARTICLE = 3, /* sum of weights == 0 */
  TSDFLOC_E_EMPTY_SCAN = 4,        /* P == 0: reference returns a default pose and leaves weights untouched (cuda_evaluator.cu:122-125) */
  TSDFLOC_E_CAPACITY = 5,          /* resampling produced more particles than the output capacity */
  TSDFLOC_E_STATE = 6              /* stage called out of order (e.g. resample before sensor_update) */
};

/* Geometry of the two-level sparse voxel map; field-for-field CudaSubVoxelMap<float,float>::MapCoef
 * (include/tsdf_localization/cuda/cuda_sub_voxel_map.h:22-50) with fixed-width integers. */
typedef struct tsdfloc_map_desc
{
  uint64_t dim[3];        /* ceil(|max-min| / resolution) per axis                 */
  float min[3];
  float max[3];
  float resolution;
  float init_value;       /* value returned for unmapped space                     */
  uint64_t up_dim[3];     /* ceil(|max-min| / 1 m) per axis                        */
  uint64_t up_dim_2;      /* up_dim[0] * up_dim[1]                                 */
  uint64_t sub_dim;       /* ceil(1 m / resolution)                                */
  uint64_t sub_dim_2;
  uint64_t grid_occ_size; /* up_dim[0]*up_dim[1]*up_dim[2]                         */
  uint64_t data_size;     /* number of fp32 voxels in data (bricks * sub_dim^3)    */
} tsdfloc_map_desc;

/* Sensor model; constructor arguments of CudaEvaluator (cuda_evaluator.h:112) / TSDFEvaluator (tsdf_evaluator.h:72). */
typedef struct tsdfloc_params
{
  float a_hit;     /* default 0.9  (util.h:16) */
  float a_range;   /* default 0.1  (util.h:17) */
  float a_max;     /* default 0.0  (util.h:18) */
  float max_range; /* default 100  (util.h:13); the range term uses 1/max_range like the CPU evaluator (tsdf_evaluator.cpp:60) */
  int32_t per_point; /* accepted for signature parity; both reference variants compute the same weights, one kernel serves both */
  int32_t neg_policy; /* enum tsdfloc_neg_policy: lookups below map.min on an axis */
} tsdfloc_params;

/* Lookups whose offset x - map.min is negative on some axis. The reference converts the negative float to unsigned, which
 * is undefined behaviour: its CUDA build (cuda_eval_particles.h:14-16,34-36,62-64; cvt.rzi.u32.f32 saturates) clamps them
 * onto the min face, its x86 CPU build (cuda_sub_voxel_map.tcc:50-137) wraps the flat index into a neighbouring brick.
 *   TSDFLOC_NEG_MISS                   such a lookup is a miss (init_value), like every other lookup outside the map (default)
 *   TSDFLOC_NEG_SATURATE_LIKE_REF_GPU  bit-for-bit what the reference's CUDA evaluator does (a drop-in for that evaluator
 *                                      reproduces it, indices and weights included)                                        */
enum tsdfloc_neg_policy
{
  TSDFLOC_NEG_MISS = 0,
  TSDFLOC_NEG_SATURATE_LIKE_REF_GPU = 1
};

void tsdfloc_default_params(tsdfloc_params* p);
int tsdfloc_abi_version(void);
const char* tsdfloc_status_string(int status);
/* Message of the last failing call on this ctx (or of the last failing tsdfloc_create when ctx == NULL). */
const char* tsdfloc_last_error(const tsdfloc_ctx* ctx);

/* ---- host-side sparse map builder -----------------------------------------------------------------------------
 * Stand-alone replacement for the host half of CudaSubVoxelMap<float,float>
 * (include/tsdf_localization/cuda/cuda_sub_voxel_map.{h,tcc}): constructor geometry (tcc:4-48) and setData's brick
 * allocation in increasing upper-cell order (tcc:170-230), producing the grid_occ / data arrays tsdfloc_create
 * uploads. A caller that already owns a reference CudaSubVoxelMap passes coef()/rawGridOcc()/rawData() to
 * tsdfloc_create instead and never needs these. Pure host code, no GPU required. */
typedef struct tsdfloc_host_map tsdfloc_host_map;
int tsdfloc_map_create(const float min[3], const float max[3], float resolution, float init_value, tsdfloc_host_map** out);
/* cells: n x (x, y, z, value) fp32. Cells outside [min, max) on any axis are an error (the reference is undefined there). */
int tsdfloc_map_set_data(tsdfloc_host_map* m, const float* cells, uint64_t n);
const tsdfloc_map_desc* tsdfloc_map_get_desc(const tsdfloc_host_map* m);
const int32_t* tsdfloc_map_grid_occ(const tsdfloc_host_map* m);
const float* tsdfloc_map_data(const tsdfloc_host_map* m);
void tsdfloc_map_destroy(tsdfloc_host_map* m);
/* Map ingest — createTSDFMap without its HDF5 layer (include/tsdf_localization/map/map_util.h:17-154). chunk_pos: n_chunks x
 * (cx, cy, cz), the integers of the dataset names "/map/<cx>_<cy>_<cz>"; chunk_data: n_chunks x 64^3 raw TSDFValue words
 * ({int16 value_mm, int16 weight}, util/tsdf.h:11-87) indexed 64*64*i + 64*j + k (:106). Bounding box, voxel corner
 * positions, the +-600 mm truncation, the likelihood^3 transform (sigma; the node uses 0.1) and setData are the reference's,
 * bit for bit; voxels with weight != 0 outside the truncation band become the free-space points global localisation
 * samples from (:131-145; particle_cloud.cpp:105-148), in the reference's order (datasets in increasing name order). */
int tsdfloc_map_from_chunks(const int32_t* chunk_pos, const uint32_t* chunk_data, uint64_t n_chunks, float sigma, tsdfloc_host_map** out);
/* The same ingest on the GPU (`device`): the raw chunk words go to the device once, one thread per voxel marks the touched
 * 1 m cells, a scan allocates the bricks in the reference's order and a second pass scatters the LUT values and compacts the
 * free-space points in the reference's order. The result — host arrays, like the reference's map object — is bit-identical to
 * tsdfloc_map_from_chunks (incl. the reference's last-writer-wins resolution of voxels whose fp32 cell selection collides);
 * message of a failure via tsdfloc_last_error(NULL). Fewer than 16,384 chunks (16 GB of words). */
int tsdfloc_map_from_chunks_gpu(const int32_t* chunk_pos, const uint32_t* chunk_data, uint64_t n_chunks, float sigma, int device,
                                tsdfloc_host_map** out);
/* createTSDFMap + the CudaEvaluator constructor in one step, on the device: the chunk words go to `device`, the bricks are
 * built there and become the evaluation context's voxel array WITHOUT visiting the host (only the brick table, 4 B per 1 m
 * cell, is read back to be laid out as the padded table). The free-space points stay on the device too: tsdfloc_init_particles
 * with TSDFLOC_INIT_FREE_MAP and free_map == NULL samples from them. Same voxels, bit for bit, as tsdfloc_map_from_chunks ->
 * tsdfloc_create. Replaces map_util.h:17-154 followed by src/cuda/cuda_evaluator.cu:21-59. */
int tsdfloc_create_from_chunks(const int32_t* chunk_pos, const uint32_t* chunk_data, uint64_t n_chunks, float sigma,
                               const tsdfloc_params* params, int device, tsdfloc_ctx** out);
/* Geometry of the map a ctx evaluates against (coef() of the reference's map object). */
int tsdfloc_map_desc_of(const tsdfloc_ctx* ctx, tsdfloc_map_desc* desc);
/* Device pointer to / number of the free-space points a ctx created by tsdfloc_create_from_chunks keeps (NULL / 0 otherwise). */
int tsdfloc_free_map_device(const tsdfloc_ctx* ctx, const float** d_points, uint64_t* n_points);
/* The free-space points of a map built by tsdfloc_map_from_chunks[_gpu]: *n points x 3 fp32 (NULL / 0 for other maps). */
const float* tsdfloc_map_free_points(const tsdfloc_host_map* m, uint64_t* n);
/* TSDF (mm) -> likelihood^3 LUT value and the value for unmapped space, createTSDFMap's transform
 * (include/tsdf_localization/map/map_util.h:68-71, 124-126). */
float tsdfloc_likelihood_value(float tsdf_mm, float sigma);
float tsdfloc_likelihood_init(float sigma);

/* ---- lifetime -------------------------------------------------------------------------------------------
 * Replaces CudaEvaluator::CudaEvaluator(map, per_point, a_hit, a_range, a_max, max_range)
 * (src/cuda/cuda_evaluator.cu:21-59): deep-copies the map's host arrays to the device (grid_occ = brick table,
 * data = bricks). Host arrays are only read during the call. No process-wide globals: any number of ctx. */
int tsdfloc_create(const tsdfloc_map_desc* map, const int32_t* grid_occ, const float* data, const tsdfloc_params* params,
                   int device, tsdfloc_ctx** out);
void tsdfloc_destroy(tsdfloc_ctx* ctx);

/* ---- (A) host-buffer calls ------------------------------------------------------------------------------ */

/* One sensor update on host buffers.
 *   particles : n x 7 fp32, the reference's `Particle` layout (x y z roll pitch yaw weight; particle.h:50-53).
 *               Poses are read; slot [6] of every particle is overwritten with the NORMALISED weight
 *               (cuda_eval_particles.h:556, cuda_evaluator.cu:405-408). Order is preserved.
 *   points    : p x 3 fp32 sensor-frame points, the reference's `CudaPoint` (cuda_evaluator.h:22-39).
 *   tf        : 16 fp32 row-major scanner->robot matrix; rows 0-2 are used (tsdf_evaluator.cpp:132-145).
 *   mean_pose : out, x y z roll pitch yaw of the weighted mean (atan2 of weighted sin/cos sums, cuda_evaluator.cu:387-392).
 * The particle set stays resident on the device for tsdfloc_resample_systematic. */
int tsdfloc_sensor_update(tsdfloc_ctx* ctx, float* particles, uint64_t n, const float* points, uint64_t p, const float tf[16],
                          float mean_pose[6]);

/* Systematic resampling of the particle set left on the device by tsdfloc_sensor_update, with the random
 * offset u0 in [0, 1/n) supplied by the caller (the reference draws it from std::mt19937, novel_resampling.h:49).
 * Reproduces the reference's fp32 running-U / fp64 running-sum recurrence exactly, including output lengths
 * != n. particles_out: capacity `cap` x 7 fp32; copies keep the parent's normalised weight. parents (optional):
 * index of the source particle per output slot. Only entries [0, *n_out) are meaningful; entries up to min(cap, n) may be
 * written. Page-locked output buffers (cudaMallocHost / cudaHostRegister) are filled by the copy engine directly, pageable
 * ones through the ctx's staging buffer; the same holds for the input buffers of tsdfloc_sensor_update. */
int tsdfloc_resample_systematic(tsdfloc_ctx* ctx, float u0, float* particles_out, uint64_t cap, uint64_t* n_out,
                                uint32_t* parents);

/* Systematic resampling of a weighted particle set in host memory — the drop-in for
 * SystematicResampler::resample(ParticleCloud&) (novel_resampling.h:41-72) as mcl_3d calls it (src/mcl_3d.cpp:424-426),
 * independent of any preceding sensor update. particles: n x 7 fp32 whose slot [6] holds the (normalised) weights; they
 * are used as they are, exactly like the reference's running sum `s += particle.second`. Outputs as above. */
int tsdfloc_resample_particles(tsdfloc_ctx* ctx, const float* particles, uint64_t n, float u0, float* particles_out, uint64_t cap,
                               uint64_t* n_out, uint32_t* parents);

/* Parity/debug: per-(particle, point) flat voxel index (data_size = miss) and per-particle hit counts for the
 * given inputs, computed by the same device index function the evaluation kernel uses.
 * idx (optional): n*p uint32, particle-major. hits (optional): n uint32. raw_weights (optional): n un-normalised. */
int tsdfloc_debug_eval(tsdfloc_ctx* ctx, const float* particles, uint64_t n, const float* points, uint64_t p, const float tf[16],
                       uint32_t* idx, uint32_t* hits, float* raw_weights);

/* ---- .mcl snapshot files ---------------------------------------------------------------------------------------
 * MCLFile::read / MCLFile::write (include/tsdf_localization/util/mcl_file.h:21-32, src/util/mcl_file.cpp:14-113): the text
 * fixture format of one scan (points + rings), a particle set, the scanner->robot transform and a reference pose
 * (x y z q1 q2 q3 q4). Written files are byte-identical to the reference's. Failures: TSDFLOC_E_STATE when the file cannot
 * be opened / written, TSDFLOC_E_BAD_ARG when its content does not parse; text via tsdfloc_last_error(NULL) (the reference's
 * exception messages). Pure host code. */
typedef struct tsdfloc_mcl tsdfloc_mcl;
int tsdfloc_mcl_read(const char* path, tsdfloc_mcl** out);
void tsdfloc_mcl_free(tsdfloc_mcl* m);
uint64_t tsdfloc_mcl_n_points(const tsdfloc_mcl* m);
uint64_t tsdfloc_mcl_n_particles(const tsdfloc_mcl* m);
const float* tsdfloc_mcl_points(const tsdfloc_mcl* m);      /* n_points x 3 */
const int32_t* tsdfloc_mcl_rings(const tsdfloc_mcl* m);     /* n_points */
const float* tsdfloc_mcl_particles(const tsdfloc_mcl* m);   /* n_particles x 7 */
const float* tsdfloc_mcl_tf(const tsdfloc_mcl* m);          /* 16 */
const float* tsdfloc_mcl_pose(const tsdfloc_mcl* m);        /* 7 */
int tsdfloc_mcl_write(const char* path, const float* points, const int32_t* rings, uint64_t n_points, const float* particles,
                      uint64_t n_particles, const float tf[16], const float pose7[7]);

/* ---- the reference's other resamplers --------------------------------------------------------------------------
 * mcl_3d selects its resampler at run time (src/mcl_3d.cpp:243-263); the compiled-in default is ResidualSystematic (:765),
 * the dynamic-reconfigure default Residual (cfg/MCL.cfg:54). Both are SEQUENTIAL recurrences over the weights with no exact
 * parallel form, so they are split: the recurrence runs on the host over 4 B per particle and yields RUNS (how many copies
 * of which particle follow each other in the output); the device expands the runs, so the 28 B particles stay where the
 * sensor update left them. Parents are identical to the reference's given the same weights and the same random draws. */
enum tsdfloc_resample_method
{
  TSDFLOC_RESAMPLE_SYSTEMATIC = 0,          /* SystematicResampler          novel_resampling.h:38-74   */
  TSDFLOC_RESAMPLE_RESIDUAL = 1,            /* ResidualResampler            novel_resampling.h:9-36    */
  TSDFLOC_RESAMPLE_RESIDUAL_SYSTEMATIC = 2  /* ResidualSystematicResampler  novel_resampling.h:76-104  */
};
/* One draw of the reference's std::uniform_int_distribution<size_t>(0, n - 1)(*m_generator_ptr) (novel_resampling.h:14,21). */
typedef uint64_t (*tsdfloc_index_draw_fn)(void* user);

/* Host half of ResidualSystematicResampler::resample (:86-99): counts[m] = copies of particle m, from the fp32 remainder
 * recurrence started at u0 (the reference's uniform_real_distribution<float>(0, 1) draw). weights[m * stride]; *total = sum
 * of the counts (about n). Pure host code. TSDFLOC_E_BAD_ARG for a negative / NaN weight (undefined in the reference). */
int tsdfloc_residual_systematic_counts(const float* weights, uint64_t stride, uint64_t n, float u0, uint32_t* counts, uint64_t* total);
/* Host half of ResidualResampler::resample (:14-31): calls draw() until n output slots are filled; run k copies particle
 * run_parent[k] run_count[k] times (ceil(w * n), cut to the slots left). run_cap >= n always suffices. max_draws bounds the
 * loop (the reference spins forever on all-zero weights): TSDFLOC_E_CAPACITY when it is hit. Pure host code. */
int tsdfloc_residual_runs(const float* weights, uint64_t stride, uint64_t n, tsdfloc_index_draw_fn draw, void* user, uint64_t max_draws,
                          uint32_t* run_parent, uint32_t* run_count, uint64_t run_cap, uint64_t* n_runs, uint64_t* n_draws);
/* Device half: expands runs over the particle set left on the device by tsdfloc_sensor_update / tsdfloc_resample*.
 * run_parent == NULL: run r copies particle r (ResidualSystematic). Outputs as tsdfloc_resample_systematic. */
int tsdfloc_resample_expand(tsdfloc_ctx* ctx, const uint32_t* run_parent, const uint32_t* run_count, uint64_t n_runs, float* particles_out,
                            uint64_t cap, uint64_t* n_out, uint32_t* parents);
/* The same kernel on caller-owned device memory (multi-GPU drivers): d_run_off = n_runs + 1 exclusive prefix sums of the run
 * counts; output slots [first_out, first_out + count_out) go to d_particles_out (and the same slice of every peer buffer). */
int tsdfloc_resample_expand_device(tsdfloc_ctx* ctx, const float* d_particles, const uint32_t* d_run_off, const uint32_t* d_run_parent,
                                   uint64_t n_runs, uint64_t first_out, uint64_t count_out, float* d_particles_out,
                                   float* const* d_out_peers, uint32_t n_peers, uint32_t* d_parents, void* stream);
/* Resampler::resample(ParticleCloud&) for any of the three methods (resampling/resampler.h:26). particles != NULL: a weighted
 * cloud in host memory (n x 7 fp32); particles == NULL: the set the last sensor update left on the device (n ignored).
 * u: the method's uniform_real draw (Systematic: in [0, 1/n); ResidualSystematic: in [0, 1); unused by Residual);
 * draw/user: the index draws of Residual (NULL otherwise). */
int tsdfloc_resample(tsdfloc_ctx* ctx, int method, const float* particles, uint64_t n, float u, tsdfloc_index_draw_fn draw, void* user,
                     float* particles_out, uint64_t cap, uint64_t* n_out, uint32_t* parents);

/* ---- the three remaining choices of mcl_3d's resampling_method switch (src/mcl_3d.cpp:243-263) ---------------------
 * Wheel (case 0), Metropolis (case 4) and Rejection (default) pick every output particle from random draws on the resampler's
 * std::mt19937. The draws stay with the caller (callbacks, like the index draws of Residual) so that equal seeds give equal
 * outputs; the host half turns weights + draws into one parent per output slot (4 B per particle on the host), the device
 * copies the 28 B particles (k_expand_runs with runs of length one). */
enum tsdfloc_resample_method_drawn
{
  TSDFLOC_RESAMPLE_WHEEL = 3,       /* WheelResampler       src/resampling/wheel_resampler.cpp:6-34 */
  TSDFLOC_RESAMPLE_METROPOLIS = 4,  /* MetropolisResampler  novel_resampling.h:106-144              */
  TSDFLOC_RESAMPLE_REJECTION = 5    /* RejectionResampler   novel_resampling.h:146-189              */
};
/* One uniform_real draw on [0, 1) already converted to the method's FLOAT_T: Metropolis / Rejection draw
 * std::uniform_real_distribution<FLOAT_T>(0.0, 1.0) (novel_resampling.h:115, 151); Wheel draws
 * std::uniform_real_distribution<>(0.0, 1.0) — a double — and narrows it to FLOAT_T (wheel_resampler.cpp:9, 15). */
typedef float (*tsdfloc_real_draw_fn)(void* user);
typedef struct tsdfloc_draws
{
  tsdfloc_real_draw_fn real;    /* all three methods */
  tsdfloc_index_draw_fn index;  /* Metropolis, Rejection: std::uniform_int_distribution<size_t>(0, n - 1); NULL for Wheel */
  void* user;                   /* handed to both callbacks */
  uint64_t metropolis_steps;    /* MetropolisResampler's sampling_steps_ (mcl_3d passes 50, src/mcl_3d.cpp:258) */
  uint64_t max_draws;           /* Rejection: give up (TSDFLOC_E_CAPACITY) after this many index draws; 0 = never, like the reference */
} tsdfloc_draws;

/* Host half of WheelResampler::resample: per output slot one draw u and the first index whose fp32 running weight sum (restarted
 * at 0 for every slot in the reference, hence the same prefix array for all) reaches u; a slot whose u exceeds the last sum
 * keeps its own particle (wheel_resampler.cpp:13-31 leaves particle_cloud[particle_index] untouched). O(n) instead of the
 * reference's O(n^2): prefix once, guide table, bracketed search. Any finite weights are accepted (a running maximum makes
 * "first index whose sum reaches u" searchable even where negative weights make the sums non-monotone). Pure host code. */
int tsdfloc_wheel_parents(const float* weights, uint64_t stride, uint64_t n, tsdfloc_real_draw_fn real, void* user, uint32_t* parents);
/* Host half of MetropolisResampler::resample: per output slot `steps` rounds of (u, j) draws, k = j whenever
 * u <= w[j] / w[0] — the reference binds `particle_k` to particle 0 once (:125) and never rebinds it, so every ratio is
 * against particle 0; reproduced as written. Pure host code. */
int tsdfloc_metropolis_parents(const float* weights, uint64_t stride, uint64_t n, uint64_t steps, tsdfloc_real_draw_fn real,
                               tsdfloc_index_draw_fn index, void* user, uint32_t* parents);
/* Host half of RejectionResampler::resample: slot i proposes itself first, then uniformly drawn particles, until
 * u <= w[j] / sup_w (fp64 quotient of an fp32 weight and the fp64 maximum, :157-180). *n_draws (optional) = index draws used. */
int tsdfloc_rejection_parents(const float* weights, uint64_t stride, uint64_t n, tsdfloc_real_draw_fn real, tsdfloc_index_draw_fn index,
                              void* user, uint64_t max_draws, uint32_t* parents, uint64_t* n_draws);
/* Resampler::resample(ParticleCloud&) for the three drawn methods; particles / n / outputs as tsdfloc_resample (n_out = n). */
int tsdfloc_resample_drawn(tsdfloc_ctx* ctx, int method, const float* particles, uint64_t n, const tsdfloc_draws* draws,
                           float* particles_out, uint64_t cap, uint64_t* n_out, uint32_t* parents);

/* ---- (B) device-pointer stage calls ---------------------------------------------------------------------
 * All pointers are device pointers on ctx's device; `stream` is a cudaStream_t passed as void* (NULL = the
 * ctx's own non-blocking stream; pass cudaStreamLegacy / cudaStreamPerThread explicitly for the default streams).
 * Calls enqueue work and return without synchronising unless stated. */

/* Scan upload + preparation: packs xyz into float4 with the per-point range term
 * (a_range/max_range inside max_range, else a_max; cuda_eval_particles.h:200-209) and reduces their sum. */
int tsdfloc_set_scan_device(tsdfloc_ctx* ctx, const float* d_points_xyz, uint64_t p, void* stream);
int tsdfloc_set_scan_host(tsdfloc_ctx* ctx, const float* points_xyz, uint64_t p, void* stream);

/* Evaluation of particles [first, first+count) of a pose array of n_total particles.
 *   d_particles : n_total x 7 fp32 (Particle layout); only poses are read.
 *   d_raw_weights: n_total fp32; entries [first, first+count) are written with the un-normalised weight
 *                 sum_p (a_hit * tsdf(T_i p) + range_term(p)) (cuda_eval_particles.h:167-215). */
int tsdfloc_eval_device(tsdfloc_ctx* ctx, const float* d_particles, uint64_t n_total, uint64_t first, uint64_t count,
                        const float tf[16], float* d_raw_weights, void* stream);

/* Normalisation + weighted mean + CDF over ALL n_total particles (redundantly on every rank in a multi-GPU run):
 * writes normalised weights into d_particles[:,6], the mean pose (6 fp32) into d_mean_pose, and builds the fp64
 * running-sum CDF kept inside ctx. Sets a device-side flag if the weight sum is 0 (query with tsdfloc_check). */
int tsdfloc_normalize_device(tsdfloc_ctx* ctx, float* d_particles, uint64_t n_total, const float* d_raw_weights,
                             float* d_mean_pose, void* stream);

/* Same CDF build for particles that already carry their weights in slot [6] (no normalisation, nothing rewritten):
 * the device half of tsdfloc_resample_particles. */
int tsdfloc_cdf_device(tsdfloc_ctx* ctx, float* d_particles, uint64_t n_total, float* d_mean_pose, void* stream);

/* Draw output slots [first_out, first_out+count_out) from the CDF: d_particles_out[j] = d_particles[parent(j)].
 * Slots >= n_out (see tsdfloc_check) are filled with parent = last valid parent so buffers stay defined.
 * d_parents optional (uint32 per output slot, indexed from first_out). */
int tsdfloc_draw_device(tsdfloc_ctx* ctx, const float* d_particles, uint64_t n_total, float u0, uint64_t first_out,
                        uint64_t count_out, float* d_particles_out, uint32_t* d_parents, void* stream);

/* Multi-GPU, fused compute + all-gather (no reference counterpart; the reference is single-GPU). Same kernels as
 * tsdfloc_eval_device / tsdfloc_draw_device, but every result is stored not only into this rank's buffer but straight into
 * the same position of every peer's buffer — plain stores through peer-mapped pointers (CUDA IPC / symmetric memory) over
 * NVLink — so no collective follows the kernel; the caller only orders the ranks with a signal barrier.
 *   d_raw_peers / d_out_peers: host array of n_peers (<= 8) device pointers, entry r = rank r's buffer BASE (same layout as
 *   d_raw_weights / the buffer d_particles_out is a slice of: d_out_peers[r] must point at the slot of output `first_out`);
 *   entries that are NULL or equal to the local pointer are skipped. */
int tsdfloc_eval_device_peers(tsdfloc_ctx* ctx, const float* d_particles, uint64_t n_total, uint64_t first, uint64_t count,
                              const float tf[16], float* d_raw_weights, float* const* d_raw_peers, uint32_t n_peers, void* stream);
int tsdfloc_draw_device_peers(tsdfloc_ctx* ctx, const float* d_particles, uint64_t n_total, float u0, uint64_t first_out,
                              uint64_t count_out, float* d_particles_out, float* const* d_out_peers, uint32_t n_peers,
                              uint32_t* d_parents, void* stream);

/* ---- one process, several GPUs ---------------------------------------------------------------------------------
 * The same two host-buffer calls as layer (A), sharded over up to 8 peer-capable devices of one box (no reference
 * counterpart: the reference is single-GPU; its caller, the mcl_3d node, is ONE process — this is how it can use the whole
 * box). Map and scan are replicated, the particles sharded contiguously; the evaluation / draw kernels store their results
 * straight into every device's buffers over NVLink (the *_peers variants above) and CUDA events order the devices — no NCCL,
 * no helper threads. Results are bit-identical to the single-GPU calls. `devices` may name the same device more than once
 * (testing on a one-GPU machine). */
typedef struct tsdfloc_multi tsdfloc_multi;
int tsdfloc_multi_create(const tsdfloc_map_desc* map, const int32_t* grid_occ, const float* data, const tsdfloc_params* params,
                         const int* devices, int n_devices, tsdfloc_multi** out);
void tsdfloc_multi_destroy(tsdfloc_multi* m);
int tsdfloc_multi_device_count(const tsdfloc_multi* m);
const char* tsdfloc_multi_last_error(const tsdfloc_multi* m);
/* The single-device context of rank `rank` (owned by m), e.g. for tsdfloc_resample_particles / tsdfloc_reduce_scan. */
tsdfloc_ctx* tsdfloc_multi_ctx(tsdfloc_multi* m, int rank);
/* tsdfloc_sensor_update_cloud over all devices (TSDFEvaluator::evaluateParticles, tsdf_evaluator.cpp:247-378): the first device
 * reduces the raw cloud, the reduced scan reaches the others over NVLink, then the sharded update. */
int tsdfloc_multi_sensor_update_cloud(tsdfloc_multi* m, float* particles, uint64_t n, const void* xyz_base, uint64_t xyz_stride,
                                      const void* ring_base, uint64_t ring_stride, int ring_bytes, uint64_t n_points, double cell_size,
                                      uint32_t n_rings, uint32_t flags, const float tf[16], float mean_pose[6], uint64_t* n_points_used);
/* tsdfloc_sensor_update / tsdfloc_resample_systematic over all devices (same arguments and status codes). */
int tsdfloc_multi_sensor_update(tsdfloc_multi* m, float* particles, uint64_t n, const float* points, uint64_t p, const float tf[16],
                                float mean_pose[6]);
int tsdfloc_multi_resample_systematic(tsdfloc_multi* m, float u0, float* particles_out, uint64_t cap, uint64_t* n_out);

/* The whole update on device pointers in one call and four kernels (scan preparation + pose matrices, evaluation,
 * normalisation + mean + CDF, draw) on `stream`, without a host synchronisation: d_points_xyz p x 3 fp32; d_particles n x 7
 * (slot 6 receives the normalised weights); output slots [0, count_out) of the systematic resampling with offset u0 go to
 * d_particles_out; d_mean_pose (optional) 6 fp32. Outcome (n_out, zero weight sum) through tsdfloc_check. The single-GPU
 * form of the stage calls below; replaces CudaEvaluator::evaluate + SystematicResampler::resample for a device-resident filter. */
int tsdfloc_update_device(tsdfloc_ctx* ctx, const float* d_points_xyz, uint64_t p, float* d_particles, uint64_t n, const float tf[16],
                          float u0, float* d_particles_out, uint64_t count_out, float* d_mean_pose, void* stream);

/* Synchronises `stream` and reports what the device recorded for the last normalize/draw:
 * n_out = number of particles the reference recurrence emits; returns TSDFLOC_E_NO_VALID_PARTICLE if sum == 0. */
int tsdfloc_check(tsdfloc_ctx* ctx, uint64_t* n_out, double* weight_sum, void* stream);

/* ---- scan reduction (the step in front of the evaluation) -------------------------------------------------------
 * Replaces the serial host reduction inside TSDFEvaluator::evaluateParticles (src/evaluation/tsdf_evaluator.cpp:304-376):
 * points nearer than 1 m are dropped, per (ring, reduction cell of `cell_size`) the FIRST point in cloud order is kept,
 * and the ORIGINAL points are emitted ring by ring, inside a ring in cloud order. Bit-identical to the reference for
 * every input it defines; defined divergences: points with a non-finite coordinate (or an overflowing cell centre) are
 * dropped, and a ring outside [0, n_rings) fails with TSDFLOC_E_BAD_ARG (the reference indexes 64 ring buckets out of
 * bounds, :342,358). n_rings <= 1024, n_points <= 2^22. */
#define TSDFLOC_REDUCE_RING_DESYNC_LIKE_REFERENCE 1u /* reproduce :319-322: a dropped point does not advance the ring
                                                        iterator, so survivor #k is paired with the ring of cloud point #k.
                                                        Default (flag clear): every point keeps its own ring. */
#define TSDFLOC_REDUCE_EMIT_CENTRES 2u /* the cell-CENTRE variant of CudaEvaluator::evaluate(PointCloud2)
                                          (src/cuda/cuda_evaluator.cu:78-116) and src/num_particles_eval.cpp:134-191: no 1 m
                                          test, centres computed in double (`floor(x / cell) * cell + cell/2`), the centres are
                                          emitted instead of the points; ring == NULL deduplicates across rings. */

/* Device pointers: d_points_xyz n x 3 fp32, d_ring n int32 (NULL = one ring), d_points_out capacity n x 3 fp32,
 * d_src_index (optional) capacity n uint32 = cloud position of every emitted point. Enqueues on `stream`. */
int tsdfloc_reduce_scan_device(tsdfloc_ctx* ctx, const float* d_points_xyz, const int32_t* d_ring, uint64_t n_points, double cell_size,
                               uint32_t n_rings, uint32_t flags, float* d_points_out, uint32_t* d_src_index, void* stream);
/* Synchronises `stream`; n_out = number of points the last tsdfloc_reduce_scan_device emitted. */
int tsdfloc_reduce_result(tsdfloc_ctx* ctx, uint64_t* n_out, void* stream);

/* Host buffers, strided so that a sensor_msgs::PointCloud2 byte buffer can be passed as it is: point i has its x, y, z
 * (3 consecutive fp32, like the reference's iter_x[0..2]) at xyz_base + i * xyz_stride and its ring (ring_bytes = 2:
 * int16 as the reference reads it, :305; 4: int32) at ring_base + i * ring_stride; ring_base NULL = one ring.
 * points_out: capacity cap x 3 fp32; src_index optional. */
int tsdfloc_reduce_scan(tsdfloc_ctx* ctx, const void* xyz_base, uint64_t xyz_stride, const void* ring_base, uint64_t ring_stride,
                        int ring_bytes, uint64_t n_points, double cell_size, uint32_t n_rings, uint32_t flags, float* points_out,
                        uint32_t* src_index, uint64_t cap, uint64_t* n_out);

/* One sensor update on a raw cloud — the drop-in for TSDFEvaluator::evaluateParticles(cloud, ..., use_cuda = true)
 * (tsdf_evaluator.cpp:247-378): reduction as above, then exactly tsdfloc_sensor_update on the reduced scan, which never
 * leaves the device. n_points_used (optional) receives the size of the reduced scan. */
int tsdfloc_sensor_update_cloud(tsdfloc_ctx* ctx, float* particles, uint64_t n, const void* xyz_base, uint64_t xyz_stride,
                                const void* ring_base, uint64_t ring_stride, int ring_bytes, uint64_t n_points, double cell_size,
                                uint32_t n_rings, uint32_t flags, const float tf[16], float mean_pose[6], uint64_t* n_points_used);

/* ---- motion update -------------------------------------------------------------------------------------------------
 * ParticleCloud::motionUpdate (src/particle_cloud.cpp:153-617) split in two: a scalar model (host) and the per-particle
 * application (device). Variants and their inputs `in`:
 *   TSDFLOC_MOTION_NOISE      motionUpdate(lin_scale, ang_scale)  :388-420   {lin_scale, ang_scale}
 *   TSDFLOC_MOTION_ODOM       motionUpdate(odom)                  :153-331   {twist.linear.x, twist.angular.z}
 *   TSDFLOC_MOTION_IMU        motionUpdate(imu_data)              :333-386   {linear_vel, angular_yaw}
 *   TSDFLOC_MOTION_NOISE_IMU  motionUpdate(lin_scale, imu_data)   :422-462   {lin_scale, delta_roll, delta_pitch, delta_yaw} */
#define TSDFLOC_MOTION_NOISE 0
#define TSDFLOC_MOTION_ODOM 1
#define TSDFLOC_MOTION_IMU 2
#define TSDFLOC_MOTION_NOISE_IMU 3

/* Mean and standard deviation of the six normal distributions (x y z roll pitch yaw) the variant samples, bit-identical to the
 * reference's (operand types follow it variant by variant). time_diff = seconds since the previous motion update (the
 * reference's FLOAT_T time_diff), a = a_1_..a_12_ (particle_cloud.h:57-68). ref_pose (optional, 6 fp32, in/out): the reference
 * pose whose travelled distance / angle gate the sensor update (particle_cloud.h refDist/refAngle, src/mcl_3d.cpp:353); it is
 * advanced by the ODOM and IMU variants (:180-182, :352-354). Pure host code. */
int tsdfloc_motion_model(int variant, const double in[4], float time_diff, const float a[12], double mean[6], double sigma[6],
                         float ref_pose[6]);

/* apply_model (:496-617) on the device, in place: every particle is moved by its own sample of the six distributions and its
 * Euler angles are read back from the composed rotation; the weight slot is untouched.
 *   d_draws != NULL : n x 6 doubles, the samples themselves (parity mode: with the reference's draws the result is the
 *                     reference's bit for bit); mean/sigma are ignored.
 *   d_draws == NULL : samples = mean + sigma * z with z from a counter-based Philox4x32-10 + Box-Muller stream keyed by
 *                     (seed, sequence, particle index); pass a new `sequence` every update. */
int tsdfloc_motion_update_device(tsdfloc_ctx* ctx, float* d_particles, uint64_t n, const double mean[6], const double sigma[6],
                                 const double* d_draws, uint64_t seed, uint64_t sequence, void* stream);
/* Same on host buffers (particles n x 7 fp32 in place; draws optional n x 6 doubles). */
int tsdfloc_motion_update(tsdfloc_ctx* ctx, float* particles, uint64_t n, const double mean[6], const double sigma[6],
                          const double* draws, uint64_t seed, uint64_t sequence);

/* Particle initialisation on the device — ParticleCloud::initialize x 3 (src/particle_cloud.cpp:32-148). Every particle gets
 * weight 1/n; the six pose components come from the same Philox stream family as the motion update.
 *   TSDFLOC_INIT_NORMAL    ~ N(mean[k], spread[k])                                  initialize(center, n, sigma_x ...)   :32-62
 *   TSDFLOC_INIT_UNIFORM   ~ U(mean[k] - spread[k], mean[k] + spread[k])            initialize(n, center, dx ...)        :64-103
 *   TSDFLOC_INIT_FREE_MAP  xyz = a uniformly drawn free-space voxel with z - 0.5, angles ~ U(mean +- spread)
 *                          initialize(n, free_map, center, droll ...)  :105-148 (global localisation); d_free_map: n_free x 3
 *                          fp32 device array. The reference's index distribution includes size() (:109, out of bounds);
 *                          here it is uniform over [0, n_free).
 * mean = x y z roll pitch yaw of the centre pose (the reference extracts roll/pitch/yaw from the pose's quaternion). */
#define TSDFLOC_INIT_NORMAL 0
#define TSDFLOC_INIT_UNIFORM 1
#define TSDFLOC_INIT_FREE_MAP 2
int tsdfloc_init_particles_device(tsdfloc_ctx* ctx, float* d_particles, uint64_t n, int mode, const double mean[6], const double spread[6],
                                  const float* d_free_map, uint64_t n_free, uint64_t seed, uint64_t sequence, void* stream);
/* Same into a host buffer (particles: n x 7 fp32; free_map: n_free x 3 fp32 host array, uploaded for the call). */
int tsdfloc_init_particles(tsdfloc_ctx* ctx, float* particles, uint64_t n, int mode, const double mean[6], const double spread[6],
                           const float* free_map, uint64_t n_free, uint64_t seed, uint64_t sequence);

/* Arg-max particle of the last normalisation (tsdfloc_sensor_update*, tsdfloc_normalize_device, tsdfloc_cdf_device), the
 * "best pose" mcl_3d picks between evaluation and resampling (src/mcl_3d.cpp:382-399: `if (value > max_value)` starting from
 * 0, i.e. the FIRST particle carrying the largest weight > 0). Reduced inside the normalisation kernels; this call only
 * synchronises `stream` and reads the result. index = -1 (pose zeroed) when no particle has a weight > 0. */
int tsdfloc_best_particle(tsdfloc_ctx* ctx, int64_t* index, float pose[6], float* weight, void* stream);

/* Host-only test hook: the reference's fp32 U recurrence U_{j+1} = (float)((double)U_j + 1/n) evaluated through the
 * same segment-table code the device uses; writes U_j for all j with U_j < limit (at most cap) and returns their count. */
uint64_t tsdfloc_host_u_sequence(float u0, uint64_t n, double limit, float* out, uint64_t cap, uint32_t* n_segs, uint32_t* flags);

/* Measurement probe (not part of the product path): the chip's ceiling for independent 4 B gathers out of the
 * first `bytes` of the uploaded voxel array (0 = all of it; L2-resident up to ~100 MB). spread_sectors = 0: every lane its own
 * random word (32 sectors per warp request, the reference kernel's access shape); 1..32: the 32 lanes of a request fall into
 * that many consecutive 32 B sectors (the evaluation kernel measures 12.5). Average time of `reps` launches after a warm-up. */
int tsdfloc_probe_gather(tsdfloc_ctx* ctx, uint64_t bytes, uint32_t spread_sectors, uint32_t reps, float* ms_per_launch,
                         uint64_t* gathers_per_launch);

/* Cumulative statistics of the evaluation kernel's summation blocks (synchronises the device):
 * out[0] = (particle, block) pairs processed, out[1] = of those folded sequentially (binade crossing, early phase or tie),
 * out[2] = of those caused by an exact rounding tie, out[3] = blocks evaluated a second time with the exact division
 * because a bracketed sub-voxel quotient was open. */
int tsdfloc_eval_stats(tsdfloc_ctx* ctx, uint64_t out[4]);

/* Test / tuning hook (never needed for correct results: every setting produces the same bits). No environment variables
 * are read anywhere in the library.
 *   TSDFLOC_TUNE_SPATIAL_ORDER  -1 automatic (map larger than L2 and >= 16,384 particles), 0 off, 1 on (tsdfloc_sort.cuh)
 *   TSDFLOC_TUNE_EVAL_PAIRING    0 automatic (two points per lane while particle pairs would not fill the warp slots once),
 *                                1 two particles per warp, 2 two points per lane (tsdfloc_eval.cuh)
 *   TSDFLOC_TUNE_DIVISION       -1 what tsdfloc_create proved for the resolution, 0 IEEE division, 1 three-instruction
 *                                quotient, 2 bracketed quotient (an unproven mode is never run)                            */
enum tsdfloc_tune_knob
{
  TSDFLOC_TUNE_SPATIAL_ORDER = 0,
  TSDFLOC_TUNE_EVAL_PAIRING = 1,
  TSDFLOC_TUNE_DIVISION = 2,
  TSDFLOC_TUNE_STAGE_TIMERS = 3,  /* 0 off (default), 1 record CUDA events around the stages (tsdfloc_stage_times) */
  TSDFLOC_TUNE_EVAL_REGISTERS = 4, /* 0 automatic (128 registers per thread below 1.5 waves of 64-register warps, 64 beyond), 1 64, 2 128 */
  TSDFLOC_TUNE_GRAPHS = 5,         /* 1 (default) replay fixed-shape device-resident updates as CUDA graphs, 0 always launch kernel by kernel */
  TSDFLOC_TUNE_EVAL_CHUNKS = 6     /* 0 automatic, 1 every warp walks the whole scan, 2..64 chained scan chunks per particle (tsdfloc_eval.cuh) */
};
int tsdfloc_tune(tsdfloc_ctx* ctx, int knob, int value);

/* Per-stage device times of the most recent update, the counterpart of the reference's RuntimeEvaluator tasks
 * (src/util/runtime_evaluator.cpp:130-149; src/cuda/cuda_evaluator.cu:127,299,362): ms[0] init_kernel (scan preparation,
 * spatial order, pose matrices), ms[1] exec_kernel (k_eval), ms[2] weight_update (normalisation + moments + CDF), ms[3]
 * resampling (draw); 0 for a stage that did not run. Needs TSDFLOC_TUNE_STAGE_TIMERS = 1; synchronises the device. The host
 * side of every stage is also an NVTX range ("tsdfloc:prep_scan", ":eval", ":weight_update", ":resample") for Nsight Systems. */
int tsdfloc_stage_times(tsdfloc_ctx* ctx, float ms[4]);

/* Steady-state CUDA graphs. tsdfloc_update_device records its whole chain of launches the second time it is called with the
 * same buffers, sizes and tuning modes, and replays it with one cudaGraphLaunch from then on (the sensor transform and u0 may
 * differ from call to call: they are patched into the recorded kernel nodes; up to four buffer sets are remembered, so
 * double-buffered outputs are steady state too). This replaces the per-scan re-issue of src/cuda/cuda_evaluator.cu:118-428;
 * results are bit-identical to the kernel-by-kernel launches. A call whose shape was not seen before runs kernel by kernel
 * at no extra cost. The host-buffer calls (tsdfloc_sensor_update) are launched kernel by kernel on purpose: their caller
 * waits for the result and a graph launch measured slower there (profiles/r02_graphs.md). out[0] = recordings made, out[1] =
 * updates served by a graph launch; *note (optional) = why the last recording attempt was abandoned, "" if none was (the
 * update then ran kernel by kernel — still on the GPU; there is no CPU path). */
int tsdfloc_graph_stats(const tsdfloc_ctx* ctx, uint64_t out[2], const char** note);

/* 1 when every fp64 addition of the parallel CDF scan of the last update / resampling call was exact (the result then cannot
 * depend on the order), 0 when one rounded and the CDF was redone in the reference's serial order (k_cdf_exact), -1 without
 * a ctx. Reflects the status block of the last call that read it back. */
int tsdfloc_last_cdf_was_exact(const tsdfloc_ctx* ctx);

/* Quotient mode tsdfloc_create proved for the map's resolution (0 IEEE, 1 three-instruction, 2 bracket); *open_brackets =
 * how many of the 2^30 floats in [0, 1) leave the bracket open (those blocks are evaluated twice). */
int tsdfloc_division_mode(tsdfloc_ctx* ctx, uint64_t* open_brackets);

/* Device time of the most recent evaluation-kernel launch (k_eval alone, CUDA events recorded on the stream it was
 * launched on); waits for that launch to finish. This is the figure bench.py's roofline is computed from. */
int tsdfloc_last_eval_ms(tsdfloc_ctx* ctx, float* ms);

/* Number of kernels this library has launched on this ctx so far (for bench accounting). */
uint64_t tsdfloc_kernel_launches(const tsdfloc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TSDFLOC_H */
