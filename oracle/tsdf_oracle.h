/* oracle/tsdf_oracle.h — CPU restatement of the reference's MCL sensor-update path in plain C.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call this; the product (libtsdfloc.so) never does and
 * fails loudly without its CUDA code.
 *
 * Parity status: PINNED against the unmodified reference compiled here (oracle/_ref/libtsdf_ref.so,
 * built by oracle/Makefile from /root/reference sources): tests/test_oracle_vs_ref.py runs both on the
 * same seeded inputs in this container, and tests/golden/ holds vectors minted from the verbatim build
 * (oracle/gen_golden.py) that travel to the GPU box. The reference itself ships no tests or goldens
 * (SURVEY §4).
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef TSDF_ORACLE_H
#define TSDF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* How a NEGATIVE (or NaN) axis offset x - min is converted to an unsigned integer. The reference code is
 * undefined behaviour there; three behaviours exist in the wild: */
enum oracle_neg_mode
{
  ORACLE_NEG_AS_MISS = 0,      /* product policy: any axis offset not in [0, 2^64) -> miss (init_value)   */
  ORACLE_NEG_REF_HOST_X86 = 1, /* what the reference CPU build does on x86-64 (cvttss2si wrap-around)      */
  ORACLE_NEG_REF_DEVICE_SAT = 2 /* what the reference CUDA build does (cvt.rzi.u32.f32 saturates to 0)    */
};

/* CudaSubVoxelMap<float,float>::MapCoef, include/tsdf_localization/cuda/cuda_sub_voxel_map.h:22-50 */
typedef struct oracle_map_coef
{
  uint64_t dim[3];
  float min[3];
  float max[3];
  float resolution;
  float init_value;
  uint64_t up_dim[3];
  uint64_t up_dim_2;
  uint64_t sub_dim;
  uint64_t sub_dim_2;
  uint64_t grid_occ_size;
  uint64_t data_size;
} oracle_map_coef;

typedef struct oracle_map
{
  oracle_map_coef coef;
  int32_t* grid_occ; /* [grid_occ_size]  -1 or element offset of the brick in data */
  float* data;       /* [data_size] */
} oracle_map;

/* ctor, cuda_sub_voxel_map.tcc:4-48 */
oracle_map* oracle_map_create(const float min[3], const float max[3], float resolution, float init_value);
void oracle_map_destroy(oracle_map* m);
/* setData, cuda_sub_voxel_map.tcc:170-230. cells = n x (x,y,z,value). Returns 0, or 1 on "Upper voxel index overflow!" */
int oracle_map_set_data(oracle_map* m, const float* cells, uint64_t n);
/* Adopt arrays built elsewhere (e.g. by the verbatim reference) without rebuilding. Copies. */
oracle_map* oracle_map_from_arrays(const oracle_map_coef* coef, const int32_t* grid_occ, const float* data);

/* getIndex, cuda_sub_voxel_map.tcc:50-137 (host) / cuda_eval_particles.h:12-67 (device). Returns the flat
 * element offset into data, or coef.data_size for a miss. */
uint64_t oracle_get_index(const oracle_map* m, float x, float y, float z, int neg_mode);
/* getEntry, cuda_sub_voxel_map.tcc:139-157 */
float oracle_get_entry(const oracle_map* m, float x, float y, float z, int neg_mode);
void oracle_get_entries(const oracle_map* m, const float* xyz, uint64_t n, int neg_mode, float* out);

/* createTSDFMap value transform, include/tsdf_localization/map/map_util.h:68-71,124-126 */
float oracle_likelihood_value(float tsdf_mm, float sigma);
float oracle_likelihood_init(float sigma);

/* Pose (x y z roll pitch yaw) o tf_matrix -> 3x4 row-major sensor->map matrix,
 * src/evaluation/tsdf_evaluator.cpp:102-145 */
void oracle_pose_matrix(const float pose6[6], const float tf[16], float out12[12]);

typedef struct oracle_params
{
  float a_hit, a_range, a_max, max_range; /* inv_max_range = 1/max_range, max_range_squared as tsdf_evaluator.h:72-76 */
} oracle_params;

/* evaluatePose, src/evaluation/tsdf_evaluator.cpp:27-76: sequential fp32 sum over points.
 * Optional outputs: idx_out[p] flat index per point (data_size = miss); *hits_out = #points with idx < data_size;
 * *w64_out = the same sum accumulated in double (for the tolerance analysis of SURVEY §7.3(3)). */
float oracle_pose_weight(const oracle_map* m, const oracle_params* prm, const float mat12[12], const float* points_xyz,
                         uint64_t n_points, int neg_mode, uint32_t* idx_out, uint32_t* hits_out, double* w64_out);

/* evaluate (CPU branch), src/evaluation/tsdf_evaluator.cpp:85-221.
 * particles: N x 7 (x y z r p y w). On return w holds the NORMALISED weight; raw_out (optional, N) the
 * un-normalised one; mean_pose6 = weighted mean xyz + atan2 of weighted sin/cos sums; idx_out optional N*P
 * uint32; hits_out optional N. weight_sum is accumulated in particle order in fp32 (one legal ordering of
 * the reference's OpenMP reduction). Returns 0, or 1 when weight_sum == 0 ("No particle is valid!"). */
int oracle_evaluate(const oracle_map* m, const oracle_params* prm, float* particles7, uint64_t n, const float* points_xyz,
                    uint64_t n_points, const float tf[16], int neg_mode, float* raw_out, float mean_pose6[6],
                    uint32_t* idx_out, uint32_t* hits_out, float* weight_sum_out);

/* SystematicResampler::resample, include/tsdf_localization/resampling/novel_resampling.h:41-72, with the
 * random offset U0 injected. weights = N normalised fp32 weights. parents_out[j] = index of the particle
 * copied into output slot j. Returns the number of output particles the reference loop produces (can differ
 * from N); at most cap parents are written. */
uint64_t oracle_systematic_resample(const float* weights, uint64_t n, float u0, uint32_t* parents_out, uint64_t cap);

/* ResidualSystematicResampler::resample, novel_resampling.h:76-104, with the uniform(0,1) draw u0 handed in. Returns the
 * output length; parents_out receives min(length, cap) source indices. */
uint64_t oracle_residual_systematic_resample(const float* weights, uint64_t n, float u0, uint32_t* parents_out, uint64_t cap);

/* ResidualResampler::resample, novel_resampling.h:9-36, with the uniform index draws handed in. parents_out: n entries.
 * Returns the output length (n unless the draws ran out); *draws_used = draws consumed. */
uint64_t oracle_residual_resample(const float* weights, uint64_t n, const uint64_t* draws, uint64_t n_draws, uint32_t* parents_out,
                                  uint64_t* draws_used);

/* The remaining resamplers of mcl_3d's switch (src/mcl_3d.cpp:243-263), every random draw handed in through callbacks:
 * real() = one uniform_real draw as FLOAT_T, index_draw() = one uniform_int_distribution<size_t>(0, n - 1) draw.
 * parents_out: n entries (all three emit exactly n particles).
 *   WheelResampler::resample       src/resampling/wheel_resampler.cpp:6-34
 *   MetropolisResampler::resample  include/tsdf_localization/resampling/novel_resampling.h:106-144
 *   RejectionResampler::resample   include/tsdf_localization/resampling/novel_resampling.h:146-189 */
typedef float (*oracle_real_draw_fn)(void* user);
typedef uint64_t (*oracle_index_draw_fn)(void* user);
void oracle_wheel_resample(const float* weights, uint64_t n, oracle_real_draw_fn real, void* user, uint32_t* parents_out);
void oracle_metropolis_resample(const float* weights, uint64_t n, uint64_t steps, oracle_real_draw_fn real, oracle_index_draw_fn index_draw,
                                void* user, uint32_t* parents_out);
void oracle_rejection_resample(const float* weights, uint64_t n, oracle_real_draw_fn real, oracle_index_draw_fn index_draw, void* user,
                               uint32_t* parents_out);
/* Deterministic draw source for tests (splitmix64); pass the handle as `user`. */
void* oracle_draws_create(uint64_t seed, uint64_t n);
void oracle_draws_destroy(void* draws);
float oracle_draw_real(void* draws);
uint64_t oracle_draw_index(void* draws);
void oracle_draws_used(void* draws, uint64_t* n_real, uint64_t* n_index);


/* Scan reduction of TSDFEvaluator::evaluateParticles, src/evaluation/tsdf_evaluator.cpp:304-376: drop points nearer than
 * 1 m (:317-322), keep per (ring, reduction cell) the FIRST point in cloud order (unordered_set insert of SortClass keyed
 * on ring + cell centre, :324-329; equality cuda_evaluator.h:56-59), emit the ORIGINAL points ordered by ring, then by the
 * running index of the kept points (:340-376).
 *   ring_desync != 0 reproduces the reference's iterator bug: iter_ring is not advanced for a dropped point (:319-322 `continue`
 *   skips `++iter_ring`), so the k-th point that survives the 1 m test is paired with the ring field of the k-th cloud point.
 *   ring_desync == 0 is the product's default: every point keeps its own ring.
 * Defined divergences (reference is undefined there): points with a non-finite coordinate are dropped (the reference's
 * hash casts NaN*1000 to long long); rings outside [0, n_rings) make the call fail with -1 (the reference indexes a
 * 64-entry vector out of bounds, :342,358).
 * src_index_out[j] = cloud position of output point j. Returns the number of output points, or -1. */
int64_t oracle_reduce_scan(const float* points_xyz, const int32_t* ring, uint64_t n, float cell_size, uint32_t n_rings,
                           int ring_desync, float* points_out, uint32_t* src_index_out);

/* The ring-agnostic cell-CENTRE reduction the reference's PointCloud2 overload and benchmark driver use
 * (src/cuda/cuda_evaluator.cu:78-116, src/num_particles_eval.cpp:134-191): every point is replaced by the centre of its
 * `cell` (double arithmetic, `floor(x / 0.064) * 0.064 + 0.032` rounded to float), duplicates (per ring when ring != NULL,
 * else globally) are dropped keeping the first, order = ring, then cloud position (the ring-agnostic reference variant
 * emits unordered_set iteration order, which is implementation-defined; as a SET the output is identical). */
int64_t oracle_reduce_scan_centres(const float* points_xyz, const int32_t* ring, uint64_t n, double cell_size, uint32_t n_rings,
                                   float* points_out, uint32_t* src_index_out);

/* ---- motion update ------------------------------------------------------------------------------
 * The four ParticleCloud::motionUpdate variants (src/particle_cloud.cpp:153-462) differ only in the mean / standard
 * deviation of six normal distributions (x y z roll pitch yaw); the per-particle work is apply_model (:496-617).
 *   variant 0  motionUpdate(lin_scale, ang_scale)  :388-420   in = {lin_scale, ang_scale}            (fp32 arithmetic)
 *   variant 1  motionUpdate(odom)                  :153-331   in = {linear.x, angular.z}             (fp64 arithmetic)
 *   variant 2  motionUpdate(imu_data)              :333-386   in = {linear_vel, angular_yaw}         (fp32)
 *   variant 3  motionUpdate(lin_scale, imu_data)   :422-462   in = {lin_scale, d_roll, d_pitch, d_yaw} (fp32)
 * time_diff is the reference's FLOAT_T time_diff; a = a_1_..a_12_ (particle_cloud.h:57-68). Also advances the reference
 * pose (ref_pose, :180-182 / :352-354) when ref_pose != NULL (variants 1 and 2 only). Returns 0, or 1 for a bad variant. */
int oracle_motion_model(int variant, const double in[4], float time_diff, const float a[12], double mean[6], double sigma[6],
                        float ref_pose[6]);
/* apply_model: particles n x 7 in place (weights untouched), draws n x 6 doubles (the values the six distributions returned). */
void oracle_motion_apply(float* particles7, uint64_t n, const double* draws6);

#ifdef __cplusplus
}
#endif
#endif
