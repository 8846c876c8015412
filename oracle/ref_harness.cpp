// oracle/ref_harness.cpp — C ABI over the UNMODIFIED reference classes (test infrastructure only).
//
// This translation unit is compiled together with the reference's own sources, taken by path from
// /root/reference (never copied into this repo):
//   src/evaluation/tsdf_evaluator.cpp            (TSDFEvaluator::evaluate / evaluatePose, CPU/OpenMP)
//   src/evaluation/model/likelihood_evaluation.cpp
//   include/tsdf_localization/cuda/cuda_sub_voxel_map.{h,tcc}   (two-level sparse map)
//   include/tsdf_localization/resampling/novel_resampling.h     (SystematicResampler)
// against the stub ROS headers in oracle/ref_stubs/. The result (oracle/_ref/libtsdf_ref*.so) is
// used ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
// as the checker and the timed CPU baseline. Nothing in the product path links or loads it.
//
// What is verbatim and what is glue:
//   * map build (setData), host getEntry, evaluate(), evaluatePose(), SystematicResampler::resample()
//     run the reference's own code, bit for bit.
//   * createTSDFMap (map/map_util.h:17-154) runs verbatim on an in-memory stand-in for the HDF5 file
//     (ref_stubs/highfive/H5File.hpp; libhdf5 is not in this image).
//   * ParticleCloud's three trivial accessors (default ctor, operator[], size) are defined here because
//     src/particle_cloud.cpp drags in tf2/ROS-time code unrelated to this path
//     (reference: src/particle_cloud.cpp:11-15, 619-632).
//   * CudaEvaluator's four symbols are defined as "no CUDA" stubs unless TSDF_REF_WITH_B200_SHIM is set,
//     in which case the B200 drop-in shim provides them (tsdf_evaluator.h:78 constructs it unconditionally), or
//     TSDF_REF_WITH_REF_CUDA, in which case the reference's OWN CUDA evaluator (src/cuda/*.cu, compiled unmodified by nvcc)
//     provides them: libtsdf_ref_cuda.so, the reference GPU arm of bench.py.
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include <tsdf_localization/cuda/cuda_sub_voxel_map.h>
#include <tsdf_localization/evaluation/tsdf_evaluator.h>
#include <tsdf_localization/evaluation/model/likelihood_evaluation.h>
#include <tsdf_localization/map/map_util.h>
#include <tsdf_localization/resampling/novel_resampling.h>
#include <tsdf_localization/resampling/wheel_resampler.h>
#include <tsdf_localization/util/constant.h>
#include <tsdf_localization/util/mcl_file.h>
#ifdef TSDF_REF_WITH_B200_SHIM
#include "tsdfloc_shim.h"
#endif

namespace tsdf_localization
{

// --- glue: trivial ParticleCloud members (see header comment) ---------------------------------
ParticleCloud::ParticleCloud()
{
  m_generator_ptr.reset(new std::mt19937(0u));
}
Particle& ParticleCloud::operator[](unsigned int index)
{
  if (index >= size())
  {
    throw std::out_of_range("Index exceeds number of particles");
  }
  return m_particles[index];
}
std::size_t ParticleCloud::size() const
{
  return m_particles.size();
}

#if !defined(TSDF_REF_WITH_B200_SHIM) && !defined(TSDF_REF_WITH_REF_CUDA)
// --- glue: CPU-only build has no CUDA back-end; mirror the base class behaviour ----------------
CudaEvaluator::CudaEvaluator(CudaSubVoxelMap<FLOAT_T, FLOAT_T>&, bool per_point, FLOAT_T a_hit, FLOAT_T a_range, FLOAT_T a_max, FLOAT_T max_range)
: d_map_(nullptr), per_point_(per_point), d_grid_occ_(nullptr), d_data_(nullptr), d_particles_(nullptr), d_particles_ordered_(nullptr),
  particles_reserved_(0), d_points_(nullptr), d_points_ordered_(nullptr), points_reserved_(0), d_transform_(nullptr), d_new_weights_(nullptr),
  d_point_weights_(nullptr), point_weights_size_(0), p_x_(nullptr), p_y_(nullptr), p_z_(nullptr), sin_a_(nullptr), cos_a_(nullptr),
  sin_b_(nullptr), cos_b_(nullptr), sin_c_(nullptr), cos_c_(nullptr), a_hit_(a_hit), a_range_(a_range), a_max_(a_max), max_range_(max_range),
  inv_max_range_(1.0 / max_range), max_range_squared_(max_range * max_range)
{
}
CudaEvaluator::~CudaEvaluator() {}
geometry_msgs::PoseWithCovariance CudaEvaluator::evaluate(std::vector<Particle>&, const sensor_msgs::PointCloud2&, FLOAT_T*)
{
  throw std::runtime_error("CUDA acceleration is not supported. Please install CUDA!");
}
geometry_msgs::PoseWithCovariance CudaEvaluator::evaluate(std::vector<Particle>&, const std::vector<CudaPoint>&, FLOAT_T*)
{
  throw std::runtime_error("CUDA acceleration is not supported. Please install CUDA!");
}
#endif

namespace
{

using RefMap = CudaSubVoxelMap<FLOAT_T, FLOAT_T>;

// With the B200 shim the facade is the shim's TSDFEvaluatorB200 (evaluateParticles override: GPU scan reduction).
#ifdef TSDF_REF_WITH_B200_SHIM
using EvaluatorBase = TSDFEvaluatorB200;
#else
using EvaluatorBase = TSDFEvaluator;
#endif

// evaluatePose is protected virtual (tsdf_evaluator.h:59) — expose it unchanged.
struct ExposedEvaluator : public EvaluatorBase
{
  using EvaluatorBase::EvaluatorBase;
  FLOAT_T pose_weight(FLOAT_T* pose, const std::vector<CudaPoint>& cloud)
  {
    LikelihoodEvaluation eval(10000);
    return evaluatePose(pose, cloud, eval);
  }
};

// evaluate() is virtual (tsdf_evaluator.h:114): capture the reduced, ordered scan evaluateParticles hands to it
// (tsdf_evaluator.cpp:378) instead of evaluating it.
struct CaptureEvaluator : public TSDFEvaluator
{
  using TSDFEvaluator::TSDFEvaluator;
  std::vector<CudaPoint> captured;
  geometry_msgs::PoseWithCovariance evaluate(std::vector<Particle>&, const std::vector<CudaPoint>& points, FLOAT_T*, bool) override
  {
    captured = points;
    return geometry_msgs::PoseWithCovariance();
  }
};

// m_generator_ptr is protected (resampler.h:28) — reseed it so U0 is reproducible.
struct SeededSystematic : public SystematicResampler
{
  void seed(uint32_t s) { m_generator_ptr.reset(new std::mt19937(s)); }
};
struct SeededResidual : public ResidualResampler
{
  void seed(uint32_t s) { m_generator_ptr.reset(new std::mt19937(s)); }
};
struct SeededResidualSystematic : public ResidualSystematicResampler
{
  void seed(uint32_t s) { m_generator_ptr.reset(new std::mt19937(s)); }
};
// sampling_steps_ of the Metropolis resamplers constructed below (mcl_3d passes 50, src/mcl_3d.cpp:258)
size_t g_metropolis_steps = 50;
struct SeededWheel : public WheelResampler
{
  void seed(uint32_t s) { m_generator_ptr.reset(new std::mt19937(s)); }
};
struct SeededMetropolis : public MetropolisResampler
{
  SeededMetropolis() : MetropolisResampler(g_metropolis_steps) {}
  void seed(uint32_t s) { m_generator_ptr.reset(new std::mt19937(s)); }
};
struct SeededRejection : public RejectionResampler
{
  void seed(uint32_t s) { m_generator_ptr.reset(new std::mt19937(s)); }
};
#ifdef TSDF_REF_WITH_B200_SHIM
struct SeededGpuMetropolis : public GpuMetropolisResampler
{
  SeededGpuMetropolis() : GpuMetropolisResampler(g_metropolis_steps) {}
};
#endif
// A seeded std::mt19937 with the distribution objects the reference's resamplers construct, behind C callbacks: feeds
// restatements the very draws — in the very interleaving — the reference consumes.
struct DrawSource
{
  std::mt19937 gen;
  std::uniform_real_distribution<FLOAT_T> real;    // novel_resampling.h:115, 151
  std::uniform_real_distribution<> real_wheel;     // wheel_resampler.cpp:9
  std::uniform_int_distribution<size_t> index;     // novel_resampling.h:116, 152
  uint64_t n_real = 0, n_index = 0;
};

struct MapHandle
{
  std::shared_ptr<RefMap> map;
};

struct EvalHandle
{
  std::shared_ptr<RefMap> map;
  std::unique_ptr<ExposedEvaluator> eval;
  std::string last_error;
};

thread_local std::string g_last_error;

}  // namespace
}  // namespace tsdf_localization

using namespace tsdf_localization;

// ResidualSystematicResampler::resample (novel_resampling.h:79-103) / ResidualResampler::resample (:12-34), verbatim, with a
// seeded generator. method: 1 = Residual, 2 = ResidualSystematic, 3 = Wheel, 4 = Metropolis, 5 = Rejection (the numbering of
// include/tsdfloc.h). u_out: the uniform(0,1) draw of ResidualSystematic.
template <typename R>
static uint64_t run_seeded_resampler(const float* particles, uint64_t n, uint32_t seed, float* particles_out, uint64_t cap)
{
  ParticleCloud cloud;
  cloud.particles().resize(n);
  std::memcpy(static_cast<void*>(cloud.particles().data()), particles, n * sizeof(Particle));
  R rs;
  rs.seed(seed);
  Resampler& base = rs;
  base.resample(cloud);
  const uint64_t m = cloud.size();
  const uint64_t c = m < cap ? m : cap;
  if (particles_out && c) std::memcpy(particles_out, static_cast<void*>(cloud.particles().data()), c * sizeof(Particle));
  return m;
}

extern "C"
{

// POD mirror of CudaSubVoxelMap::MapCoef (cuda_sub_voxel_map.h:22-50) with fixed-width fields.
struct ref_map_coef
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
};

const char* ref_last_error() { return g_last_error.c_str(); }
unsigned ref_omp_threads() { return OMP_THREADS; }

void* ref_map_create(const float mn[3], const float mx[3], float resolution, float init_value)
{
  try
  {
    auto* h = new MapHandle;
    h->map = std::make_shared<RefMap>(RefMap(mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], resolution, init_value));
    return h;
  }
  catch (std::exception& e)
  {
    g_last_error = e.what();
    return nullptr;
  }
}

// NOTE: std::make_shared<RefMap>(RefMap(...)) move-constructs from a temporary exactly as
// map_util.h:73 does (the class is move-only, cuda_sub_voxel_map.h:121-137).

void ref_map_destroy(void* h) { delete static_cast<MapHandle*>(h); }

// createTSDFMap<CudaSubVoxelMap<float,float>, float, float>(file, free_map, sigma) (map_util.h:17-154) on n_chunks chunks of
// 64^3 raw TSDFValue words, chunk_pos = n_chunks x (cx, cy, cz). Returns a map handle; free_out (optional, capacity
// free_cap points x 3 floats) receives the free-space voxels, *n_free their number.
void* ref_create_tsdf_map(const int32_t* chunk_pos, const uint32_t* chunk_data, uint64_t n_chunks, float sigma, float* free_out,
                          uint64_t free_cap, uint64_t* n_free)
{
  try
  {
    const size_t words = static_cast<size_t>(CHUNK_SIZE) * CHUNK_SIZE * CHUNK_SIZE;
    auto& group = HighFive::stub_files()["mem.h5"]["/map"];
    group.clear();
    for (uint64_t c = 0; c < n_chunks; ++c)
    {
      const std::string tag = std::to_string(chunk_pos[3 * c]) + "_" + std::to_string(chunk_pos[3 * c + 1]) + "_" + std::to_string(chunk_pos[3 * c + 2]);
      group[tag].assign(chunk_data + c * words, chunk_data + (c + 1) * words);
    }
    std::vector<CudaPoint> free_map;
    auto* h = new MapHandle;
    h->map = createTSDFMap<RefMap, FLOAT_T, FLOAT_T>("mem.h5", free_map, sigma);
    group.clear();
    if (n_free) *n_free = free_map.size();
    if (free_out)
    {
      const uint64_t c = free_map.size() < free_cap ? free_map.size() : free_cap;
      if (c) std::memcpy(free_out, static_cast<void*>(free_map.data()), c * sizeof(CudaPoint));
    }
    return h;
  }
  catch (std::exception& ex)
  {
    g_last_error = ex.what();
    return nullptr;
  }
}

// cells: n × (x, y, z, value) fp32 — the tuple list createTSDFMap hands to setData (map_util.h:129,152).
int ref_map_set_data(void* h, const float* cells, uint64_t n)
{
  try
  {
    std::vector<std::tuple<FLOAT_T, FLOAT_T, FLOAT_T, FLOAT_T>> data;
    data.reserve(n);
    for (uint64_t i = 0; i < n; ++i)
    {
      data.push_back(std::make_tuple(cells[4 * i], cells[4 * i + 1], cells[4 * i + 2], cells[4 * i + 3]));
    }
    static_cast<MapHandle*>(h)->map->setData(data);
    return 0;
  }
  catch (std::exception& e)
  {
    g_last_error = e.what();
    return 1;
  }
}

// Large synthetic maps (BASELINE configs C4 / C5: 1.4 GB of bricks) are generated as arrays, not as 3.5e8 tuples: hand the
// reference's map object the two arrays setData would have produced (layout of cuda_sub_voxel_map.tcc:196-224), through its
// own public accessors (rawGridOcc(), rawDataPtr(), coef(); cuda_sub_voxel_map.h:92-125, 271). Lookups then run verbatim.
int ref_map_adopt_arrays(void* h, const int32_t* grid_occ, const float* data, uint64_t data_size)
{
  try
  {
    RefMap& m = *static_cast<MapHandle*>(h)->map;
    std::memcpy(m.rawGridOcc(), grid_occ, m.gridOccBytes());
    float** slot = m.rawDataPtr();
    delete[] *slot;
    *slot = new float[data_size ? data_size : 1];
    std::memcpy(*slot, data, sizeof(float) * data_size);
    m.coef().data_size_ = data_size;
    return 0;
  }
  catch (std::exception& e)
  {
    g_last_error = e.what();
    return 1;
  }
}

void ref_map_get_coef(void* h, ref_map_coef* out)
{
  const auto& c = static_cast<MapHandle*>(h)->map->coef();
  out->dim[0] = c.dim_x_; out->dim[1] = c.dim_y_; out->dim[2] = c.dim_z_;
  out->min[0] = c.min_x_; out->min[1] = c.min_y_; out->min[2] = c.min_z_;
  out->max[0] = c.max_x_; out->max[1] = c.max_y_; out->max[2] = c.max_z_;
  out->resolution = c.resolution_;
  out->init_value = c.init_value_;
  out->up_dim[0] = c.up_dim_x_; out->up_dim[1] = c.up_dim_y_; out->up_dim[2] = c.up_dim_z_;
  out->up_dim_2 = c.up_dim_2_;
  out->sub_dim = c.sub_dim_;
  out->sub_dim_2 = c.sub_dim_2_;
  out->grid_occ_size = c.grid_occ_size_;
  out->data_size = c.data_size_;
}

const int* ref_map_grid_occ(void* h) { return static_cast<MapHandle*>(h)->map->rawGridOcc(); }
const float* ref_map_data(void* h) { return static_cast<MapHandle*>(h)->map->rawData(); }

// Host getEntry (cuda_sub_voxel_map.tcc:139-157) on n query points.
void ref_map_get_entries(void* h, const float* xyz, uint64_t n, float* out)
{
  auto& map = *static_cast<MapHandle*>(h)->map;
  for (uint64_t i = 0; i < n; ++i)
  {
    out[i] = map.getEntry(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  }
}

void* ref_eval_create(void* map_h, float a_hit, float a_range, float a_max, float max_range)
{
  try
  {
    auto* e = new EvalHandle;
    e->map = static_cast<MapHandle*>(map_h)->map;
    e->eval.reset(new ExposedEvaluator(e->map, false, a_hit, a_range, a_max, max_range));
    return e;
  }
  catch (std::exception& ex)
  {
    g_last_error = ex.what();
    return nullptr;
  }
}

// Same with TSDFEvaluator's 7th constructor argument (reduction_cell_size, tsdf_evaluator.h:72).
void* ref_eval_create_cell(void* map_h, float a_hit, float a_range, float a_max, float max_range, float reduction_cell_size)
{
  try
  {
    auto* e = new EvalHandle;
    e->map = static_cast<MapHandle*>(map_h)->map;
    e->eval.reset(new ExposedEvaluator(e->map, false, a_hit, a_range, a_max, max_range, reduction_cell_size));
    return e;
  }
  catch (std::exception& ex)
  {
    g_last_error = ex.what();
    return nullptr;
  }
}

void ref_eval_destroy(void* e) { delete static_cast<EvalHandle*>(e); }

namespace
{
// Packed test cloud: x y z float32 at byte 0/4/8, ring int16 at byte 12, point_step 16.
sensor_msgs::PointCloud2 make_cloud(const float* xyz, const int16_t* ring, uint64_t n)
{
  sensor_msgs::PointCloud2 cloud;
  cloud.width = static_cast<uint32_t>(n);
  cloud.height = 1;
  cloud.point_step = 16;
  cloud.row_step = cloud.point_step * cloud.width;
  cloud.fields.resize(4);
  const char* names[4] = {"x", "y", "z", "ring"};
  for (int f = 0; f < 4; ++f)
  {
    cloud.fields[f].name = names[f];
    cloud.fields[f].offset = 4u * f;
  }
  cloud.data.assign(static_cast<size_t>(n) * 16, 0);
  for (uint64_t i = 0; i < n; ++i)
  {
    std::memcpy(&cloud.data[16 * i], xyz + 3 * i, 12);
    std::memcpy(&cloud.data[16 * i + 12], ring + i, 2);
  }
  return cloud;
}
}  // namespace

// TSDFEvaluator::evaluateParticles(particle_cloud, cloud, "", "", use_cuda, ignore_tf = true) (tsdf_evaluator.cpp:247-378) on a
// packed cloud. CPU build: the reference's own reduction + CPU evaluation. Shim build: the facade is TSDFEvaluatorB200, so
// use_cuda != 0 runs reduction + evaluation on the B200 (desync: reproduce the reference's ring-iterator bug).
// reduced_out (shim build, use_cuda only): size of the reduced scan.
int ref_evaluate_cloud(void* eh, float* particles, uint64_t n, const float* xyz, const int16_t* ring, uint64_t n_points, int use_cuda,
                       int desync, uint32_t n_rings, double pose_out[7], uint64_t* reduced_out)
{
  auto* e = static_cast<EvalHandle*>(eh);
  try
  {
    ParticleCloud pc;
    pc.particles().resize(n);
    std::memcpy(static_cast<void*>(pc.particles().data()), particles, n * sizeof(Particle));
    const sensor_msgs::PointCloud2 cloud = make_cloud(xyz, ring, n_points);
#ifdef TSDF_REF_WITH_B200_SHIM
    e->eval->ring_desync_like_reference = desync != 0;
    e->eval->n_rings = n_rings;
#else
    (void)desync;
    (void)n_rings;
#endif
    auto pose = e->eval->evaluateParticles(pc, cloud, "", "", use_cuda != 0, true);
    std::memcpy(particles, static_cast<void*>(pc.particles().data()), n * sizeof(Particle));
    if (pose_out)
    {
      pose_out[0] = pose.pose.position.x; pose_out[1] = pose.pose.position.y; pose_out[2] = pose.pose.position.z;
      pose_out[3] = pose.pose.orientation.x; pose_out[4] = pose.pose.orientation.y;
      pose_out[5] = pose.pose.orientation.z; pose_out[6] = pose.pose.orientation.w;
    }
    if (reduced_out)
    {
#ifdef TSDF_REF_WITH_B200_SHIM
      *reduced_out = use_cuda ? e->eval->last_reduced_size() : 0;
#else
      *reduced_out = 0;
#endif
    }
    return 0;
  }
  catch (std::exception& ex)
  {
    g_last_error = ex.what();
    return 1;
  }
}

// TSDFEvaluator::evaluate(particles, points, tf, use_cuda) (tsdf_evaluator.cpp:78-245).
// particles: N × 7 fp32 (x y z roll pitch yaw weight), weights overwritten with the NORMALISED weights.
// pose_out: x y z qx qy qz qw (doubles).  Returns 0, or 1 on an exception (e.g. "No particle is valid!").
int ref_evaluate(void* eh, float* particles, uint64_t n, const float* points, uint64_t p, const float tf[16], int use_cuda, double pose_out[7])
{
  auto* e = static_cast<EvalHandle*>(eh);
  static_assert(sizeof(Particle) == 7 * sizeof(float), "Particle must be 7 packed floats");
  static_assert(sizeof(CudaPoint) == 3 * sizeof(float), "CudaPoint must be 3 packed floats");
  try
  {
    std::vector<Particle> ps(n);
    std::memcpy(static_cast<void*>(ps.data()), particles, n * sizeof(Particle));
    std::vector<CudaPoint> pts(p);
    std::memcpy(static_cast<void*>(pts.data()), points, p * sizeof(CudaPoint));
    FLOAT_T tfm[16];
    std::memcpy(tfm, tf, sizeof(tfm));
    auto pose = e->eval->evaluate(ps, pts, tfm, use_cuda != 0);
    std::memcpy(particles, static_cast<void*>(ps.data()), n * sizeof(Particle));
    if (pose_out)
    {
      pose_out[0] = pose.pose.position.x; pose_out[1] = pose.pose.position.y; pose_out[2] = pose.pose.position.z;
      pose_out[3] = pose.pose.orientation.x; pose_out[4] = pose.pose.orientation.y;
      pose_out[5] = pose.pose.orientation.z; pose_out[6] = pose.pose.orientation.w;
    }
    return 0;
  }
  catch (std::exception& ex)
  {
    g_last_error = ex.what();
    return 1;
  }
}

// evaluatePose (tsdf_evaluator.cpp:27-76) for n pre-built 3x4 (row-major, 12 floats) sensor→map matrices:
// the un-normalised per-particle weight.
void ref_pose_weights(void* eh, const float* matrices12, uint64_t n, const float* points, uint64_t p, float* weights_out)
{
  auto* e = static_cast<EvalHandle*>(eh);
  std::vector<CudaPoint> pts(p);
  std::memcpy(static_cast<void*>(pts.data()), points, p * sizeof(CudaPoint));
  for (uint64_t i = 0; i < n; ++i)
  {
    FLOAT_T pose[16] = {0};
    std::memcpy(pose, matrices12 + 12 * i, 12 * sizeof(float));
    pose[15] = 1;
    weights_out[i] = e->eval->pose_weight(pose, pts);
  }
}

// SystematicResampler::resample (novel_resampling.h:41-72) with a seeded generator.
// u0_out receives the U the resampler drew (same generator state, same distribution type).
// Returns the number of particles the reference produced (may differ from n); at most cap are copied out.
uint64_t ref_systematic_resample(const float* particles, uint64_t n, uint32_t seed, float* particles_out, uint64_t cap, float* u0_out)
{
  ParticleCloud cloud;
  cloud.particles().resize(n);
  std::memcpy(static_cast<void*>(cloud.particles().data()), particles, n * sizeof(Particle));
  if (u0_out)
  {
    std::mt19937 gen(seed);
    auto inverse_M = 1.0 / n;
    std::uniform_real_distribution<FLOAT_T> uniform_distribution(0.0, inverse_M);
    *u0_out = uniform_distribution(gen);
  }
  SeededSystematic rs;
  rs.seed(seed);
  rs.resample(cloud);
  const uint64_t m = cloud.size();
  const uint64_t c = m < cap ? m : cap;
  if (particles_out && c)
  {
    std::memcpy(particles_out, static_cast<void*>(cloud.particles().data()), c * sizeof(Particle));
  }
  return m;
}

uint64_t ref_resample_method(int method, const float* particles, uint64_t n, uint32_t seed, float* particles_out, uint64_t cap, float* u_out)
{
  if (u_out)
  {
    std::mt19937 gen(seed);
    std::uniform_real_distribution<FLOAT_T> uniform_distribution(0.0, 1.0);
    *u_out = uniform_distribution(gen);
  }
  if (method == 1) return run_seeded_resampler<SeededResidual>(particles, n, seed, particles_out, cap);
  if (method == 2) return run_seeded_resampler<SeededResidualSystematic>(particles, n, seed, particles_out, cap);
  if (method == 3) return run_seeded_resampler<SeededWheel>(particles, n, seed, particles_out, cap);        // wheel_resampler.cpp:6-34
  if (method == 4) return run_seeded_resampler<SeededMetropolis>(particles, n, seed, particles_out, cap);   // novel_resampling.h:106-144
  if (method == 5) return run_seeded_resampler<SeededRejection>(particles, n, seed, particles_out, cap);    // novel_resampling.h:146-189
  return ~0ull;
}
void ref_set_metropolis_steps(uint64_t steps) { g_metropolis_steps = steps; }

// Draw callbacks on a seeded generator (see DrawSource). `n` is the particle count of the index distribution.
void* ref_draws_create(uint32_t seed, uint64_t n)
{
  return new DrawSource{std::mt19937(seed), std::uniform_real_distribution<FLOAT_T>(0.0, 1.0), std::uniform_real_distribution<>(0.0, 1.0),
                        std::uniform_int_distribution<size_t>(0, n - 1)};
}
void ref_draws_destroy(void* d) { delete static_cast<DrawSource*>(d); }
float ref_draw_real(void* p)
{
  DrawSource* d = static_cast<DrawSource*>(p);
  ++d->n_real;
  return d->real(d->gen);
}
float ref_draw_real_wheel(void* p)
{
  DrawSource* d = static_cast<DrawSource*>(p);
  ++d->n_real;
  const FLOAT_T random_value = d->real_wheel(d->gen);
  return random_value;
}
uint64_t ref_draw_index(void* p)
{
  DrawSource* d = static_cast<DrawSource*>(p);
  ++d->n_index;
  return d->index(d->gen);
}
void ref_draws_used(void* p, uint64_t* n_real, uint64_t* n_index)
{
  DrawSource* d = static_cast<DrawSource*>(p);
  *n_real = d->n_real;
  *n_index = d->n_index;
}

// The first `count` draws of std::uniform_int_distribution<size_t>(0, n - 1) on std::mt19937(seed): the index stream
// ResidualResampler::resample consumes (novel_resampling.h:14,21), for feeding restatements the same draws.
void ref_uniform_index_draws(uint32_t seed, uint64_t n, uint64_t count, uint64_t* out)
{
  std::mt19937 gen(seed);
  std::uniform_int_distribution<size_t> uniform_distribution(0, n - 1);
  for (uint64_t i = 0; i < count; ++i) out[i] = uniform_distribution(gen);
}

// MCLFile::write / MCLFile::read (src/util/mcl_file.cpp:14-113), verbatim. pose7 = x y z q1 q2 q3 q4.
int ref_mcl_write(const char* name, const float* points, const int32_t* rings, uint64_t p, const float* particles, uint64_t n,
                  const float tf[16], const float pose7[7])
{
  try
  {
    std::vector<CudaPoint> pts(p);
    if (p) std::memcpy(static_cast<void*>(pts.data()), points, p * sizeof(CudaPoint));
    std::vector<int> rg(rings, rings + p);
    std::vector<Particle> ps(n);
    if (n) std::memcpy(static_cast<void*>(ps.data()), particles, n * sizeof(Particle));
    std::array<FLOAT_T, 16> t;
    std::memcpy(t.data(), tf, sizeof(float) * 16);
    MCLFile(name).write(pts, rg, ps, t, pose7[0], pose7[1], pose7[2], pose7[3], pose7[4], pose7[5], pose7[6]);
    return 0;
  }
  catch (std::exception& ex)
  {
    g_last_error = ex.what();
    return 1;
  }
}

// Two-step read: sizes first (points_out == NULL), then the payload into caller buffers of those sizes.
int ref_mcl_read(const char* name, uint64_t* p, uint64_t* n, float* points_out, int32_t* rings_out, float* particles_out, float tf_out[16],
                 float pose7_out[7])
{
  try
  {
    std::vector<CudaPoint> pts;
    std::vector<int> rg;
    std::vector<Particle> ps;
    std::array<FLOAT_T, 16> t;
    FLOAT_T v[7];
    MCLFile(name).read(pts, rg, ps, t, v[0], v[1], v[2], v[3], v[4], v[5], v[6]);
    *p = pts.size();
    *n = ps.size();
    if (points_out)
    {
      if (!pts.empty()) std::memcpy(points_out, static_cast<void*>(pts.data()), pts.size() * sizeof(CudaPoint));
      for (size_t i = 0; i < rg.size(); ++i) rings_out[i] = rg[i];
      if (!ps.empty()) std::memcpy(particles_out, static_cast<void*>(ps.data()), ps.size() * sizeof(Particle));
      std::memcpy(tf_out, t.data(), sizeof(float) * 16);
      for (int k = 0; k < 7; ++k) pose7_out[k] = v[k];
    }
    return 0;
  }
  catch (std::exception& ex)
  {
    g_last_error = ex.what();
    return 1;
  }
}

#ifndef TSDF_REF_WITH_B200_SHIM
// The scan reduction inside TSDFEvaluator::evaluateParticles (tsdf_evaluator.cpp:304-376), run verbatim on a packed cloud
// (x y z float32 at byte 0/4/8, ring int16 at byte 12, point_step 16) with ignore_tf = true; returns the number of points
// handed to evaluate() and copies at most cap of them (3 floats each). Rings must stay below the reference's 64 buckets.
int64_t ref_reduce_scan(const float* xyz, const int16_t* ring, uint64_t n, float cell_size, float* points_out, uint64_t cap)
{
  try
  {
    const FLOAT_T lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    auto map = std::make_shared<RefMap>(lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], 0.5f, 0.0f);
    CaptureEvaluator ev(map, false, 0.9f, 0.1f, 0.0f, 100.0f, cell_size);
    const sensor_msgs::PointCloud2 cloud = make_cloud(xyz, ring, n);
    ParticleCloud pc;
    ev.evaluateParticles(pc, cloud, "", "", false, true);
    const uint64_t m = ev.captured.size();
    const uint64_t c = m < cap ? m : cap;
    if (points_out && c) std::memcpy(points_out, static_cast<void*>(ev.captured.data()), c * sizeof(CudaPoint));
    return static_cast<int64_t>(m);
  }
  catch (std::exception& ex)
  {
    g_last_error = ex.what();
    return -1;
  }
}
#endif

#ifdef TSDF_REF_WITH_B200_SHIM
// GpuSystematicResampler (the product's Resampler subclass) through the reference's Resampler interface, same seed
// handling as ref_systematic_resample. Needs a live evaluator created by ref_eval_create (it owns the GPU context).
uint64_t ref_gpu_systematic_resample(const float* particles, uint64_t n, uint32_t seed, float* particles_out, uint64_t cap)
{
  try
  {
    ParticleCloud cloud;
    cloud.particles().resize(n);
    std::memcpy(static_cast<void*>(cloud.particles().data()), particles, n * sizeof(Particle));
    GpuSystematicResampler rs;
    rs.seed(seed);
    Resampler& base = rs;
    base.resample(cloud);
    const uint64_t m = cloud.size();
    const uint64_t c = m < cap ? m : cap;
    if (particles_out && c) std::memcpy(particles_out, static_cast<void*>(cloud.particles().data()), c * sizeof(Particle));
    return m;
  }
  catch (std::exception& ex)
  {
    g_last_error = ex.what();
    return ~0ull;
  }
}
// GpuResidualResampler / GpuResidualSystematicResampler (product) through the reference's Resampler interface.
uint64_t ref_gpu_resample_method(int method, const float* particles, uint64_t n, uint32_t seed, float* particles_out, uint64_t cap)
{
  try
  {
    if (method == 1) return run_seeded_resampler<GpuResidualResampler>(particles, n, seed, particles_out, cap);
    if (method == 2) return run_seeded_resampler<GpuResidualSystematicResampler>(particles, n, seed, particles_out, cap);
    if (method == 3) return run_seeded_resampler<GpuWheelResampler>(particles, n, seed, particles_out, cap);
    if (method == 4) return run_seeded_resampler<SeededGpuMetropolis>(particles, n, seed, particles_out, cap);
    if (method == 5) return run_seeded_resampler<GpuRejectionResampler>(particles, n, seed, particles_out, cap);
    g_last_error = "unknown method";
    return ~0ull;
  }
  catch (std::exception& ex)
  {
    g_last_error = ex.what();
    return ~0ull;
  }
}
int ref_has_b200_shim() { return 1; }
#else
int ref_has_b200_shim() { return 0; }
#endif
#ifdef TSDF_REF_WITH_REF_CUDA
int ref_has_reference_cuda() { return 1; }
#else
int ref_has_reference_cuda() { return 0; }
#endif

}  // extern "C"
