// cuda_evaluator_b200.cpp — drop-in replacement for the reference's src/cuda/cuda_evaluator.cu.
//
// Compiled INTO the reference package in place of src/cuda/cuda_evaluator.cu (+ cuda_sum.cu, cuda_util.cu) and linked
// against libtsdfloc.so: it defines the four CudaEvaluator symbols the reference's own header declares
// (include/tsdf_localization/cuda/cuda_evaluator.h:109-121), so TSDFEvaluator (evaluation/tsdf_evaluator.h:78),
// mcl_3d (src/mcl_3d.cpp:749) and num_particles_eval (src/num_particles_eval.cpp:206) compile and link UNCHANGED.
// All numeric work happens behind the C ABI (include/tsdfloc.h); this file only adapts types and error behaviour:
//
//   reference behaviour                                                         here
//   ctor uploads the map, wraps failures in "Error while creating ..."          tsdfloc_create            (cuda_evaluator.cu:21-59)
//   evaluate(): empty scan -> default pose, weights untouched                   same                      (:122-125)
//   evaluate(): normalised weights written to particles[i].second, in order     tsdfloc_sensor_update     (:397-408)
//   evaluate(): "No particle is valid!" when the weight sum is 0                TSDFLOC_E_NO_VALID_PARTICLE -> same text (:366-369)
//   evaluate(): pose = weighted mean xyz + setRPY(atan2 means), covariance 0     same                      (:410-423)
//   any CUDA failure -> std::runtime_error                                      TSDFLOC_E_CUDA -> std::runtime_error(last_error)
//
// The header's private data members stay as they are (binary layout unchanged); the tsdfloc context handle is kept
// in the otherwise unused `d_transform_` pointer. With TSDFLOC_DEVICES="0,1,2,..." (more than one device) the evaluator
// shards every update over those GPUs from this one process (tsdfloc_multi_*; handle in the unused `d_new_weights_` slot);
// results are bit-identical to the single-GPU path.
#include <tsdf_localization/cuda/cuda_evaluator.h>

#include <sensor_msgs/point_cloud2_iterator.h>
#include <tf2_geometry_msgs/tf2_geometry_msgs.h>
#include <tf2_ros/transform_listener.h>

#include <cmath>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "tsdfloc.h"
#include "tsdfloc_shim.h"

namespace tsdf_localization
{

static_assert(sizeof(Particle) == 7 * sizeof(float), "Particle must be 7 packed floats (particle.h:26-55)");
static_assert(sizeof(CudaPoint) == 3 * sizeof(float), "CudaPoint must be 3 packed floats (cuda_evaluator.h:22-39)");

namespace
{
std::mutex g_ctx_mutex;
tsdfloc_ctx* g_last_ctx = nullptr;
tsdfloc_multi* g_last_multi = nullptr;

inline tsdfloc_ctx* ctx_of(FLOAT_T* slot) { return reinterpret_cast<tsdfloc_ctx*>(slot); }
inline tsdfloc_multi* multi_of(FLOAT_T* slot) { return reinterpret_cast<tsdfloc_multi*>(slot); }

// "0,1,2" -> {0, 1, 2}
std::vector<int> device_list(const char* text)
{
  std::vector<int> out;
  std::string cur;
  for (const char* p = text; ; ++p)
  {
    if (*p == ',' || *p == '\0')
    {
      if (!cur.empty()) out.push_back(std::atoi(cur.c_str()));
      cur.clear();
      if (*p == '\0') break;
    }
    else cur.push_back(*p);
  }
  return out;
}
}  // namespace

tsdfloc_ctx* tsdfloc_shim_context()
{
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  return g_last_ctx;
}

tsdfloc_multi* tsdfloc_shim_multi()
{
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  return g_last_multi;
}

CudaEvaluator::CudaEvaluator(CudaSubVoxelMap<FLOAT_T, FLOAT_T>& map, bool per_point, FLOAT_T a_hit, FLOAT_T a_range, FLOAT_T a_max, FLOAT_T max_range)
: d_map_(nullptr), per_point_(per_point), d_grid_occ_(nullptr), d_data_(nullptr), d_particles_(nullptr), d_particles_ordered_(nullptr),
  particles_reserved_(0), d_points_(nullptr), d_points_ordered_(nullptr), points_reserved_(0), d_transform_(nullptr), d_new_weights_(nullptr),
  d_point_weights_(nullptr), point_weights_size_(0), p_x_(nullptr), p_y_(nullptr), p_z_(nullptr), sin_a_(nullptr), cos_a_(nullptr),
  sin_b_(nullptr), cos_b_(nullptr), sin_c_(nullptr), cos_c_(nullptr), a_hit_(a_hit), a_range_(a_range), a_max_(a_max), max_range_(max_range),
  inv_max_range_(1.0 / max_range), max_range_squared_(max_range * max_range)
{
  const auto& c = map.coef();
  tsdfloc_map_desc d{};
  d.dim[0] = c.dim_x_; d.dim[1] = c.dim_y_; d.dim[2] = c.dim_z_;
  d.min[0] = c.min_x_; d.min[1] = c.min_y_; d.min[2] = c.min_z_;
  d.max[0] = c.max_x_; d.max[1] = c.max_y_; d.max[2] = c.max_z_;
  d.resolution = c.resolution_;
  d.init_value = c.init_value_;
  d.up_dim[0] = c.up_dim_x_; d.up_dim[1] = c.up_dim_y_; d.up_dim[2] = c.up_dim_z_;
  d.up_dim_2 = c.up_dim_2_;
  d.sub_dim = c.sub_dim_;
  d.sub_dim_2 = c.sub_dim_2_;
  d.grid_occ_size = c.grid_occ_size_;
  d.data_size = c.data_size_;

  tsdfloc_params prm{};
  tsdfloc_default_params(&prm);
  prm.a_hit = a_hit;
  prm.a_range = a_range;
  prm.a_max = a_max;
  prm.max_range = max_range;
  prm.per_point = per_point ? 1 : 0;

  int device = 0;
  if (const char* e = std::getenv("TSDFLOC_DEVICE")) device = std::atoi(e);
  std::vector<int> devices;
  if (const char* e = std::getenv("TSDFLOC_DEVICES")) devices = device_list(e);
  tsdfloc_ctx* ctx = nullptr;
  static_assert(sizeof(OCC_T) == sizeof(int32_t), "OCC_T must be a 32-bit int (cuda_sub_voxel_map.h:13)");
  if (devices.size() > 1)
  {
    tsdfloc_multi* multi = nullptr;
    const int rc = tsdfloc_multi_create(&d, reinterpret_cast<const int32_t*>(map.rawGridOcc()), map.rawData(), &prm, devices.data(),
                                        static_cast<int>(devices.size()), &multi);
    if (rc != TSDFLOC_OK)
      throw std::runtime_error(std::string("Error while creating the CUDA context for the map! ") + tsdfloc_multi_last_error(nullptr));
    d_new_weights_ = reinterpret_cast<FLOAT_T*>(multi);
    ctx = tsdfloc_multi_ctx(multi, 0);   // resampler / reduction entry points keep using one device's context
  }
  else
  {
    if (devices.size() == 1) device = devices[0];
    const int rc = tsdfloc_create(&d, reinterpret_cast<const int32_t*>(map.rawGridOcc()), map.rawData(), &prm, device, &ctx);
    if (rc != TSDFLOC_OK)
    {
      // same wrapper text as cuda_evaluator.cu:52-55, with the cause appended
      throw std::runtime_error(std::string("Error while creating the CUDA context for the map! ") + tsdfloc_last_error(nullptr));
    }
  }
  d_map_ = &map;
  d_transform_ = reinterpret_cast<FLOAT_T*>(ctx);
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  g_last_ctx = ctx;
  g_last_multi = d_new_weights_ ? multi_of(d_new_weights_) : nullptr;
}

CudaEvaluator::~CudaEvaluator()
{
  tsdfloc_ctx* ctx = ctx_of(d_transform_);
  {
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    if (g_last_ctx == ctx)
    {
      g_last_ctx = nullptr;
      g_last_multi = nullptr;
    }
  }
  if (d_new_weights_)
  {
    tsdfloc_multi_destroy(multi_of(d_new_weights_));   // owns the per-device contexts, incl. ctx
    d_new_weights_ = nullptr;
  }
  else
  {
    tsdfloc_destroy(ctx);
  }
  d_transform_ = nullptr;
}

namespace
{
// Byte offset of a named PointCloud2 field; throws like sensor_msgs::PointCloud2ConstIterator does for a missing field.
uint32_t field_offset(const sensor_msgs::PointCloud2& cloud, const std::string& name)
{
  for (const auto& f : cloud.fields)
    if (f.name == name) return f.offset;
  throw std::runtime_error("Field " + name + " does not exist");
}

void throw_multi_error(tsdfloc_multi* multi, int rc)
{
  if (rc == TSDFLOC_E_NO_VALID_PARTICLE) throw std::runtime_error("No particle is valid!");
  throw std::runtime_error(std::string("Error occured during the sensor update on the gpu! ") + tsdfloc_multi_last_error(multi));
}

void throw_update_error(tsdfloc_ctx* ctx, int rc)
{
  if (rc == TSDFLOC_E_NO_VALID_PARTICLE) throw std::runtime_error("No particle is valid!");
  throw std::runtime_error(std::string("Error occured during the sensor update on the gpu! ") + tsdfloc_last_error(ctx));
}
}  // namespace

geometry_msgs::PoseWithCovariance tsdfloc_shim_pose(const float mean[6])
{
  geometry_msgs::PoseWithCovariance average_pose;
  average_pose.pose.position.x = mean[0];
  average_pose.pose.position.y = mean[1];
  average_pose.pose.position.z = mean[2];
  // tf2::Quaternion::setRPY(roll, pitch, yaw) written out (ZYX half-angle formula), as cuda_evaluator.cu:414-416 uses it
  const double hr = 0.5 * mean[3], hp = 0.5 * mean[4], hy = 0.5 * mean[5];
  const double cr = std::cos(hr), sr = std::sin(hr), cp = std::cos(hp), sp = std::sin(hp), cy = std::cos(hy), sy = std::sin(hy);
  average_pose.pose.orientation.x = sr * cp * cy - cr * sp * sy;
  average_pose.pose.orientation.y = cr * sp * cy + sr * cp * sy;
  average_pose.pose.orientation.z = cr * cp * sy - sr * sp * cy;
  average_pose.pose.orientation.w = cr * cp * cy + sr * sp * sy;
  for (auto& v : average_pose.covariance) v = 0.0;  // the reference leaves all six variances at 0 (cuda_evaluator.cu:418-423)
  return average_pose;
}

// PointCloud2 overload (cuda_evaluator.cu:78-116): the ring-agnostic 6.4 cm cell-CENTRE reduction, then the evaluation.
// The reference builds a ring-ordered multimap and an unordered_set on the host; here the raw cloud goes to the device
// and is reduced there (TSDFLOC_REDUCE_EMIT_CENTRES, one ring): the same SET of centres, emitted in cloud order (the
// reference's order is its unordered_set's iteration order). No caller in the reference; kept for interface completeness.
geometry_msgs::PoseWithCovariance CudaEvaluator::evaluate(std::vector<Particle>& particles, const sensor_msgs::PointCloud2& real_cloud, FLOAT_T tf_matrix[16])
{
  const uint64_t n_points = static_cast<uint64_t>(real_cloud.width) * real_cloud.height;
  if (n_points == 0) return geometry_msgs::PoseWithCovariance();
  field_offset(real_cloud, "ring");  // the reference's iterator throws when the field is missing
  tsdfloc_ctx* ctx = ctx_of(d_transform_);
  float mean[6] = {0, 0, 0, 0, 0, 0};
  if (d_new_weights_)   // TSDFLOC_DEVICES names several GPUs: reduce on the first, evaluate on all
  {
    tsdfloc_multi* multi = multi_of(d_new_weights_);
    const int rc = tsdfloc_multi_sensor_update_cloud(multi, reinterpret_cast<float*>(particles.data()), particles.size(),
                                                     real_cloud.data.data() + field_offset(real_cloud, "x"), real_cloud.point_step, nullptr, 0, 4,
                                                     n_points, 0.064, 1, TSDFLOC_REDUCE_EMIT_CENTRES, tf_matrix, mean, nullptr);
    if (rc == TSDFLOC_E_EMPTY_SCAN) return geometry_msgs::PoseWithCovariance();
    if (rc != TSDFLOC_OK) throw_multi_error(multi, rc);
    return tsdfloc_shim_pose(mean);
  }
  const int rc = tsdfloc_sensor_update_cloud(ctx, reinterpret_cast<float*>(particles.data()), particles.size(),
                                             real_cloud.data.data() + field_offset(real_cloud, "x"), real_cloud.point_step, nullptr, 0, 4,
                                             n_points, 0.064, 1, TSDFLOC_REDUCE_EMIT_CENTRES, tf_matrix, mean, nullptr);
  if (rc == TSDFLOC_E_EMPTY_SCAN) return geometry_msgs::PoseWithCovariance();
  if (rc != TSDFLOC_OK) throw_update_error(ctx, rc);
  return tsdfloc_shim_pose(mean);
}

// TSDFEvaluator::evaluateParticles with the reduction on the GPU (tsdf_evaluator.cpp:247-378).
geometry_msgs::PoseWithCovariance TSDFEvaluatorB200::evaluateParticles(ParticleCloud& particle_cloud, const sensor_msgs::PointCloud2& real_cloud,
                                                                       const std::string& robot_frame, const std::string& scan_frame,
                                                                       bool use_cuda, bool ignore_tf)
{
  if (!use_cuda) return TSDFEvaluator::evaluateParticles(particle_cloud, real_cloud, robot_frame, scan_frame, use_cuda, ignore_tf);
  if (!ctx_) throw std::runtime_error("TSDFEvaluatorB200: no CudaEvaluator context alive");

  FLOAT_T tf_matrix[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  if (!ignore_tf)
  {
    // scanner -> robot transform, as tsdf_evaluator.cpp:251-287
    static tf2_ros::Buffer tf_buffer;
    static tf2_ros::TransformListener tf_listener(tf_buffer);
    geometry_msgs::TransformStamped scan_to_base = tf_buffer.lookupTransform(robot_frame, scan_frame, real_cloud.header.stamp, ros::Duration(0.5));
    tf2::Transform t;
    tf2::convert(scan_to_base.transform, t);
    for (int r = 0; r < 3; ++r)
    {
      for (int c = 0; c < 3; ++c) tf_matrix[4 * r + c] = t.getBasis()[r][c];
    }
    tf_matrix[3] = t.getOrigin().getX();
    tf_matrix[7] = t.getOrigin().getY();
    tf_matrix[11] = t.getOrigin().getZ();
  }

  const uint64_t n_points = static_cast<uint64_t>(real_cloud.width) * real_cloud.height;
  const uint8_t* base = real_cloud.data.data();
  std::vector<Particle>& particles = particle_cloud.particles();
  float mean[6] = {0, 0, 0, 0, 0, 0};
  last_reduced_ = 0;
  // x y z are read as three consecutive floats at the "x" field (iter_x[0..2], :311-315), the ring as a short (:305)
  if (multi_)   // TSDFLOC_DEVICES names several GPUs: reduce on the first, evaluate on all
  {
    const int rc = tsdfloc_multi_sensor_update_cloud(multi_, reinterpret_cast<float*>(particles.data()), particles.size(),
                                                     base + field_offset(real_cloud, "x"), real_cloud.point_step,
                                                     base + field_offset(real_cloud, "ring"), real_cloud.point_step, 2, n_points, cell_, n_rings,
                                                     ring_desync_like_reference ? TSDFLOC_REDUCE_RING_DESYNC_LIKE_REFERENCE : 0u, tf_matrix,
                                                     mean, &last_reduced_);
    if (rc == TSDFLOC_E_EMPTY_SCAN) return geometry_msgs::PoseWithCovariance();
    if (rc != TSDFLOC_OK) throw_multi_error(multi_, rc);
    return tsdfloc_shim_pose(mean);
  }
  const int rc = tsdfloc_sensor_update_cloud(ctx_, reinterpret_cast<float*>(particles.data()), particles.size(),
                                             base + field_offset(real_cloud, "x"), real_cloud.point_step,
                                             base + field_offset(real_cloud, "ring"), real_cloud.point_step, 2, n_points, cell_, n_rings,
                                             ring_desync_like_reference ? TSDFLOC_REDUCE_RING_DESYNC_LIKE_REFERENCE : 0u, tf_matrix, mean,
                                             &last_reduced_);
  if (rc == TSDFLOC_E_EMPTY_SCAN) return geometry_msgs::PoseWithCovariance();  // evaluate() on an empty scan, cuda_evaluator.cu:122-125
  if (rc != TSDFLOC_OK) throw_update_error(ctx_, rc);
  return tsdfloc_shim_pose(mean);
}

geometry_msgs::PoseWithCovariance CudaEvaluator::evaluate(std::vector<Particle>& particles, const std::vector<CudaPoint>& points, FLOAT_T tf_matrix[16])
{
  if (points.size() == 0)
  {
    return geometry_msgs::PoseWithCovariance();
  }
  tsdfloc_ctx* ctx = ctx_of(d_transform_);
  float mean[6] = {0, 0, 0, 0, 0, 0};
  if (d_new_weights_)
  {
    tsdfloc_multi* multi = multi_of(d_new_weights_);
    const int rc = tsdfloc_multi_sensor_update(multi, reinterpret_cast<float*>(particles.data()), particles.size(),
                                               reinterpret_cast<const float*>(points.data()), points.size(), tf_matrix, mean);
    if (rc == TSDFLOC_E_NO_VALID_PARTICLE) throw std::runtime_error("No particle is valid!");
    if (rc != TSDFLOC_OK)
      throw std::runtime_error(std::string("Error occured during the sensor update on the gpu! ") + tsdfloc_multi_last_error(multi));
    return tsdfloc_shim_pose(mean);
  }
  const int rc = tsdfloc_sensor_update(ctx, reinterpret_cast<float*>(particles.data()), particles.size(),
                                       reinterpret_cast<const float*>(points.data()), points.size(), tf_matrix, mean);
  if (rc != TSDFLOC_OK) throw_update_error(ctx, rc);
  return tsdfloc_shim_pose(mean);
}

}  // namespace tsdf_localization
