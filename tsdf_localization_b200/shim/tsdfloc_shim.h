// tsdfloc_shim.h — C++ glue between the reference's class interfaces and the C ABI (include/tsdfloc.h).
//
//   * cuda_evaluator_b200.cpp defines the reference's own `CudaEvaluator` (header unchanged) on top of libtsdfloc.so.
//   * TSDFEvaluatorB200 overrides the reference's virtual TSDFEvaluator::evaluateParticles (evaluation/tsdf_evaluator.h:102,
//     src/evaluation/tsdf_evaluator.cpp:247-378) so that the scan reduction (:304-376) also runs on the GPU and the reduced
//     scan never visits the host: construct it in mcl_3d instead of TSDFEvaluator (src/mcl_3d.cpp:749).
//   * GpuSystematicResampler plugs into the reference's `Resampler` interface (resampling/resampler.h:16-31) exactly
//     like SystematicResampler (resampling/novel_resampling.h:38-74): select it in mcl_3d's reconfigure callback
//     (src/mcl_3d.cpp:243-263) instead of `new SystematicResampler()`.
#pragma once

#include <tsdf_localization/evaluation/tsdf_evaluator.h>
#include <tsdf_localization/particle_cloud.h>
#include <tsdf_localization/resampling/resampler.h>

#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "tsdfloc.h"

namespace tsdf_localization
{

// Context of the most recently constructed CudaEvaluator (the reference keeps one per process, cuda_data.h:17-26);
// nullptr when none is alive.
tsdfloc_ctx* tsdfloc_shim_context();
// The multi-GPU handle of that evaluator when TSDFLOC_DEVICES names several devices, else nullptr.
tsdfloc_multi* tsdfloc_shim_multi();

// geometry_msgs::PoseWithCovariance from (x y z roll pitch yaw): position + setRPY quaternion, covariance 0, exactly what
// CudaEvaluator::evaluate returns (src/cuda/cuda_evaluator.cu:410-423).
geometry_msgs::PoseWithCovariance tsdfloc_shim_pose(const float mean[6]);

class TSDFEvaluatorB200 : public TSDFEvaluator
{
public:
  TSDFEvaluatorB200(const std::shared_ptr<CudaSubVoxelMap<FLOAT_T, FLOAT_T>>& map_ptr, bool per_point = false, FLOAT_T a_hit = 0.9,
                    FLOAT_T a_range = 0.1, FLOAT_T a_max = 0.0, FLOAT_T max_range = 100.0, FLOAT_T reduction_cell_size = 0.064)
  : TSDFEvaluator(map_ptr, per_point, a_hit, a_range, a_max, max_range, reduction_cell_size), cell_(reduction_cell_size),
    ctx_(tsdfloc_shim_context()),  // the CudaEvaluator the base class just constructed (tsdf_evaluator.h:78)
    multi_(tsdfloc_shim_multi())
  {
  }

  // use_cuda == false falls through to the reference's own CPU implementation.
  geometry_msgs::PoseWithCovariance evaluateParticles(ParticleCloud& particle_cloud, const sensor_msgs::PointCloud2& real_cloud,
                                                      const std::string& robot_frame = "base_footprint",
                                                      const std::string& scan_frame = "scanner", bool use_cuda = false,
                                                      bool ignore_tf = false) override;

  // Size of the reduced scan of the last GPU evaluateParticles call.
  uint64_t last_reduced_size() const { return last_reduced_; }

  // true: pair survivor #k of the 1 m test with the ring of cloud point #k like the reference (tsdf_evaluator.cpp:319-322
  // skips ++iter_ring); false (default): every point keeps its own ring.
  bool ring_desync_like_reference = false;
  // rings must lie in [0, n_rings); the reference has 64 buckets (:342) and overruns them with an OS1-128.
  uint32_t n_rings = 128;

private:
  FLOAT_T cell_;
  tsdfloc_ctx* ctx_;
  tsdfloc_multi* multi_;
  uint64_t last_reduced_ = 0;
};

class GpuSystematicResampler : public Resampler
{
public:
  // ctx == nullptr: use the live CudaEvaluator's context at resample() time.
  explicit GpuSystematicResampler(tsdfloc_ctx* ctx = nullptr) : ctx_(ctx) {}

  void resample(ParticleCloud& particle_cloud) override
  {
    tsdfloc_ctx* ctx = ctx_ ? ctx_ : tsdfloc_shim_context();
    if (!ctx) throw std::runtime_error("GpuSystematicResampler: no CudaEvaluator context alive");
    const std::size_t n = particle_cloud.size();
    if (n == 0) return;
    // U ~ uniform_real_distribution<FLOAT_T>(0, 1/N) from the base class's mt19937, exactly as novel_resampling.h:43-49
    auto inverse_M = 1.0 / n;
    std::uniform_real_distribution<FLOAT_T> uniform_distribution(0.0, inverse_M);
    const FLOAT_T U = uniform_distribution(*m_generator_ptr);

    std::vector<Particle> new_particles(n + n / 8 + 64);
    uint64_t n_out = 0;
    const int rc = tsdfloc_resample_particles(ctx, reinterpret_cast<const float*>(particle_cloud.particles().data()), n, U,
                                              reinterpret_cast<float*>(new_particles.data()), new_particles.size(), &n_out, nullptr);
    if (rc != TSDFLOC_OK) throw std::runtime_error(std::string("GpuSystematicResampler: ") + tsdfloc_last_error(ctx));
    new_particles.resize(n_out);
    particle_cloud.particles() = std::move(new_particles);
  }

  void seed(uint32_t s) { m_generator_ptr.reset(new std::mt19937(s)); }

private:
  tsdfloc_ctx* ctx_;
};

// The reference's other two resamplers on the same split (host recurrence over the weights, device expansion of the
// particles): drop-ins for `new ResidualResampler()` / `new ResidualSystematicResampler()` in src/mcl_3d.cpp:243-263,765.
// The random draws come from the base class's mt19937 through the very distribution objects the reference constructs
// (novel_resampling.h:14, 81), so equal seeds give equal outputs.
class GpuResidualSystematicResampler : public Resampler
{
public:
  explicit GpuResidualSystematicResampler(tsdfloc_ctx* ctx = nullptr) : ctx_(ctx) {}

  void resample(ParticleCloud& particle_cloud) override
  {
    tsdfloc_ctx* ctx = ctx_ ? ctx_ : tsdfloc_shim_context();
    if (!ctx) throw std::runtime_error("GpuResidualSystematicResampler: no CudaEvaluator context alive");
    const std::size_t n = particle_cloud.size();
    if (n == 0) return;
    std::uniform_real_distribution<FLOAT_T> uniform_distribution(0.0, 1.0);
    const FLOAT_T u = uniform_distribution(*m_generator_ptr);
    std::vector<Particle> new_particles(2 * n + 64);
    uint64_t n_out = 0;
    const int rc = tsdfloc_resample(ctx, TSDFLOC_RESAMPLE_RESIDUAL_SYSTEMATIC, reinterpret_cast<const float*>(particle_cloud.particles().data()),
                                    n, u, nullptr, nullptr, reinterpret_cast<float*>(new_particles.data()), new_particles.size(), &n_out, nullptr);
    if (rc != TSDFLOC_OK) throw std::runtime_error(std::string("GpuResidualSystematicResampler: ") + tsdfloc_last_error(ctx));
    new_particles.resize(n_out);
    particle_cloud.particles() = std::move(new_particles);
  }

  void seed(uint32_t s) { m_generator_ptr.reset(new std::mt19937(s)); }

private:
  tsdfloc_ctx* ctx_;
};

class GpuResidualResampler : public Resampler
{
public:
  explicit GpuResidualResampler(tsdfloc_ctx* ctx = nullptr) : ctx_(ctx) {}

  void resample(ParticleCloud& particle_cloud) override
  {
    tsdfloc_ctx* ctx = ctx_ ? ctx_ : tsdfloc_shim_context();
    if (!ctx) throw std::runtime_error("GpuResidualResampler: no CudaEvaluator context alive");
    const std::size_t n = particle_cloud.size();
    if (n == 0) return;
    Draw d{m_generator_ptr.get(), std::uniform_int_distribution<size_t>(0, n - 1)};
    std::vector<Particle> new_particles(n);
    uint64_t n_out = 0;
    const int rc = tsdfloc_resample(ctx, TSDFLOC_RESAMPLE_RESIDUAL, reinterpret_cast<const float*>(particle_cloud.particles().data()), n, 0.0f,
                                    &Draw::next, &d, reinterpret_cast<float*>(new_particles.data()), new_particles.size(), &n_out, nullptr);
    if (rc != TSDFLOC_OK) throw std::runtime_error(std::string("GpuResidualResampler: ") + tsdfloc_last_error(ctx));
    new_particles.resize(n_out);
    particle_cloud.particles() = std::move(new_particles);
  }

  void seed(uint32_t s) { m_generator_ptr.reset(new std::mt19937(s)); }

private:
  struct Draw
  {
    std::mt19937* gen;
    std::uniform_int_distribution<size_t> dist;
    static uint64_t next(void* self)
    {
      Draw* d = static_cast<Draw*>(self);
      return d->dist(*d->gen);
    }
  };
  tsdfloc_ctx* ctx_;
};

// The remaining three choices of the node's resampling_method switch (src/mcl_3d.cpp:243-263: case 0 Wheel, case 4
// Metropolis(50), default Rejection). Every output slot is decided by draws on the base class's mt19937 through the very
// distribution types the reference constructs (wheel_resampler.cpp:9, novel_resampling.h:115-116, 151-152), handed to the
// library as callbacks; the decisions are made over the weights alone, the particles are copied on the device.
class GpuDrawnResampler : public Resampler
{
public:
  void resample(ParticleCloud& particle_cloud) override
  {
    tsdfloc_ctx* ctx = ctx_ ? ctx_ : tsdfloc_shim_context();
    if (!ctx) throw std::runtime_error(std::string(name_) + ": no CudaEvaluator context alive");
    const std::size_t n = particle_cloud.size();
    if (n == 0) return;
    Draw d{m_generator_ptr.get(), std::uniform_real_distribution<FLOAT_T>(0.0, 1.0), std::uniform_real_distribution<>(0.0, 1.0),
           std::uniform_int_distribution<size_t>(0, n - 1)};
    tsdfloc_draws draws{};
    draws.real = method_ == TSDFLOC_RESAMPLE_WHEEL ? &Draw::next_real_wheel : &Draw::next_real;
    draws.index = &Draw::next_index;
    draws.user = &d;
    draws.metropolis_steps = steps_;
    draws.max_draws = 0;   // like the reference: the loop ends when the draws say so
    std::vector<Particle> new_particles(n);
    uint64_t n_out = 0;
    const int rc = tsdfloc_resample_drawn(ctx, method_, reinterpret_cast<const float*>(particle_cloud.particles().data()), n, &draws,
                                          reinterpret_cast<float*>(new_particles.data()), new_particles.size(), &n_out, nullptr);
    if (rc != TSDFLOC_OK) throw std::runtime_error(std::string(name_) + ": " + tsdfloc_last_error(ctx));
    new_particles.resize(n_out);
    particle_cloud.particles() = std::move(new_particles);
  }

  void seed(uint32_t s) { m_generator_ptr.reset(new std::mt19937(s)); }

protected:
  GpuDrawnResampler(int method, const char* name, tsdfloc_ctx* ctx, std::size_t steps) : method_(method), name_(name), ctx_(ctx), steps_(steps) {}

private:
  struct Draw
  {
    std::mt19937* gen;
    std::uniform_real_distribution<FLOAT_T> real;   // Metropolis / Rejection: auto u = uniform_real_distribution(*m_generator_ptr)
    std::uniform_real_distribution<> real_wheel;    // Wheel: FLOAT_T random_value = uniform_distribution(*m_generator_ptr)
    std::uniform_int_distribution<size_t> index;
    static float next_real(void* self)
    {
      Draw* d = static_cast<Draw*>(self);
      return d->real(*d->gen);
    }
    static float next_real_wheel(void* self)
    {
      Draw* d = static_cast<Draw*>(self);
      const FLOAT_T random_value = d->real_wheel(*d->gen);
      return random_value;
    }
    static uint64_t next_index(void* self)
    {
      Draw* d = static_cast<Draw*>(self);
      return d->index(*d->gen);
    }
  };
  int method_;
  const char* name_;
  tsdfloc_ctx* ctx_;
  std::size_t steps_;
};

class GpuWheelResampler : public GpuDrawnResampler
{
public:
  explicit GpuWheelResampler(tsdfloc_ctx* ctx = nullptr) : GpuDrawnResampler(TSDFLOC_RESAMPLE_WHEEL, "GpuWheelResampler", ctx, 0) {}
};

class GpuMetropolisResampler : public GpuDrawnResampler
{
public:
  explicit GpuMetropolisResampler(size_t sampling_steps, tsdfloc_ctx* ctx = nullptr)
  : GpuDrawnResampler(TSDFLOC_RESAMPLE_METROPOLIS, "GpuMetropolisResampler", ctx, sampling_steps)
  {
  }
};

class GpuRejectionResampler : public GpuDrawnResampler
{
public:
  explicit GpuRejectionResampler(tsdfloc_ctx* ctx = nullptr) : GpuDrawnResampler(TSDFLOC_RESAMPLE_REJECTION, "GpuRejectionResampler", ctx, 0) {}
};

}  // namespace tsdf_localization
