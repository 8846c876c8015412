// tsdfloc_shim.h — C++ glue between the reference's class interfaces and the C ABI (include/tsdfloc.h).
//
//   * cuda_evaluator_b200.cpp defines the reference's own `CudaEvaluator` (header unchanged) on top of libtsdfloc.so.
//   * GpuSystematicResampler plugs into the reference's `Resampler` interface (resampling/resampler.h:16-31) exactly
//     like SystematicResampler (resampling/novel_resampling.h:38-74): select it in mcl_3d's reconfigure callback
//     (src/mcl_3d.cpp:243-263) instead of `new SystematicResampler()`.
#pragma once

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

}  // namespace tsdf_localization
