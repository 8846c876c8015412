// host_resample.cpp — the sequential halves of the reference's Residual and ResidualSystematic resamplers
// (include/tsdf_localization/resampling/novel_resampling.h:9-36, 76-104). Both are recurrences with no exact parallel form:
// ResidualSystematic carries an fp32 remainder `u` from particle to particle (every step rounds), Residual consumes a random
// index stream until the output is full. They run here, on the host, over the N weights (4 B per particle); what they
// produce — how many copies of which particle, in output order — is expanded on the device (k_expand_runs), so the 28 B
// particles never leave the GPU. Operand types follow the reference statement by statement (`auto` there resolves to float
// for u / temp / expected_insertions / insertions, to double for temp + 1.0). Built with -ffp-contract=off.
#include "../../include/tsdfloc.h"

#include <cmath>
#include <cstddef>

extern "C" int tsdfloc_residual_systematic_counts(const float* weights, uint64_t stride, uint64_t n, float u0, uint32_t* counts, uint64_t* total)
{
  if (!weights || !counts || stride == 0) return TSDFLOC_E_BAD_ARG;
  if (n > (1ull << 24)) return TSDFLOC_E_BAD_ARG;   // size_t -> float conversions below are exact up to 2^24
  float u = u0;                                     // auto u = uniform_distribution(*m_generator_ptr)          (:86)
  const float size_f = static_cast<float>(n);       // particle_cloud.size() * particle.second: size_t -> float  (:91)
  uint64_t sum = 0;
  for (uint64_t m = 0; m < n; ++m)
  {
    const float w = weights[m * stride];
    const float temp = size_f * w - u;              // two fp32 roundings                                        (:91)
    const double t1 = static_cast<double>(temp) + 1.0;                                                        // (:92)
    if (!(t1 >= 0.0) || t1 >= 4294967296.0) return TSDFLOC_E_BAD_ARG;   // negative / NaN weight: the reference's cast is undefined
    const uint64_t o = static_cast<uint64_t>(t1);   // static_cast<size_t>(temp + 1.0)                           (:92)
    u = static_cast<float>(o) - temp;               // size_t - float -> float                                   (:93)
    counts[m] = static_cast<uint32_t>(o);
    sum += o;
  }
  if (total) *total = sum;
  return TSDFLOC_OK;
}

extern "C" int tsdfloc_residual_runs(const float* weights, uint64_t stride, uint64_t n, tsdfloc_index_draw_fn draw, void* user,
                                     uint64_t max_draws, uint32_t* run_parent, uint32_t* run_count, uint64_t run_cap, uint64_t* n_runs,
                                     uint64_t* n_draws)
{
  if (!weights || !draw || !run_parent || !run_count || !n_runs || stride == 0 || n == 0) return TSDFLOC_E_BAD_ARG;
  if (n > (1ull << 24)) return TSDFLOC_E_BAD_ARG;
  const float size_f = static_cast<float>(n);
  uint64_t filled = 0, runs = 0, draws = 0;
  while (filled < n)                                                  // while (new_particles.size() < particle_cloud.size())  (:19)
  {
    if (draws >= max_draws) { *n_runs = runs; if (n_draws) *n_draws = draws; return TSDFLOC_E_CAPACITY; }
    const uint64_t idx = draw(user);                                  // uniform_distribution(*m_generator_ptr)                (:21)
    ++draws;
    if (idx >= n) return TSDFLOC_E_BAD_ARG;
    const float expected = weights[idx * stride] * size_f;            // float * size_t -> float                               (:23)
    const float left = static_cast<float>(n - filled);                // size_t, converted where it meets the float            (:24-25)
    const float insertions = expected <= left ? expected : left;      //                                                       (:25)
    // for (size_t index = 0; index < insertions; ++index): the smallest k with (float)k >= insertions copies             (:27-30)
    uint64_t k = 0;
    if (insertions > 0.0f)
    {
      const float c = std::ceil(insertions);
      k = static_cast<uint64_t>(c);
    }
    if (k == 0) continue;
    if (k > n - filled) k = n - filled;                               // cannot trigger (insertions <= left, left integral); defensive
    if (runs >= run_cap) { *n_runs = runs; if (n_draws) *n_draws = draws; return TSDFLOC_E_CAPACITY; }
    run_parent[runs] = static_cast<uint32_t>(idx);
    run_count[runs] = static_cast<uint32_t>(k);
    ++runs;
    filled += k;
  }
  *n_runs = runs;
  if (n_draws) *n_draws = draws;
  return TSDFLOC_OK;
}
