// host_resample.cpp — the host halves of the reference's resamplers other than Systematic: Residual and ResidualSystematic
// (include/tsdf_localization/resampling/novel_resampling.h:9-36, 76-104) first, Wheel / Metropolis / Rejection below. Both are recurrences with no exact parallel form:
// ResidualSystematic carries an fp32 remainder `u` from particle to particle (every step rounds), Residual consumes a random
// index stream until the output is full. They run here, on the host, over the N weights (4 B per particle); what they
// produce — how many copies of which particle, in output order — is expanded on the device (k_expand_runs), so the 28 B
// particles never leave the GPU. Operand types follow the reference statement by statement (`auto` there resolves to float
// for u / temp / expected_insertions / insertions, to double for temp + 1.0). Built with -ffp-contract=off.
#include "../../include/tsdfloc.h"

#include <cmath>
#include <cstddef>
#include <vector>

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

// ---- Wheel / Metropolis / Rejection: the remaining choices of mcl_3d's resampling_method switch (src/mcl_3d.cpp:243-263).
// Each output slot is decided by random draws; the draws come through callbacks from the caller's generator, the decisions
// are made here (one parent per slot), the copies on the device.

extern "C" int tsdfloc_wheel_parents(const float* weights, uint64_t stride, uint64_t n, tsdfloc_real_draw_fn real, void* user, uint32_t* parents)
{
  if (!weights || !real || !parents || stride == 0 || n == 0) return TSDFLOC_E_BAD_ARG;
  if (n > (1ull << 24)) return TSDFLOC_E_BAD_ARG;
  // wheel_resampler.cpp:16-29 restarts `FLOAT_T weight_sum = 0.0` for every output slot and adds the weights in index order,
  // so every slot walks the same fp32 running sums; slot i takes the first index whose sum s satisfies u <= s. With
  // reach[k] = max(s_0 .. s_k) (NaN sums never satisfy the test and are skipped) that index is the first k with
  // u <= reach[k], and reach is non-decreasing whatever the signs of the weights.
  std::vector<float> reach(n);
  float sum = 0.0f;                                                    // FLOAT_T weight_sum = 0.0;               (:16)
  float top = -INFINITY;
  for (uint64_t k = 0; k < n; ++k)
  {
    sum += weights[k * stride];                                        // weight_sum += current_weight;          (:21-22)
    if (sum > top) top = sum;
    reach[k] = top;
  }
  // guide[b] = first k with reach[k] >= b / n, b = 0 .. n (n when there is none): u in [b/n, (b+1)/n) finds its index inside
  // [guide[b], guide[b + 1]]. reach * n and u * n are exact in fp64 (24-bit significands, n <= 2^24).
  const double nd = static_cast<double>(n);
  std::vector<uint32_t> guide(n + 2);
  {
    uint64_t k = 0;
    for (uint64_t b = 0; b <= n; ++b)
    {
      while (k < n && !(static_cast<double>(reach[k]) * nd >= static_cast<double>(b))) ++k;
      guide[b] = static_cast<uint32_t>(k);
    }
    guide[n + 1] = static_cast<uint32_t>(n);
  }
  for (uint64_t i = 0; i < n; ++i)
  {
    const float u = real(user);                                        // FLOAT_T random_value = uniform_distribution(gen);  (:15)
    uint64_t lo = 0, hi = n;                                           // the answer lies in [lo, hi]; n: no sum reaches u
    if (u != u) lo = n;                                                // a NaN draw satisfies no test (a std distribution never makes one)
    else if (u >= 0.0f)
    {
      const double scaled = static_cast<double>(u) * nd;
      const uint64_t b = scaled >= nd ? n : static_cast<uint64_t>(scaled);
      lo = guide[b];
      hi = guide[b + 1];
    }
    while (lo < hi)                                                    // first k in [lo, hi) with u <= reach[k], else hi
    {
      const uint64_t mid = lo + ((hi - lo) >> 1);
      if (u <= reach[mid]) hi = mid; else lo = mid + 1;
    }
    // lo == hi: either reach[hi] >= (b + 1) / n > u, or hi == n and no sum reaches u: the slot keeps its own particle (:31 is
    // commented out in the reference)
    parents[i] = static_cast<uint32_t>(lo < n ? lo : i);
  }
  return TSDFLOC_OK;
}

extern "C" int tsdfloc_metropolis_parents(const float* weights, uint64_t stride, uint64_t n, uint64_t steps, tsdfloc_real_draw_fn real,
                                          tsdfloc_index_draw_fn index, void* user, uint32_t* parents)
{
  if (!weights || !real || !index || !parents || stride == 0 || n == 0) return TSDFLOC_E_BAD_ARG;
  if (n > (1ull << 24)) return TSDFLOC_E_BAD_ARG;
  const float w_k = weights[0];                                        // auto& particle_k = particle_cloud[k] with k == 0: bound
                                                                       // once per slot, never rebound                     (:123-125)
  for (uint64_t i = 0; i < n; ++i)
  {
    uint64_t k = 0;                                                    // auto k = 0;                                       (:123)
    for (uint64_t s = 0; s < steps; ++s)                               // for (auto n = 0u; n < sampling_steps_; ++n)       (:127)
    {
      const float u = real(user);                                      // u first, then j                                   (:129-130)
      const uint64_t j = index(user);
      if (j >= n) return TSDFLOC_E_BAD_ARG;
      const float ratio = weights[j * stride] / w_k;                   // float / float                                     (:133)
      if (u <= ratio) k = j;                                           //                                                   (:133-136)
    }
    parents[i] = static_cast<uint32_t>(k);                             // new_particles.push_back(particle_cloud[k]);       (:139)
  }
  return TSDFLOC_OK;
}

extern "C" int tsdfloc_rejection_parents(const float* weights, uint64_t stride, uint64_t n, tsdfloc_real_draw_fn real, tsdfloc_index_draw_fn index,
                                         void* user, uint64_t max_draws, uint32_t* parents, uint64_t* n_draws)
{
  if (!weights || !real || !index || !parents || stride == 0 || n == 0) return TSDFLOC_E_BAD_ARG;
  if (n > (1ull << 24)) return TSDFLOC_E_BAD_ARG;
  double sup_w = 0.0;                                                  // auto sup_w = 0.0;  — a double                     (:157)
  for (uint64_t k = 0; k < n; ++k)
  {
    const float w = weights[k * stride];
    if (sup_w < w) sup_w = w;                                          //                                                   (:163-166)
  }
  uint64_t draws = 0;
  for (uint64_t i = 0; i < n; ++i)
  {
    uint64_t j = i;                                                    // auto j = i;                                       (:171)
    float u = real(user);                                              //                                                   (:172)
    while (static_cast<double>(u) > static_cast<double>(weights[j * stride]) / sup_w)   // float / double -> double        (:174)
    {
      if (max_draws && draws >= max_draws) { if (n_draws) *n_draws = draws; return TSDFLOC_E_CAPACITY; }
      j = index(user);                                                 // j first, then u                                   (:176-177)
      ++draws;
      if (j >= n) return TSDFLOC_E_BAD_ARG;
      u = real(user);
    }
    parents[i] = static_cast<uint32_t>(j);                             //                                                   (:180)
  }
  if (n_draws) *n_draws = draws;
  return TSDFLOC_OK;
}
