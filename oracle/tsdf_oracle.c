/* oracle/tsdf_oracle.c — CPU restatement of the reference's MCL sensor-update path (plain C).
 * TEST INFRASTRUCTURE ONLY — see tsdf_oracle.h for who may use it and for the parity-pinning status.
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off, no -march=native: every a*b+c rounds twice,
 * exactly like the reference CPU build on baseline x86-64). */
#define _GNU_SOURCE
#include "tsdf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- float -> unsigned conversions under the three policies ------------------------------------- */

/* What gcc emits for (size_t)f on x86-64 without AVX-512: cvttss2si on f when f < 2^63 (negative values
 * wrap modulo 2^64, NaN gives 0x8000...0), else cvttss2si(f - 2^63) ^ 2^63 (out of range gives 0). */
static uint64_t cvt_u64_x86(float f)
{
  const float two63 = 9223372036854775808.0f;
  if (f != f) return 0x8000000000000000ull;
  if (f < two63)
  {
    if (f <= -two63) return 0x8000000000000000ull;
    return (uint64_t)(int64_t)f;
  }
  {
    float g = f - two63;
    if (!(g < two63)) return 0ull; /* indefinite ^ 2^63 */
    return ((uint64_t)(int64_t)g) ^ 0x8000000000000000ull;
  }
}

/* CUDA cvt.rzi.u32.f32: saturating, NaN -> 0 */
static uint64_t cvt_u32_sat(float f)
{
  if (!(f > 0.0f)) return 0u;
  if (f >= 4294967296.0f) return 0xffffffffu;
  return (uint32_t)f;
}

/* ---- map ------------------------------------------------------------------------------------------ */

oracle_map* oracle_map_create(const float mn[3], const float mx[3], float resolution, float init_value)
{
  /* cuda_sub_voxel_map.tcc:16-34; sub_voxel_size = 1.0 (cuda_sub_voxel_map.h:15) */
  oracle_map* m = (oracle_map*)calloc(1, sizeof(oracle_map));
  oracle_map_coef* c = &m->coef;
  for (int a = 0; a < 3; ++a)
  {
    c->dim[a] = (uint64_t)ceilf(fabsf(mx[a] - mn[a]) / resolution);
    c->min[a] = mn[a];
    c->max[a] = mx[a];
    c->up_dim[a] = (uint64_t)ceilf(fabsf(mx[a] - mn[a]) / 1.0f);
  }
  c->resolution = resolution;
  c->init_value = init_value;
  c->up_dim_2 = c->up_dim[0] * c->up_dim[1];
  c->sub_dim = (uint64_t)ceilf(1.0f / resolution);
  c->sub_dim_2 = c->sub_dim * c->sub_dim;
  c->grid_occ_size = c->up_dim[0] * c->up_dim[1] * c->up_dim[2];
  c->data_size = 0;
  m->grid_occ = (int32_t*)malloc(sizeof(int32_t) * (c->grid_occ_size ? c->grid_occ_size : 1));
  for (uint64_t i = 0; i < c->grid_occ_size; ++i) m->grid_occ[i] = -1;
  m->data = NULL;
  return m;
}

void oracle_map_destroy(oracle_map* m)
{
  if (!m) return;
  free(m->grid_occ);
  free(m->data);
  free(m);
}

oracle_map* oracle_map_from_arrays(const oracle_map_coef* coef, const int32_t* grid_occ, const float* data)
{
  oracle_map* m = (oracle_map*)calloc(1, sizeof(oracle_map));
  m->coef = *coef;
  m->grid_occ = (int32_t*)malloc(sizeof(int32_t) * (coef->grid_occ_size ? coef->grid_occ_size : 1));
  memcpy(m->grid_occ, grid_occ, sizeof(int32_t) * coef->grid_occ_size);
  m->data = (float*)malloc(sizeof(float) * (coef->data_size ? coef->data_size : 1));
  memcpy(m->data, data, sizeof(float) * coef->data_size);
  return m;
}

uint64_t oracle_get_index(const oracle_map* m, float x, float y, float z, int neg_mode)
{
  /* cuda_sub_voxel_map.tcc:50-137 (host, size_t) == cuda_eval_particles.h:12-67 (device, unsigned int) for
   * non-negative offsets; they differ only in how negative/NaN offsets convert (neg_mode). */
  const oracle_map_coef* c = &m->coef;
  const float p[3] = {x, y, z};
  float off[3];
  uint64_t up[3];
  float sub_pos[3];
  for (int a = 0; a < 3; ++a)
  {
    off[a] = p[a] - c->min[a];
    /* product policy: every offset whose float->unsigned conversion is undefined in the reference (negative, NaN,
     * >= 2^64) is a miss; for all other offsets the three modes are identical */
    if (neg_mode == ORACLE_NEG_AS_MISS && (!(off[a] >= 0.0f) || off[a] >= 18446744073709551616.0f)) return c->data_size;
  }
  for (int a = 0; a < 3; ++a)
  {
    /* tcc:53-60: metre-truncated offset, then "/= resolution" through float */
    uint64_t g = (neg_mode == ORACLE_NEG_REF_DEVICE_SAT) ? cvt_u32_sat(off[a]) : cvt_u64_x86(off[a]);
    float q = (neg_mode == ORACLE_NEG_REF_DEVICE_SAT) ? (float)(uint32_t)g / c->resolution : (float)g / c->resolution;
    g = (neg_mode == ORACLE_NEG_REF_DEVICE_SAT) ? cvt_u32_sat(q) : cvt_u64_x86(q);
    if (g >= c->dim[a]) return c->data_size; /* tcc:62-65 */
  }
  for (int a = 0; a < 3; ++a)
  {
    /* tcc:74-84 */
    float t = off[a] / 1.0f;
    if (neg_mode == ORACLE_NEG_REF_DEVICE_SAT)
    {
      up[a] = cvt_u32_sat(t);
      sub_pos[a] = off[a] - (float)(uint32_t)up[a] * 1.0f;
    }
    else
    {
      up[a] = cvt_u64_x86(t);
      sub_pos[a] = off[a] - (float)up[a] * 1.0f;
    }
  }
  uint64_t up_index = up[0] + up[1] * c->up_dim[0] + up[2] * c->up_dim_2; /* tcc:86 */
  if (neg_mode == ORACLE_NEG_REF_DEVICE_SAT) up_index = (uint32_t)up_index;
  if (up_index >= c->grid_occ_size) return c->data_size; /* tcc:108-111 */
  int32_t sub_index = m->grid_occ[up_index];
  if (sub_index < 0) return c->data_size; /* tcc:120-128 */
  uint64_t sub[3];
  for (int a = 0; a < 3; ++a)
  {
    float q = sub_pos[a] / c->resolution; /* tcc:132-134 */
    sub[a] = (neg_mode == ORACLE_NEG_REF_DEVICE_SAT) ? cvt_u32_sat(q) : cvt_u64_x86(q);
  }
  uint64_t r = (uint64_t)(int64_t)sub_index + sub[0] + sub[1] * c->sub_dim + sub[2] * c->sub_dim_2; /* tcc:136 */
  if (neg_mode == ORACLE_NEG_REF_DEVICE_SAT) r = (uint32_t)r;
  return r;
}

float oracle_get_entry(const oracle_map* m, float x, float y, float z, int neg_mode)
{
  /* tcc:139-157 */
  if (!m->data || !m->grid_occ) return m->coef.init_value;
  uint64_t idx = oracle_get_index(m, x, y, z, neg_mode);
  return idx < m->coef.data_size ? m->data[idx] : m->coef.init_value;
}

void oracle_get_entries(const oracle_map* m, const float* xyz, uint64_t n, int neg_mode, float* out)
{
  for (uint64_t i = 0; i < n; ++i) out[i] = oracle_get_entry(m, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], neg_mode);
}

int oracle_map_set_data(oracle_map* m, const float* cells, uint64_t n)
{
  /* cuda_sub_voxel_map.tcc:170-230 */
  oracle_map_coef* c = &m->coef;
  if (n == 0) return 0;
  for (uint64_t i = 0; i < n; ++i)
  {
    uint64_t up[3];
    for (int a = 0; a < 3; ++a)
    {
      float off = cells[4 * i + a] - c->min[a];
      up[a] = cvt_u64_x86(off / 1.0f);
    }
    uint64_t up_index = up[0] + up[1] * c->up_dim[0] + up[2] * c->up_dim[0] * c->up_dim[1];
    if (up_index >= c->grid_occ_size) return 1; /* "Upper voxel index overflow!" */
    m->grid_occ[up_index] = 0;
  }
  int32_t current = 0;
  const uint64_t sub_size = c->sub_dim * c->sub_dim * c->sub_dim;
  for (uint64_t i = 0; i < c->grid_occ_size; ++i)
  {
    if (m->grid_occ[i] >= 0)
    {
      m->grid_occ[i] = current;
      current += (int32_t)sub_size;
    }
  }
  free(m->data);
  c->data_size = (uint64_t)(int64_t)current;
  m->data = (float*)malloc(sizeof(float) * (c->data_size ? c->data_size : 1));
  for (uint64_t i = 0; i < c->data_size; ++i) m->data[i] = c->init_value;
  for (uint64_t i = 0; i < n; ++i)
  {
    /* setEntry, tcc:159-168: writes data_[getIndex()] unguarded; a cell whose getIndex misses would write
     * data_[data_size] (out of bounds) in the reference. Cells handed to setData come from inside the
     * bounding box, so this does not happen for valid maps; the oracle skips such a write. */
    uint64_t idx = oracle_get_index(m, cells[4 * i], cells[4 * i + 1], cells[4 * i + 2], ORACLE_NEG_REF_HOST_X86);
    if (idx < c->data_size) m->data[idx] = cells[4 * i + 3];
  }
  return 0;
}

/* ---- likelihood LUT -------------------------------------------------------------------------------- */

float oracle_likelihood_init(float sigma)
{
  /* map_util.h:68-71: sigma_quad float; expf(double expr -> float); sqrtf(double expr -> float) */
  float sigma_quad = sigma * sigma;
  float init = expf((float)(-(10.0 * 10.0) / sigma_quad / 2)) / (sqrtf((float)(2 * sigma_quad * M_PI)));
  init = init * init * init;
  return init;
}

float oracle_likelihood_value(float tsdf_mm, float sigma)
{
  /* map_util.h:124-126: value is double (float * 0.001); expf/sqrtf take float arguments; the quotient and
   * the cube are evaluated in double because value is double, then stored into the float tuple (:129). */
  float sigma_quad = sigma * sigma;
  double value = tsdf_mm * 0.001;
  value = expf((float)(-(value * value) / sigma_quad / 2)) / (sqrtf((float)(2 * sigma_quad * M_PI)));
  value = value * value * value;
  return (float)value;
}

/* ---- pose -> matrix -------------------------------------------------------------------------------- */

void oracle_pose_matrix(const float pose6[6], const float tf[16], float out[12])
{
  /* tsdf_evaluator.cpp:102-145. sin/cos are the double versions, rounded to float on assignment. */
  float tp[16];
  float alpha = pose6[3], beta = pose6[4], gamma = pose6[5];
  float sin_alpha = (float)sin((double)alpha);
  float cos_alpha = (float)cos((double)alpha);
  float sin_beta = (float)sin((double)beta);
  float cos_beta = (float)cos((double)beta);
  float sin_gamma = (float)sin((double)gamma);
  float cos_gamma = (float)cos((double)gamma);

  tp[0] = cos_beta * cos_gamma;
  tp[4] = cos_beta * sin_gamma;
  tp[8] = -sin_beta;
  tp[3] = pose6[0];

  tp[1] = sin_alpha * sin_beta * cos_gamma - cos_alpha * sin_gamma;
  tp[5] = sin_alpha * sin_beta * sin_gamma + cos_alpha * cos_gamma;
  tp[9] = sin_alpha * cos_beta;
  tp[7] = pose6[1];

  tp[2] = cos_alpha * sin_beta * cos_gamma + sin_alpha * sin_gamma;
  tp[6] = cos_alpha * sin_beta * sin_gamma - sin_alpha * cos_gamma;
  tp[10] = cos_alpha * cos_beta;
  tp[11] = pose6[2];

  for (int r = 0; r < 3; ++r)
  {
    const float a = tp[4 * r], b = tp[4 * r + 1], c = tp[4 * r + 2], d = tp[4 * r + 3];
    out[4 * r + 0] = a * tf[0] + b * tf[4] + c * tf[8];
    out[4 * r + 1] = a * tf[1] + b * tf[5] + c * tf[9];
    out[4 * r + 2] = a * tf[2] + b * tf[6] + c * tf[10];
    out[4 * r + 3] = a * tf[3] + b * tf[7] + c * tf[11] + d;
  }
}

/* ---- evaluatePose ---------------------------------------------------------------------------------- */

float oracle_pose_weight(const oracle_map* m, const oracle_params* prm, const float pose[12], const float* pts, uint64_t np,
                         int neg_mode, uint32_t* idx_out, uint32_t* hits_out, double* w64_out)
{
  /* tsdf_evaluator.cpp:27-76 */
  const float inv_max_range = (float)(1.0 / prm->max_range);            /* tsdf_evaluator.h:75 */
  const float max_range_squared = prm->max_range * prm->max_range;      /* tsdf_evaluator.h:75 */
  float eval_sum = 0.0f;
  double eval_sum64 = 0.0;
  uint32_t hits = 0;
  for (uint64_t i = 0; i < np; ++i)
  {
    float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    float tx = pose[0] * x + pose[1] * y + pose[2] * z + pose[3];
    float ty = pose[4] * x + pose[5] * y + pose[6] * z + pose[7];
    float tz = pose[8] * x + pose[9] * y + pose[10] * z + pose[11];
    uint64_t idx = oracle_get_index(m, tx, ty, tz, neg_mode);
    float value = (m->data && idx < m->coef.data_size) ? m->data[idx] : m->coef.init_value;
    if (idx < m->coef.data_size) ++hits;
    if (idx_out) idx_out[i] = (uint32_t)(idx < m->coef.data_size ? idx : m->coef.data_size);
    float square_dist = x * x + y * y + z * z;
    if (square_dist < max_range_squared)
      value = prm->a_hit * value + prm->a_range * inv_max_range;
    else
      value = prm->a_hit * value + prm->a_max;
    eval_sum += value;
    eval_sum64 += (double)value;
  }
  if (hits_out) *hits_out = hits;
  if (w64_out) *w64_out = eval_sum64;
  return eval_sum;
}

/* ---- evaluate -------------------------------------------------------------------------------------- */

int oracle_evaluate(const oracle_map* m, const oracle_params* prm, float* P, uint64_t n, const float* pts, uint64_t np,
                    const float tf[16], int neg_mode, float* raw_out, float mean_pose6[6], uint32_t* idx_out,
                    uint32_t* hits_out, float* weight_sum_out)
{
  /* tsdf_evaluator.cpp:85-157 */
  int64_t i;
#pragma omp parallel for schedule(dynamic)
  for (i = 0; i < (int64_t)n; ++i)
  {
    float mat[12];
    oracle_pose_matrix(&P[7 * i], tf, mat);
    P[7 * i + 6] = oracle_pose_weight(m, prm, mat, pts, np, neg_mode, idx_out ? idx_out + (uint64_t)i * np : NULL,
                                      hits_out ? hits_out + i : NULL, NULL);
  }
  float weight_sum = 0.0f;
  for (uint64_t k = 0; k < n; ++k)
  {
    if (raw_out) raw_out[k] = P[7 * k + 6];
    weight_sum += P[7 * k + 6];
  }
  if (weight_sum_out) *weight_sum_out = weight_sum;
  if (weight_sum == 0.0f) return 1; /* :159-162 "No particle is valid!" */

  /* :198-221 */
  float avg[3] = {0, 0, 0};
  float ss[3] = {0, 0, 0}, sc[3] = {0, 0, 0};
  for (uint64_t k = 0; k < n; ++k)
  {
    float* p = &P[7 * k];
    p[6] /= weight_sum;
    avg[0] += p[0] * p[6];
    avg[1] += p[1] * p[6];
    avg[2] += p[2] * p[6];
    for (int a = 0; a < 3; ++a)
    {
      ss[a] += (float)(sin((double)p[3 + a]) * p[6]);
      sc[a] += (float)(cos((double)p[3 + a]) * p[6]);
    }
  }
  if (mean_pose6)
  {
    mean_pose6[0] = avg[0];
    mean_pose6[1] = avg[1];
    mean_pose6[2] = avg[2];
    for (int a = 0; a < 3; ++a) mean_pose6[3 + a] = (float)atan2((double)ss[a], (double)sc[a]);
  }
  return 0;
}

/* ---- systematic resampling ------------------------------------------------------------------------- */

uint64_t oracle_systematic_resample(const float* w, uint64_t n, float u0, uint32_t* parents_out, uint64_t cap)
{
  /* novel_resampling.h:41-72: inverse_M double, U float (float += double rounds to float every step),
   * s double running sum, strict s > U. */
  const double inverse_M = 1.0 / (double)n;
  float U = u0;
  double s = 0.0;
  uint64_t out = 0;
  for (uint64_t mi = 0; mi < n; ++mi)
  {
    s += w[mi];
    while (s > U)
    {
      if (parents_out && out < cap) parents_out[out] = (uint32_t)mi;
      ++out;
      U += inverse_M;
    }
  }
  return out;
}

/* ---- residual-systematic and residual resampling -------------------------------------------------------- */

uint64_t oracle_residual_systematic_resample(const float* w, uint64_t n, float u0, uint32_t* parents_out, uint64_t cap)
{
  /* novel_resampling.h:86-99. `auto u` and `auto temp` are float (size_t * float - float), `temp + 1.0` is double,
   * `u = o - temp` converts the size_t to float. */
  float u = u0;
  uint64_t out = 0;
  for (uint64_t mi = 0; mi < n; ++mi)
  {
    float prod = (float)n * w[mi];
    float temp = prod - u;
    uint64_t o = (uint64_t)((double)temp + 1.0);
    u = (float)o - temp;
    for (uint64_t k = 0; k < o; ++k)
    {
      if (parents_out && out < cap) parents_out[out] = (uint32_t)mi;
      ++out;
    }
  }
  return out;
}

uint64_t oracle_residual_resample(const float* w, uint64_t n, const uint64_t* draws, uint64_t n_draws, uint32_t* parents_out,
                                  uint64_t* draws_used)
{
  /* novel_resampling.h:14-31 with the index draws handed in (the reference takes them from
   * std::uniform_int_distribution<size_t>(0, n-1)). expected_insertions and insertions are float; the copy loop compares the
   * size_t counter with that float. Returns the output length (n, or less if the draws ran out). */
  uint64_t out = 0, d = 0;
  while (out < n && d < n_draws)
  {
    uint64_t idx = draws[d++];
    float expected = w[idx] * (float)n;
    uint64_t left_i = n - out;
    float insertions = expected <= (float)left_i ? expected : (float)left_i;
    for (uint64_t k = 0; (float)k < insertions; ++k)
    {
      if (parents_out) parents_out[out] = (uint32_t)idx;
      ++out;
    }
  }
  if (draws_used) *draws_used = d;
  return out;
}

/* ---- Wheel / Metropolis / Rejection, statement by statement, with the draws handed in as callbacks ------- */

void oracle_wheel_resample(const float* w, uint64_t n, oracle_real_draw_fn real, void* user, uint32_t* parents_out)
{
  /* src/resampling/wheel_resampler.cpp:13-32: one draw per output slot (a double narrowed to FLOAT_T, :15 — the callback
   * returns it narrowed), a fresh fp32 running sum per slot, the first index whose sum reaches the draw; a slot no sum
   * reaches keeps its own particle (:31 is commented out). O(n^2), like the reference. */
  for (uint64_t particle_index = 0; particle_index < n; ++particle_index)
  {
    float random_value = real(user);
    float weight_sum = 0.0f;
    uint32_t parent = (uint32_t)particle_index;
    for (uint64_t index = 0; index < n; ++index)
    {
      float current_weight = w[index];
      weight_sum += current_weight;
      if (random_value <= weight_sum)
      {
        parent = (uint32_t)index;
        break;
      }
    }
    parents_out[particle_index] = parent;
  }
}

void oracle_metropolis_resample(const float* w, uint64_t n, uint64_t steps, oracle_real_draw_fn real, oracle_index_draw_fn index_draw,
                                void* user, uint32_t* parents_out)
{
  /* novel_resampling.h:121-140. `auto& particle_k = particle_cloud[k]` (:125) is bound while k == 0 and a reference cannot
   * be re-seated, so the acceptance ratio is always against particle 0's weight; u is drawn before j (:129-130). */
  for (uint64_t i = 0; i < n; ++i)
  {
    uint64_t k = 0;
    const float* particle_k_second = &w[0];
    for (uint64_t s = 0; s < steps; ++s)
    {
      float u = real(user);
      uint64_t j = index_draw(user);
      if (u <= w[j] / *particle_k_second) k = j;
    }
    parents_out[i] = (uint32_t)k;
  }
}

void oracle_rejection_resample(const float* w, uint64_t n, oracle_real_draw_fn real, oracle_index_draw_fn index_draw, void* user,
                               uint32_t* parents_out)
{
  /* novel_resampling.h:157-181. `auto sup_w = 0.0` is a double, so `second / sup_w` divides in fp64 and the float draw is
   * widened for the comparison; inside the loop j is drawn before u (:176-177). */
  double sup_w = 0.0;
  for (uint64_t index = 0; index < n; ++index)
  {
    float wi = w[index];
    if (sup_w < wi) sup_w = wi;
  }
  for (uint64_t i = 0; i < n; ++i)
  {
    uint64_t j = i;
    float u = real(user);
    while (u > (w[j] / sup_w))
    {
      j = index_draw(user);
      u = real(user);
    }
    parents_out[i] = (uint32_t)j;
  }
}

/* A small deterministic draw source for tests that run where the reference's std::mt19937 is not at hand (splitmix64): equal
 * seeds give equal streams, which is all a restatement-vs-product comparison needs. */
typedef struct oracle_draws
{
  uint64_t state, n, n_real, n_index;
} oracle_draws;

static uint64_t splitmix64(uint64_t* s)
{
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void* oracle_draws_create(uint64_t seed, uint64_t n)
{
  oracle_draws* d = (oracle_draws*)malloc(sizeof(oracle_draws));
  if (!d) return NULL;
  d->state = seed;
  d->n = n ? n : 1;
  d->n_real = d->n_index = 0;
  return d;
}
void oracle_draws_destroy(void* p) { free(p); }
float oracle_draw_real(void* p)
{
  oracle_draws* d = (oracle_draws*)p;
  ++d->n_real;
  return (float)(splitmix64(&d->state) >> 40) * (1.0f / 16777216.0f); /* 24 bits: [0, 1) exactly representable */
}
uint64_t oracle_draw_index(void* p)
{
  oracle_draws* d = (oracle_draws*)p;
  ++d->n_index;
  return splitmix64(&d->state) % d->n;
}
void oracle_draws_used(void* p, uint64_t* n_real, uint64_t* n_index)
{
  oracle_draws* d = (oracle_draws*)p;
  *n_real = d->n_real;
  *n_index = d->n_index;
}

/* ---- scan reduction ---------------------------------------------------------------------------------- */

typedef struct red_key
{
  int32_t ring;
  float c[3];
  uint32_t src;  /* cloud position of the first point with this key */
  uint32_t order; /* running index among surviving points (the reference's `index`) */
} red_key;

static uint64_t red_hash(int32_t ring, const float c[3])
{
  uint32_t b[3];
  memcpy(b, c, sizeof(b));
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)(uint32_t)ring;
  for (int a = 0; a < 3; ++a)
  {
    h ^= b[a];
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
  }
  return h;
}

static int red_cmp(const void* pa, const void* pb)
{
  const red_key* a = (const red_key*)pa;
  const red_key* b = (const red_key*)pb;
  if (a->ring != b->ring) return a->ring < b->ring ? -1 : 1;
  if (a->order != b->order) return a->order < b->order ? -1 : 1;
  return 0;
}

/* shared tail: dedup (first wins) + (ring, order) ordering */
static int64_t red_finish(red_key* keys, uint64_t nk, const float* emit_xyz /* NULL: emit the centres */, float* points_out,
                          uint32_t* src_index_out)
{
  uint64_t cap = 16;
  while (cap < 2 * nk + 1) cap <<= 1;
  int64_t* table = (int64_t*)malloc(sizeof(int64_t) * cap);
  red_key* kept = (red_key*)malloc(sizeof(red_key) * (nk ? nk : 1));
  uint64_t n_kept = 0;
  for (uint64_t i = 0; i < cap; ++i) table[i] = -1;
  for (uint64_t i = 0; i < nk; ++i)
  {
    uint64_t slot = red_hash(keys[i].ring, keys[i].c) & (cap - 1);
    int dup = 0;
    while (table[slot] >= 0)
    {
      const red_key* o = &kept[table[slot]];
      /* SortClass::operator== (cuda_evaluator.h:56-59): ring and the three centre floats compare equal */
      if (o->ring == keys[i].ring && o->c[0] == keys[i].c[0] && o->c[1] == keys[i].c[1] && o->c[2] == keys[i].c[2])
      {
        dup = 1; /* unordered_set::insert keeps the element already present */
        break;
      }
      slot = (slot + 1) & (cap - 1);
    }
    if (!dup)
    {
      table[slot] = (int64_t)n_kept;
      kept[n_kept++] = keys[i];
    }
  }
  qsort(kept, n_kept, sizeof(red_key), red_cmp);
  for (uint64_t j = 0; j < n_kept; ++j)
  {
    if (points_out)
    {
      const float* src = emit_xyz ? emit_xyz + 3ull * kept[j].src : kept[j].c;
      points_out[3 * j + 0] = src[0];
      points_out[3 * j + 1] = src[1];
      points_out[3 * j + 2] = src[2];
    }
    if (src_index_out) src_index_out[j] = kept[j].src;
  }
  free(table);
  free(kept);
  return (int64_t)n_kept;
}

int64_t oracle_reduce_scan(const float* P, const int32_t* ring, uint64_t n, float cell, uint32_t n_rings, int ring_desync,
                           float* points_out, uint32_t* src_index_out)
{
  const float res = cell;        /* map_res_ (tsdf_evaluator.h:76) */
  const float half = cell / 2;   /* map_res_half_ */
  red_key* keys = (red_key*)malloc(sizeof(red_key) * (n ? n : 1));
  uint64_t nk = 0;
  uint32_t index = 0; /* tsdf_evaluator.cpp:309 */
  int64_t result = 0;
  for (uint64_t i = 0; i < n; ++i)
  {
    const float x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
    if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue; /* defined divergence, see header */
    const float dist = sqrtf(x * x + y * y + z * z); /* :317 */
    if (dist < 1.0) continue;                        /* :319-322 (index and the ring iterator stay) */
    const int32_t r = ring_desync ? ring[index] : ring[i];
    if (r < 0 || (uint32_t)r >= n_rings)
    {
      result = -1;
      break;
    }
    const float cx = floorf(x / res) * res + half; /* :324-326, fp32, every operation rounded (-ffp-contract=off) */
    const float cy = floorf(y / res) * res + half;
    const float cz = floorf(z / res) * res + half;
    if (isfinite(cx) && isfinite(cy) && isfinite(cz)) /* an overflowing centre drops the point (policy, like NaN input) */
    {
      red_key* k = &keys[nk++];
      k->ring = r;
      k->c[0] = cx;
      k->c[1] = cy;
      k->c[2] = cz;
      k->src = (uint32_t)i;
      k->order = index;
    }
    ++index; /* :331-332 */
  }
  if (result == 0) result = red_finish(keys, nk, P, points_out, src_index_out);
  free(keys);
  return result;
}

int64_t oracle_reduce_scan_centres(const float* P, const int32_t* ring, uint64_t n, double cell, uint32_t n_rings, float* points_out,
                                   uint32_t* src_index_out)
{
  const double half = cell / 2; /* 0.032 for the reference's literal 0.064 */
  red_key* keys = (red_key*)malloc(sizeof(red_key) * (n ? n : 1));
  uint64_t nk = 0;
  int64_t result = 0;
  for (uint64_t i = 0; i < n; ++i)
  {
    const float x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
    if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue;
    const int32_t r = ring ? ring[i] : 0;
    if (r < 0 || (uint32_t)r >= n_rings)
    {
      result = -1;
      break;
    }
    const float cx = (float)(floor((double)x / cell) * cell + half); /* cuda_evaluator.cu:100-102, num_particles_eval.cpp:140-142 */
    const float cy = (float)(floor((double)y / cell) * cell + half);
    const float cz = (float)(floor((double)z / cell) * cell + half);
    if (!(isfinite(cx) && isfinite(cy) && isfinite(cz))) continue;
    red_key* k = &keys[nk++];
    k->ring = r;
    k->c[0] = cx;
    k->c[1] = cy;
    k->c[2] = cz;
    k->src = (uint32_t)i;
    k->order = (uint32_t)i;
  }
  if (result == 0) result = red_finish(keys, nk, NULL, points_out, src_index_out);
  free(keys);
  return result;
}

/* ---- motion update ---------------------------------------------------------------------------------------- */

int oracle_motion_model(int variant, const double in[4], float time_diff, const float a[12], double mean[6], double sigma[6],
                        float ref_pose[6])
{
  for (int k = 0; k < 6; ++k) mean[k] = 0.0;
  if (variant == 1)
  {
    /* particle_cloud.cpp:167-217: linear_velocity / angular_velocity are doubles -> double arithmetic throughout */
    const double lv = in[0], av = in[1];
    if (ref_pose)
    {
      /* :180-182, float = float + double expression */
      const float r0 = (float)(ref_pose[0] + lv * time_diff * cos(ref_pose[5] + (av / 2 * time_diff)));
      const float r1 = (float)(ref_pose[1] + lv * time_diff * sin(ref_pose[5] + (av / 2 * time_diff)));
      const float r5 = (float)(ref_pose[5] + av * time_diff);
      ref_pose[0] = r0;
      ref_pose[1] = r1;
      ref_pose[5] = r5;
    }
    const double d = lv * time_diff, d2 = d * d, th = av * time_diff, t2 = th * th;
    sigma[0] = a[0] * d2 + a[1] * t2;
    sigma[1] = a[2] * d2 + a[3] * t2;
    sigma[2] = a[4] * d2 + a[5] * t2;
    sigma[3] = a[6] * d2 + a[7] * t2;
    sigma[4] = a[8] * d2 + a[9] * t2;
    sigma[5] = a[10] * d2 + a[11] * t2;
    mean[0] = d;
    mean[5] = th;
    return 0;
  }
  if (variant == 0 || variant == 2)
  {
    /* :388-414 / :337-374: FLOAT_T operands -> fp32 arithmetic, each operation rounded */
    const float s0 = (float)in[0], s1 = (float)in[1];
    if (variant == 2 && ref_pose)
    {
      /* :352-354: float operands, cos/sin evaluated in double (the object file imports sincos) and the product rounded to float */
      const float arg = ref_pose[5] + (s1 / 2 * time_diff);
      const float r0 = (float)(ref_pose[0] + s0 * time_diff * cos(arg));
      const float r1 = (float)(ref_pose[1] + s0 * time_diff * sin(arg));
      const float r5 = ref_pose[5] + s1 * time_diff;
      ref_pose[0] = r0;
      ref_pose[1] = r1;
      ref_pose[5] = r5;
    }
    const float d = s0 * time_diff, d2 = d * d, th = s1 * time_diff, t2 = th * th;
    sigma[0] = a[0] * d2 + a[1] * t2;
    sigma[1] = a[2] * d2 + a[3] * t2;
    sigma[2] = a[4] * d2 + a[5] * t2;
    sigma[3] = a[6] * d2 + a[7] * t2;
    sigma[4] = a[8] * d2 + a[9] * t2;
    sigma[5] = a[10] * d2 + a[11] * t2;
    if (variant == 2)
    {
      mean[0] = d;
      mean[5] = th;
    }
    return 0;
  }
  if (variant == 3)
  {
    /* :436-456 */
    const float d = (float)in[0] * time_diff, d2 = d * d;
    const float roll = (float)in[1], pitch = (float)in[2], th = (float)in[3];
    const float r2 = roll * roll, p2 = pitch * pitch, t2 = th * th;
    sigma[0] = a[0] * d2 + a[1] * t2;
    sigma[1] = a[2] * d2 + a[3] * t2;
    sigma[2] = a[4] * d2 + a[5] * t2;
    sigma[3] = a[7] * r2;
    sigma[4] = a[9] * p2;
    sigma[5] = a[11] * t2;
    mean[3] = roll;
    mean[4] = pitch;
    mean[5] = th;
    return 0;
  }
  return 1;
}

/* rotation part of R = Rz(c) Ry(b) Rx(a) from float sines / cosines, written like particle_cloud.cpp:525-539 */
static void rot_rows(float sa, float ca, float sb, float cb, float sc, float cc, float m[12])
{
  m[0] = cb * cc;
  m[4] = cb * sc;
  m[8] = -sb;
  m[1] = sa * sb * cc - ca * sc;
  m[5] = sa * sb * sc + ca * cc;
  m[9] = sa * cb;
  m[2] = ca * sb * cc + sa * sc;
  m[6] = ca * sb * sc - sa * cc;
  m[10] = ca * cb;
}

void oracle_motion_apply(float* P, uint64_t n, const double* draws)
{
  for (uint64_t i = 0; i < n; ++i)
  {
    float* p = P + 7 * i;
    const double dx = draws[6 * i], dy = draws[6 * i + 1], dz = draws[6 * i + 2];
    const double roll = draws[6 * i + 3], pitch = draws[6 * i + 4], yaw = draws[6 * i + 5];
    float o[12], q[12], tf[12];
    /* :508-513 FLOAT_T s = sin(double) */
    rot_rows((float)sin(roll), (float)cos(roll), (float)sin(pitch), (float)cos(pitch), (float)sin(yaw), (float)cos(yaw), o);
    o[3] = (float)dx;
    o[7] = (float)dy;
    o[11] = (float)dz;
    /* :543-553 the particle's own angles are floats; sin/cos still evaluate in double */
    rot_rows((float)sin((double)p[3]), (float)cos((double)p[3]), (float)sin((double)p[4]), (float)cos((double)p[4]),
             (float)sin((double)p[5]), (float)cos((double)p[5]), q);
    q[3] = 0;
    q[7] = 0;
    q[11] = 0;
    /* :575-588 */
    for (int r = 0; r < 3; ++r)
    {
      tf[4 * r + 0] = q[4 * r] * o[0] + q[4 * r + 1] * o[4] + q[4 * r + 2] * o[8];
      tf[4 * r + 1] = q[4 * r] * o[1] + q[4 * r + 1] * o[5] + q[4 * r + 2] * o[9];
      tf[4 * r + 2] = q[4 * r] * o[2] + q[4 * r + 1] * o[6] + q[4 * r + 2] * o[10];
      tf[4 * r + 3] = q[4 * r] * o[3] + q[4 * r + 1] * o[7] + q[4 * r + 2] * o[11] + q[4 * r + 3];
    }
    /* getAngleFromMat, src/util/util.cpp:86-111 */
    float t_roll, t_pitch, t_yaw;
    if (fabs((double)tf[8]) >= 1)
    {
      t_yaw = 0;
      const double delta = atan2((double)tf[9], (double)tf[10]);
      t_pitch = (float)(tf[8] < 0 ? M_PI / 2.0 : -M_PI / 2.0);
      t_roll = (float)delta;
    }
    else
    {
      t_pitch = (float)(-asin((double)tf[8]));
      t_roll = (float)atan2(tf[9] / cos((double)t_pitch), tf[10] / cos((double)t_pitch));
      t_yaw = (float)atan2(tf[4] / cos((double)t_pitch), tf[0] / cos((double)t_pitch));
    }
    /* :600-616 */
    p[0] = p[0] + tf[3];
    p[1] = p[1] + tf[7];
    p[2] = p[2] + tf[11];
    p[3] = t_roll;
    p[4] = t_pitch;
    p[5] = t_yaw;
  }
}
