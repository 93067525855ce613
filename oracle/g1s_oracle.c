/*
 * g1s_oracle.c — CPU restatement of grav1synth's `diff` hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (grav1synth_b200/, the
 * C-ABI library) links, loads or calls this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY: PINNED TO THE UPSTREAM BINARY, NOT TO THE REFERENCE BINARY.  The arithmetic of this path
 * lives in the third-party crate `av1-grain 0.4.2` (/root/reference/Cargo.toml:15, Cargo.lock:92-104),
 * whose source is not vendored in /root/reference and is not on this box; the reference holds no golden
 * vector or test for `diff` (tests/sanity_tests.rs never runs it) and cannot be built here (no Rust).
 * That crate module is a port of libaom aom_dsp/noise_model.c + mathutils.h::linsolve, and a compiled
 * libaom 3.13.1 ships in this image: oracle/aom_pin.py calls its aom_flat_block_finder_* /
 * aom_noise_model_* entry points directly.  In G1SO_GRAM_REF_ORDER + G1SO_EXP_LIBM mode this file is
 * BIT-IDENTICAL to that binary on every case of tests/aom_cases.py plus a 60-case random sweep: flat-block
 * maps, statuses, observation counts, the f64 bit patterns of the AR solution / AR gain / strength solution
 * of the latest and combined state after every frame, and every integer of every emitted segment
 * (tests/test_aom_pin.py, goldens under tests/golden/aom/).  One restated detail turned out wrong and was
 * corrected by this pin: fit_piecewise moves only the points, residual[] keeps its slots.
 * Still "parity unpinned" (crate-only behaviour that libaom cannot witness): the truncating `>> (bd-8)`
 * reduction of high-bit-depth input, timestamps, the fixed seed, swallowing of NoiseStatus::Error, and NaN
 * handling in get_grain_parameters when a chroma strength is exactly zero (Rust's f64::max / `as` casts are
 * followed, C's macros differ: tests/test_aom_pin.py::test_nan_correlation_follows_rust_semantics_not_c).
 * What follows restates the algorithm anchored on the reference's own call sites:
 *   DiffGenerator::new        src/main.rs:420-427
 *   DiffGenerator::diff_frame src/main.rs:442, 462, 482, 502
 *   DiffGenerator::finish     src/main.rs:524
 *   result conversion         src/parser/grain.rs:108-133, src/main.rs:705-713
 *   grain table text          src/main.rs:525-530, 631-696 (+ tests/example-table.tbl)
 * Each function names the crate/libaom routine it follows.
 *
 * Two accumulation modes for the AR normal equations:
 *   G1SO_GRAM_REF_ORDER  per-term f64 `A[i][j] += buf[i]*buf[j] / 255^2` in the
 *                        reference's loop order (what the crate does);
 *   G1SO_GRAM_EXACT_INT  exact int64 sums, one division at the end (what the CUDA
 *                        engine computes).  Tests assert both modes give identical
 *                        grain tables on the whole corpus; the modes differ by ~1e-14
 *                        in the solutions, which can only show where fit_piecewise
 *                        meets a structural tie.
 * Two exp() flavours for the flat-block sigmoid score:
 *   G1SO_EXP_LIBM        libm exp (what Rust's f64::exp calls);
 *   G1SO_EXP_FIXED       a fixed +,*,fma sequence that CPU and GPU evaluate
 *                        bit-identically (< 1 ulp from libm).
 *
 * Build: gcc -O3 -ffp-contract=off (no fast-math, no contraction: the crate's
 * release profile has neither, /root/reference/Cargo.toml:49-51).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/g1s.h"

#define BLOCK_SIZE 32
#define BLOCK_SIZE_SQUARED (BLOCK_SIZE * BLOCK_SIZE)
#define LOW_POLY_NUM_PARAMS 3
#define NOISE_MODEL_LAG 3
#define BLOCK_NORMALIZATION 255.0
#define NUM_BINS 20
#define TINY_NEAR_ZERO 1.0E-16
#define DEFAULT_GRAIN_SEED 10956

enum { G1SO_GRAM_REF_ORDER = 0, G1SO_GRAM_EXACT_INT = 1 };
enum { G1SO_EXP_LIBM = 0, G1SO_EXP_FIXED = 1 };
enum { ST_OK = 0, ST_DIFFERENT = 1, ST_ERROR = 2 };

/* ------------------------------------------------------------------ linsolve */

/* libaom mathutils.h::linsolve == av1-grain diff/solver/util.rs::linsolve:
 * Gaussian elimination, larger |pivot| bubbled up by adjacent-row swaps. */
static int linsolve(int n, double *A, int stride, double *b, double *x) {
  for (int k = 0; k < n - 1; k++) {
    for (int i = n - 1; i > k; i--) {
      if (fabs(A[(i - 1) * stride + k]) < fabs(A[i * stride + k])) {
        for (int j = 0; j < n; j++) {
          const double c = A[i * stride + j];
          A[i * stride + j] = A[(i - 1) * stride + j];
          A[(i - 1) * stride + j] = c;
        }
        const double c = b[i];
        b[i] = b[i - 1];
        b[i - 1] = c;
      }
    }
    for (int i = k; i < n - 1; i++) {
      if (fabs(A[k * stride + k]) < TINY_NEAR_ZERO) return 0;
      const double c = A[(i + 1) * stride + k] / A[k * stride + k];
      for (int j = 0; j < n; j++) A[(i + 1) * stride + j] -= c * A[k * stride + j];
      b[i + 1] -= c * b[k];
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    if (fabs(A[i * stride + i]) < TINY_NEAR_ZERO) return 0;
    double c = 0;
    for (int j = i + 1; j <= n - 1; j++) c += A[i * stride + j] * x[j];
    x[i] = (b[i] - c) / A[i * stride + i];
  }
  return 1;
}

/* util.rs::multiply_mat (libaom noise_model.c multiply_mat) */
static void multiply_mat(const double *m1, const double *m2, double *res, int m1_rows,
                         int inner_dim, int m2_cols) {
  for (int row = 0; row < m1_rows; ++row) {
    for (int col = 0; col < m2_cols; ++col) {
      double sum = 0;
      for (int inner = 0; inner < inner_dim; ++inner)
        sum += m1[row * inner_dim + inner] * m2[inner * m2_cols + col];
      *(res++) = sum;
    }
  }
}

/* ------------------------------------------------------------ EquationSystem */

typedef struct {
  int n;
  double *A, *b, *x;
} eqsys;

static void eq_init(eqsys *e, int n) {
  e->n = n;
  e->A = (double *)calloc((size_t)n * n, sizeof(double));
  e->b = (double *)calloc((size_t)n, sizeof(double));
  e->x = (double *)calloc((size_t)n, sizeof(double));
}
static void eq_free(eqsys *e) {
  free(e->A);
  free(e->b);
  free(e->x);
}
static void eq_clear(eqsys *e) {
  memset(e->A, 0, sizeof(double) * e->n * e->n);
  memset(e->b, 0, sizeof(double) * e->n);
  memset(e->x, 0, sizeof(double) * e->n);
}
static void eq_copy(eqsys *dst, const eqsys *src) {
  memcpy(dst->A, src->A, sizeof(double) * src->n * src->n);
  memcpy(dst->b, src->b, sizeof(double) * src->n);
  memcpy(dst->x, src->x, sizeof(double) * src->n);
}
/* equation_system_add: A and b only */
static void eq_add(eqsys *dst, const eqsys *src) {
  const int n = dst->n;
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) dst->A[i * n + j] += src->A[i * n + j];
    dst->b[i] += src->b[i];
  }
}
/* EquationSystem::solve — works on copies of A and b */
static int eq_solve(eqsys *e) {
  const int n = e->n;
  double *A = (double *)malloc(sizeof(double) * n * n);
  double *b = (double *)malloc(sizeof(double) * n);
  memcpy(A, e->A, sizeof(double) * n * n);
  memcpy(b, e->b, sizeof(double) * n);
  const int ret = linsolve(n, A, n, b, e->x);
  free(A);
  free(b);
  return ret;
}

/* ------------------------------------------------------- NoiseStrengthSolver */

typedef struct {
  eqsys eqns;
  double min_intensity, max_intensity;
  int num_bins;
  int num_equations;
  double total;
} strength_solver;

static void ss_init(strength_solver *s) {
  eq_init(&s->eqns, NUM_BINS);
  s->num_bins = NUM_BINS;
  s->min_intensity = 0;
  s->max_intensity = 255; /* (1 << 8) - 1: frames are reduced to 8 bit first */
  s->num_equations = 0;
  s->total = 0;
}
static void ss_clear(strength_solver *s) {
  eq_clear(&s->eqns);
  s->num_equations = 0;
  s->total = 0;
}
static double fclamp(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

static double ss_bin_index(const strength_solver *s, double value) {
  const double val = fclamp(value, s->min_intensity, s->max_intensity);
  const double range = s->max_intensity - s->min_intensity;
  return (s->num_bins - 1) * (val - s->min_intensity) / range;
}
static double ss_center(const strength_solver *s, int i) {
  const double range = s->max_intensity - s->min_intensity;
  const int n = s->num_bins;
  return ((double)i) / (n - 1) * range + s->min_intensity;
}
static double ss_get_value(const strength_solver *s, double x) {
  const double bin = ss_bin_index(s, x);
  const int i0 = (int)floor(bin);
  const int i1 = (s->num_bins - 1 < i0 + 1) ? s->num_bins - 1 : i0 + 1;
  const double a = bin - i0;
  return (1.0 - a) * s->eqns.x[i0] + a * s->eqns.x[i1];
}
static void ss_add_measurement(strength_solver *s, double block_mean, double noise_std) {
  const double bin = ss_bin_index(s, block_mean);
  const int i0 = (int)floor(bin);
  const int i1 = (s->num_bins - 1 < i0 + 1) ? s->num_bins - 1 : i0 + 1;
  const double a = bin - i0;
  const int n = s->num_bins;
  s->eqns.A[i0 * n + i0] += (1.0 - a) * (1.0 - a);
  s->eqns.A[i1 * n + i0] += a * (1.0 - a);
  s->eqns.A[i1 * n + i1] += a * a;
  s->eqns.A[i0 * n + i1] += a * (1.0 - a);
  s->eqns.b[i0] += (1.0 - a) * noise_std;
  s->eqns.b[i1] += a * noise_std;
  s->total += noise_std;
  s->num_equations++;
}
/* NoiseStrengthSolver::solve: smoothness + ridge on a copy of A; b is modified in
 * place (as in libaom aom_noise_strength_solver_solve). */
static int ss_solve(strength_solver *s) {
  const int n = s->num_bins;
  const double kAlpha = 2.0 * (double)(s->num_equations) / n;
  double *old_A = s->eqns.A;
  double *A = (double *)malloc(sizeof(double) * n * n);
  memcpy(A, old_A, sizeof(double) * n * n);
  for (int i = 0; i < n; ++i) {
    const int i_lo = i - 1 > 0 ? i - 1 : 0;
    const int i_hi = i + 1 < n - 1 ? i + 1 : n - 1;
    A[i * n + i_lo] -= kAlpha;
    A[i * n + i] += 2 * kAlpha;
    A[i * n + i_hi] -= kAlpha;
  }
  const double mean = s->total / s->num_equations;
  for (int i = 0; i < n; ++i) {
    A[i * n + i] += 1.0 / 8192.;
    s->eqns.b[i] += mean / 8192.;
  }
  s->eqns.A = A;
  const int result = eq_solve(&s->eqns);
  s->eqns.A = old_A;
  free(A);
  return result;
}
static void ss_add(strength_solver *dst, const strength_solver *src) {
  eq_add(&dst->eqns, &src->eqns);
  dst->num_equations += src->num_equations;
  dst->total += src->total;
}

typedef struct {
  double pts[NUM_BINS][2];
  int n;
} strength_lut;

static void update_piecewise_linear_residual(const strength_solver *s, const strength_lut *lut,
                                             double *residual, int start, int end) {
  const double dx = 255. / s->num_bins;
  const int lo = start > 1 ? start : 1;
  const int hi = end < lut->n - 1 ? end : lut->n - 1;
  for (int i = lo; i < hi; ++i) {
    int lower = (int)floor(ss_bin_index(s, lut->pts[i - 1][0]));
    if (lower < 0) lower = 0;
    int upper = (int)ceil(ss_bin_index(s, lut->pts[i + 1][0]));
    if (upper > s->num_bins - 1) upper = s->num_bins - 1;
    double r = 0;
    for (int j = lower; j <= upper; ++j) {
      const double x = ss_center(s, j);
      if (x < lut->pts[i - 1][0]) continue;
      if (x >= lut->pts[i + 1][0]) continue;
      const double y = s->eqns.x[j];
      const double a = (x - lut->pts[i - 1][0]) / (lut->pts[i + 1][0] - lut->pts[i - 1][0]);
      const double estimate_y = lut->pts[i - 1][1] * (1.0 - a) + lut->pts[i + 1][1] * a;
      r += fabs(y - estimate_y);
    }
    residual[i] = r * dx;
  }
}

/* NoiseStrengthSolver::fit_piecewise */
static void ss_fit_piecewise(const strength_solver *s, int max_output_points, strength_lut *lut) {
  const double kTolerance = s->max_intensity * 0.00625 / 255.0;
  lut->n = s->num_bins;
  for (int i = 0; i < s->num_bins; ++i) {
    lut->pts[i][0] = ss_center(s, i);
    lut->pts[i][1] = s->eqns.x[i];
  }
  double residual[NUM_BINS];
  memset(residual, 0, sizeof(residual));
  update_piecewise_linear_residual(s, lut, residual, 0, s->num_bins);
  while (lut->n > 2) {
    int min_index = 1;
    for (int j = 1; j < lut->n - 1; ++j)
      if (residual[j] < residual[min_index]) min_index = j;
    const double dx = lut->pts[min_index + 1][0] - lut->pts[min_index - 1][0];
    const double avg_residual = residual[min_index] / dx;
    if (lut->n <= max_output_points && avg_residual > kTolerance) break;
    const int num_remaining = lut->n - min_index - 1;
    memmove(lut->pts + min_index, lut->pts + min_index + 1, sizeof(lut->pts[0]) * num_remaining);
    /* libaom (and its port) shift only the points: residual[] keeps its old slots (checked against the
     * libaom 3.13.1 binary, oracle/aom_pin.py) */
    lut->n--;
    update_piecewise_linear_residual(s, lut, residual, min_index - 1, min_index + 1);
  }
}

/* ---------------------------------------------------------------- NoiseState */

typedef struct {
  eqsys eqns;
  double ar_gain;
  int64_t num_observations;
  strength_solver strength;
} noise_state;

static void ns_init(noise_state *s, int n) {
  eq_init(&s->eqns, n);
  s->ar_gain = 1.0;
  s->num_observations = 0;
  ss_init(&s->strength);
}
static void ns_free(noise_state *s) {
  eq_free(&s->eqns);
  eq_free(&s->strength.eqns);
}

/* NoiseModel::ar_equation_system_solve */
static int ar_equation_system_solve(noise_state *state, int is_chroma) {
  const int ret = eq_solve(&state->eqns);
  state->ar_gain = 1.0;
  if (!ret) return ret;
  double var = 0;
  const int n = state->eqns.n;
  for (int i = 0; i < n - is_chroma; ++i)
    var += state->eqns.A[i * n + i] / (double)state->num_observations;
  var /= (n - is_chroma);
  double sum_covar = 0;
  for (int i = 0; i < n - is_chroma; ++i) {
    double bi = state->eqns.b[i];
    if (is_chroma) bi -= state->eqns.A[i * n + (n - 1)] * state->eqns.x[n - 1];
    sum_covar += (bi * state->eqns.x[i]) / (double)state->num_observations;
  }
  const double noise_var = fmax(var - sum_covar, 1e-6);
  state->ar_gain = fmax(1, sqrt(fmax(var / noise_var, 1e-6)));
  return ret;
}

static void set_chroma_coefficient_fallback_soln(eqsys *eqns) {
  const double kTolerance = 1e-6;
  const int last = eqns->n - 1;
  memset(eqns->x, 0, sizeof(double) * eqns->n);
  if (fabs(eqns->A[last * eqns->n + last]) > kTolerance)
    eqns->x[last] = eqns->b[last] / eqns->A[last * eqns->n + last];
}

/* --------------------------------------------------------------- the oracle */

typedef struct g1s_oracle {
  int64_t fps_num, fps_den;
  int src_bd, den_bd;
  int gram_mode, exp_mode;
  int64_t frame_count;
  uint64_t prev_timestamp;
  /* FlatBlockFinder */
  double *fbA;         /* 1024 x 3 */
  double AtA_inv[9];
  /* NoiseModel */
  int n; /* 24 */
  int coords[24][2];
  noise_state latest[3], combined[3];
  /* grain table */
  g1s_segment *segs;
  size_t nsegs, capsegs;
  /* introspection of the most recent frame (tests) */
  uint8_t *last_flat;
  int last_nb_w, last_nb_h, last_num_flat;
  float *last_scores;
  double *last_feat; /* nb x 5: gxx gxy gyy mean var (normalised) */
  int64_t last_gram[3][26 * 26];
  int64_t last_gram_nobs[3];
  int last_status;
  char err[256];
} g1s_oracle;

/* diff/solver.rs FlatBlockFinder::new */
static void fbf_init(g1s_oracle *o) {
  eqsys eqns;
  eq_init(&eqns, LOW_POLY_NUM_PARAMS);
  o->fbA = (double *)calloc((size_t)LOW_POLY_NUM_PARAMS * BLOCK_SIZE_SQUARED, sizeof(double));
  const double bs_half = BLOCK_SIZE / 2;
  for (int y = 0; y < BLOCK_SIZE; ++y) {
    const double yd = ((double)y - bs_half) / bs_half;
    for (int x = 0; x < BLOCK_SIZE; ++x) {
      const double xd = ((double)x - bs_half) / bs_half;
      const double c[3] = {yd, xd, 1.0};
      const int row = y * BLOCK_SIZE + x;
      o->fbA[LOW_POLY_NUM_PARAMS * row + 0] = yd;
      o->fbA[LOW_POLY_NUM_PARAMS * row + 1] = xd;
      o->fbA[LOW_POLY_NUM_PARAMS * row + 2] = 1.0;
      for (int i = 0; i < LOW_POLY_NUM_PARAMS; ++i)
        for (int j = 0; j < LOW_POLY_NUM_PARAMS; ++j) eqns.A[LOW_POLY_NUM_PARAMS * i + j] += c[i] * c[j];
    }
  }
  for (int i = 0; i < LOW_POLY_NUM_PARAMS; ++i) {
    memset(eqns.b, 0, sizeof(double) * LOW_POLY_NUM_PARAMS);
    eqns.b[i] = 1.0;
    eq_solve(&eqns);
    for (int j = 0; j < LOW_POLY_NUM_PARAMS; ++j) o->AtA_inv[j * LOW_POLY_NUM_PARAMS + i] = eqns.x[j];
  }
  eq_free(&eqns);
}

/* Fixed-sequence exp: k = rint(x/ln2), r = x - k*ln2 (two-part), degree-13 Horner
 * in fma, scale by 2^k through the exponent field.  Valid for |x| < 700.  The
 * CUDA kernel evaluates the same sequence (csrc/g1s_kernels.cu: g1s_exp_fixed). */
static double g1s_exp_fixed(double x) {
  const double inv_ln2 = 1.4426950408889634074;
  const double ln2_hi = 6.93147180369123816490e-01;
  const double ln2_lo = 1.90821492927058770002e-10;
  const double shift = 6755399441055744.0; /* 1.5 * 2^52 */
  const double kd = (x * inv_ln2 + shift) - shift;
  const int k = (int)kd;
  double r = fma(-kd, ln2_hi, x);
  r = fma(-kd, ln2_lo, r);
  double p = 1.0 / 6227020800.0;
  p = fma(p, r, 1.0 / 479001600.0);
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  union {
    double d;
    uint64_t u;
  } sc;
  sc.u = (uint64_t)(1023 + k) << 52;
  return p * sc.d;
}

/* FlatBlockFinder::extract_block (libaom aom_flat_block_finder_extract_block) */
static void fbf_extract_block(const g1s_oracle *o, const uint8_t *data, int w, int h, int stride,
                              int offsx, int offsy, double *plane, double *block) {
  double plane_coords[LOW_POLY_NUM_PARAMS];
  double AtA_inv_b[LOW_POLY_NUM_PARAMS];
  for (int yi = 0; yi < BLOCK_SIZE; ++yi) {
    int y = offsy + yi;
    if (y > h - 1) y = h - 1;
    for (int xi = 0; xi < BLOCK_SIZE; ++xi) {
      int x = offsx + xi;
      if (x > w - 1) x = w - 1;
      block[yi * BLOCK_SIZE + xi] = ((double)data[y * stride + x]) / BLOCK_NORMALIZATION;
    }
  }
  multiply_mat(block, o->fbA, AtA_inv_b, 1, BLOCK_SIZE_SQUARED, LOW_POLY_NUM_PARAMS);
  multiply_mat(o->AtA_inv, AtA_inv_b, plane_coords, LOW_POLY_NUM_PARAMS, LOW_POLY_NUM_PARAMS, 1);
  multiply_mat(o->fbA, plane_coords, plane, BLOCK_SIZE_SQUARED, LOW_POLY_NUM_PARAMS, 1);
  for (int i = 0; i < BLOCK_SIZE_SQUARED; ++i) block[i] -= plane[i];
}

typedef struct {
  float score;
  int index;
} index_and_score;

static int cmp_score(const void *a, const void *b) {
  const float x = ((const index_and_score *)a)->score, y = ((const index_and_score *)b)->score;
  return (x > y) - (x < y);
}

/* FlatBlockFinder::run (libaom aom_flat_block_finder_run) */
static int fbf_run(g1s_oracle *o, const uint8_t *data, int w, int h, int stride, uint8_t *flat_blocks) {
  const double kTraceThreshold = 0.15 / BLOCK_SIZE_SQUARED;
  const double kRatioThreshold = 1.25;
  const double kNormThreshold = 0.08 / BLOCK_SIZE_SQUARED;
  const double kVarThreshold = 0.005 / BLOCK_SIZE_SQUARED;
  const int num_blocks_w = (w + BLOCK_SIZE - 1) / BLOCK_SIZE;
  const int num_blocks_h = (h + BLOCK_SIZE - 1) / BLOCK_SIZE;
  const int num_blocks = num_blocks_w * num_blocks_h;
  int num_flat = 0;
  double *plane = (double *)malloc(sizeof(double) * BLOCK_SIZE_SQUARED);
  double *block = (double *)malloc(sizeof(double) * BLOCK_SIZE_SQUARED);
  index_and_score *scores = (index_and_score *)malloc(sizeof(index_and_score) * num_blocks);

  for (int by = 0; by < num_blocks_h; ++by) {
    for (int bx = 0; bx < num_blocks_w; ++bx) {
      double Gxx = 0, Gxy = 0, Gyy = 0, var = 0, mean = 0;
      fbf_extract_block(o, data, w, h, stride, bx * BLOCK_SIZE, by * BLOCK_SIZE, plane, block);
      for (int yi = 1; yi < BLOCK_SIZE - 1; ++yi) {
        for (int xi = 1; xi < BLOCK_SIZE - 1; ++xi) {
          const double gx = (block[yi * BLOCK_SIZE + xi + 1] - block[yi * BLOCK_SIZE + xi - 1]) / 2;
          const double gy = (block[yi * BLOCK_SIZE + xi + BLOCK_SIZE] - block[yi * BLOCK_SIZE + xi - BLOCK_SIZE]) / 2;
          Gxx += gx * gx;
          Gxy += gx * gy;
          Gyy += gy * gy;
          const double v = block[yi * BLOCK_SIZE + xi];
          mean += v;
          var += v * v;
        }
      }
      const double nf = (BLOCK_SIZE - 2) * (BLOCK_SIZE - 2);
      mean /= nf;
      Gxx /= nf;
      Gxy /= nf;
      Gyy /= nf;
      var = var / nf - mean * mean;
      {
        const double trace = Gxx + Gyy;
        const double det = Gxx * Gyy - Gxy * Gxy;
        const double e_sub = sqrt(fmax(trace * trace - 4 * det, 0.));
        const double e1 = (trace + e_sub) / 2.;
        const double e2 = (trace - e_sub) / 2.;
        const double norm = e1;
        const double ratio = (e1 / fmax(e2, 1e-6));
        const int is_flat = (trace < kTraceThreshold) && (ratio < kRatioThreshold) &&
                            (norm < kNormThreshold) && (var > kVarThreshold);
        /* crate: nested mul_add, innermost first */
        double sum_weights =
            fma(-6682.0, var, fma(-0.2056, ratio, fma(13087.0, trace, fma(-12434.0, norm, 2.5694))));
        sum_weights = fclamp(sum_weights, -25.0, 100.0);
        const double e = (o->exp_mode == G1SO_EXP_FIXED) ? g1s_exp_fixed(-sum_weights) : exp(-sum_weights);
        const float score = (float)(1.0 / (1.0 + e));
        const int idx = by * num_blocks_w + bx;
        flat_blocks[idx] = is_flat ? 255 : 0;
        scores[idx].score = var > kVarThreshold ? score : 0;
        scores[idx].index = idx;
        num_flat += is_flat;
        if (o->last_feat) {
          double *f = o->last_feat + 5 * idx;
          f[0] = Gxx; f[1] = Gxy; f[2] = Gyy; f[3] = mean; f[4] = var;
        }
        if (o->last_scores) o->last_scores[idx] = scores[idx].score;
      }
    }
  }
  qsort(scores, num_blocks, sizeof(*scores), cmp_score);
  const int top_nth_percentile = num_blocks * 90 / 100;
  const float score_threshold = scores[top_nth_percentile].score;
  for (int i = 0; i < num_blocks; ++i) {
    if (scores[i].score >= score_threshold) {
      num_flat += flat_blocks[scores[i].index] == 0;
      flat_blocks[scores[i].index] |= 1;
    }
  }
  free(block);
  free(plane);
  free(scores);
  return num_flat;
}

/* util.rs::extract_ar_row (libaom EXTRACT_AR_ROW), 8-bit planes */
static double extract_ar_row(int (*coords)[2], int num_coords, const uint8_t *data,
                             const uint8_t *denoised, int stride, const int sub_log2[2],
                             const uint8_t *alt_data, const uint8_t *alt_denoised, int alt_stride,
                             int x, int y, double *buffer) {
  for (int i = 0; i < num_coords; ++i) {
    const int x_i = x + coords[i][0], y_i = y + coords[i][1];
    buffer[i] = (double)data[y_i * stride + x_i] - denoised[y_i * stride + x_i];
  }
  const double val = (double)data[y * stride + x] - denoised[y * stride + x];
  if (alt_data && alt_denoised) {
    double avg_data = 0, avg_denoised = 0;
    int num_samples = 0;
    for (int dy_i = 0; dy_i < (1 << sub_log2[1]); dy_i++) {
      const int y_up = (y << sub_log2[1]) + dy_i;
      for (int dx_i = 0; dx_i < (1 << sub_log2[0]); dx_i++) {
        const int x_up = (x << sub_log2[0]) + dx_i;
        avg_data += alt_data[y_up * alt_stride + x_up];
        avg_denoised += alt_denoised[y_up * alt_stride + x_up];
        num_samples++;
      }
    }
    buffer[num_coords] = (avg_data - avg_denoised) / num_samples;
  }
  return val;
}

/* NoiseModel::add_block_observations (libaom add_block_observations) */
static void add_block_observations(g1s_oracle *o, int c, const uint8_t *data, const uint8_t *denoised,
                                   int w, int h, int stride, const int sub_log2[2],
                                   const uint8_t *alt_data, const uint8_t *alt_denoised, int alt_stride,
                                   const uint8_t *flat_blocks, int num_blocks_w, int num_blocks_h) {
  const int lag = NOISE_MODEL_LAG;
  const int num_coords = o->n;
  noise_state *st = &o->latest[c];
  double *A = st->eqns.A;
  double *b = st->eqns.b;
  double buffer[26];
  const int n = st->eqns.n;
  const int block_w = BLOCK_SIZE >> sub_log2[0];
  const int block_h = BLOCK_SIZE >> sub_log2[1];
  const double norm2 = BLOCK_NORMALIZATION * BLOCK_NORMALIZATION;
  const int exact = o->gram_mode == G1SO_GRAM_EXACT_INT;
  const int nss = 1 << (sub_log2[0] + sub_log2[1]); /* luma samples behind one chroma sample */
  int64_t *G = o->last_gram[c]; /* [26][26]: rows/cols 0..23 taps, 24 = luma tap (x nss), 25 = val */
  memset(G, 0, sizeof(int64_t) * 26 * 26);
  int64_t nobs = 0;

  for (int by = 0; by < num_blocks_h; ++by) {
    const int y_o = by * block_h;
    for (int bx = 0; bx < num_blocks_w; ++bx) {
      const int x_o = bx * block_w;
      if (!flat_blocks[by * num_blocks_w + bx]) continue;
      const int y_start = (by > 0 && flat_blocks[(by - 1) * num_blocks_w + bx]) ? 0 : lag;
      const int x_start = (bx > 0 && flat_blocks[by * num_blocks_w + bx - 1]) ? 0 : lag;
      int y_end = (h >> sub_log2[1]) - by * block_h;
      if (y_end > block_h) y_end = block_h;
      int x_end = (w >> sub_log2[0]) - bx * block_w - lag;
      const int x_lim = (bx + 1 < num_blocks_w && flat_blocks[by * num_blocks_w + bx + 1]) ? block_w : block_w - lag;
      if (x_end > x_lim) x_end = x_lim;
      for (int y = y_start; y < y_end; ++y) {
        for (int x = x_start; x < x_end; ++x) {
          const double val = extract_ar_row(o->coords, num_coords, data, denoised, stride, sub_log2, alt_data,
                                            alt_denoised, alt_stride, x + x_o, y + y_o, buffer);
          if (exact) {
            int64_t ib[26];
            for (int i = 0; i < num_coords; ++i) ib[i] = (int64_t)buffer[i];
            ib[24] = alt_data ? (int64_t)(buffer[24] * nss) : 0;
            ib[25] = (int64_t)val;
            for (int i = 0; i < 26; ++i)
              for (int j = 0; j < 26; ++j) G[i * 26 + j] += ib[i] * ib[j];
          } else {
            for (int i = 0; i < n; ++i) {
              for (int j = 0; j < n; ++j) A[i * n + j] += (buffer[i] * buffer[j]) / norm2;
              b[i] += (buffer[i] * val) / norm2;
            }
          }
          nobs++;
        }
      }
    }
  }
  st->num_observations += nobs;
  o->last_gram_nobs[c] = nobs;
  if (exact) {
    /* one rounding per entry: exact integer / (power-of-two tap scale) / 255^2 */
    for (int i = 0; i < n; ++i) {
      const double si = (i == 24) ? (double)nss : 1.0;
      for (int j = 0; j < n; ++j) {
        const double sj = (j == 24) ? (double)nss : 1.0;
        A[i * n + j] += ((double)G[i * 26 + j] / (si * sj)) / norm2;
      }
      b[i] += ((double)G[i * 26 + 25] / si) / norm2;
    }
  }
}

static double get_block_mean(const uint8_t *data, int w, int h, int stride, int x_o, int y_o, int block_size) {
  const int max_h = h - y_o < block_size ? h - y_o : block_size;
  const int max_w = w - x_o < block_size ? w - x_o : block_size;
  double block_mean = 0;
  for (int y = 0; y < max_h; ++y)
    for (int x = 0; x < max_w; ++x) block_mean += data[(y_o + y) * stride + x_o + x];
  return block_mean / (max_w * max_h);
}

static double get_noise_var(const uint8_t *data, const uint8_t *denoised, int stride, int w, int h, int x_o,
                            int y_o, int block_size_x, int block_size_y) {
  const int max_h = h - y_o < block_size_y ? h - y_o : block_size_y;
  const int max_w = w - x_o < block_size_x ? w - x_o : block_size_x;
  double noise_var = 0, noise_mean = 0;
  for (int y = 0; y < max_h; ++y) {
    for (int x = 0; x < max_w; ++x) {
      const double noise = (double)data[(y_o + y) * stride + (x_o + x)] - denoised[(y_o + y) * stride + (x_o + x)];
      noise_mean += noise;
      noise_var += noise * noise;
    }
  }
  noise_mean /= (max_w * max_h);
  return noise_var / (max_w * max_h) - noise_mean * noise_mean;
}

/* NoiseModel::add_noise_std_observations */
static void add_noise_std_observations(g1s_oracle *o, int c, const double *coeffs, const uint8_t *data,
                                       const uint8_t *denoised, int w, int h, int stride, const int sub_log2[2],
                                       const uint8_t *alt_data, int alt_stride, const uint8_t *flat_blocks,
                                       int num_blocks_w, int num_blocks_h) {
  const int num_coords = o->n;
  strength_solver *solver = &o->latest[c].strength;
  const strength_solver *luma = &o->latest[0].strength;
  const double luma_gain = o->latest[0].ar_gain;
  const double noise_gain = o->latest[c].ar_gain;
  const int bsx = BLOCK_SIZE >> sub_log2[0], bsy = BLOCK_SIZE >> sub_log2[1];
  for (int by = 0; by < num_blocks_h; ++by) {
    const int y_o = by * bsy;
    for (int bx = 0; bx < num_blocks_w; ++bx) {
      const int x_o = bx * bsx;
      if (!flat_blocks[by * num_blocks_w + bx]) continue;
      int num_samples_h = (h >> sub_log2[1]) - by * bsy;
      if (num_samples_h > bsy) num_samples_h = bsy;
      int num_samples_w = (w >> sub_log2[0]) - bx * bsx;
      if (num_samples_w > bsx) num_samples_w = bsx;
      if (num_samples_w * num_samples_h > BLOCK_SIZE) {
        const double block_mean = get_block_mean(alt_data ? alt_data : data, w, h, alt_data ? alt_stride : stride,
                                                 x_o << sub_log2[0], y_o << sub_log2[1], BLOCK_SIZE);
        const double noise_var = get_noise_var(data, denoised, stride, w >> sub_log2[0], h >> sub_log2[1], x_o,
                                               y_o, bsx, bsy);
        const double luma_strength = c > 0 ? luma_gain * ss_get_value(luma, block_mean) : 0;
        const double corr = c > 0 ? coeffs[num_coords] : 0;
        const double t = corr * luma_strength;
        const double uncorr_std = sqrt(fmax(noise_var / 16, noise_var - t * t));
        const double adjusted_strength = uncorr_std / noise_gain;
        ss_add_measurement(solver, block_mean, adjusted_strength);
      }
    }
  }
}

static double normalized_cross_correlation(const double *a, const double *b, int n) {
  double c = 0, a_len = 0, b_len = 0;
  for (int i = 0; i < n; ++i) {
    a_len += a[i] * a[i];
    b_len += b[i] * b[i];
    c += a[i] * b[i];
  }
  return c / (sqrt(a_len) * sqrt(b_len));
}

/* NoiseModel::is_different (libaom is_noise_model_different), luma only */
static int is_noise_model_different(const g1s_oracle *o) {
  const double kCoeffThreshold = 0.9;
  const double kStrengthThreshold = 0.005; /* * (1 << (8 - 8)) */
  const double corr =
      normalized_cross_correlation(o->latest[0].eqns.x, o->combined[0].eqns.x, o->combined[0].eqns.n);
  if (corr < kCoeffThreshold) return 1;
  const double dx = 1.0 / o->latest[0].strength.num_bins;
  const eqsys *le = &o->latest[0].strength.eqns;
  const eqsys *ce = &o->combined[0].strength.eqns;
  double diff = 0, total_weight = 0;
  for (int j = 0; j < le->n; ++j) {
    double weight = 0;
    for (int i = 0; i < le->n; ++i) weight += le->A[i * le->n + j];
    weight = sqrt(weight);
    diff += weight * fabs(le->x[j] - ce->x[j]);
    total_weight += weight;
  }
  if (diff * dx / total_weight > kStrengthThreshold) return 1;
  return 0;
}

/* NoiseModel::update (libaom aom_noise_model_update) on 8-bit planes */
static int noise_model_update(g1s_oracle *o, const uint8_t *const data[3], const uint8_t *const denoised[3], int w,
                              int h, const int stride[3], const int chroma_sub_log2[2],
                              const uint8_t *flat_blocks) {
  const int num_blocks_w = (w + BLOCK_SIZE - 1) / BLOCK_SIZE;
  const int num_blocks_h = (h + BLOCK_SIZE - 1) / BLOCK_SIZE;
  int y_model_different = 0;
  int num_blocks = 0;
  for (int i = 0; i < 3; ++i) {
    eq_clear(&o->latest[i].eqns);
    o->latest[i].num_observations = 0;
    ss_clear(&o->latest[i].strength);
  }
  for (int i = 0; i < num_blocks_h * num_blocks_w; ++i)
    if (flat_blocks[i]) num_blocks++;
  if (num_blocks <= 1) {
    snprintf(o->err, sizeof(o->err), "Not enough flat blocks to update noise estimate");
    return ST_ERROR;
  }
  for (int channel = 0; channel < 3; ++channel) {
    const int no_subsampling[2] = {0, 0};
    const uint8_t *alt_data = channel > 0 ? data[0] : 0;
    const uint8_t *alt_denoised = channel > 0 ? denoised[0] : 0;
    const int *sub = channel > 0 ? chroma_sub_log2 : no_subsampling;
    const int is_chroma = channel != 0;
    if (!data[channel] || !denoised[channel]) break;
    add_block_observations(o, channel, data[channel], denoised[channel], w, h, stride[channel], sub, alt_data,
                           alt_denoised, stride[0], flat_blocks, num_blocks_w, num_blocks_h);
    if (!ar_equation_system_solve(&o->latest[channel], is_chroma)) {
      if (is_chroma) {
        set_chroma_coefficient_fallback_soln(&o->latest[channel].eqns);
      } else {
        snprintf(o->err, sizeof(o->err), "Solving latest noise equation system failed %d!", channel);
        return ST_ERROR;
      }
    }
    add_noise_std_observations(o, channel, o->latest[channel].eqns.x, data[channel], denoised[channel], w, h,
                               stride[channel], sub, alt_data, stride[0], flat_blocks, num_blocks_w, num_blocks_h);
    if (!ss_solve(&o->latest[channel].strength)) {
      snprintf(o->err, sizeof(o->err), "Solving latest noise strength failed!");
      return ST_ERROR;
    }
    if (channel == 0 && o->combined[channel].strength.num_equations > 0 && is_noise_model_different(o))
      y_model_different = 1;
    if (y_model_different) continue;

    o->combined[channel].num_observations += o->latest[channel].num_observations;
    eq_add(&o->combined[channel].eqns, &o->latest[channel].eqns);
    if (!ar_equation_system_solve(&o->combined[channel], is_chroma)) {
      if (is_chroma) {
        set_chroma_coefficient_fallback_soln(&o->combined[channel].eqns);
      } else {
        snprintf(o->err, sizeof(o->err), "Solving combined noise equation system failed %d!", channel);
        return ST_ERROR;
      }
    }
    ss_add(&o->combined[channel].strength, &o->latest[channel].strength);
    if (!ss_solve(&o->combined[channel].strength)) {
      snprintf(o->err, sizeof(o->err), "Solving combined noise strength failed!");
      return ST_ERROR;
    }
  }
  return y_model_different ? ST_DIFFERENT : ST_OK;
}

/* NoiseModel::save_latest (libaom aom_noise_model_save_latest) */
static void noise_model_save_latest(g1s_oracle *o) {
  for (int c = 0; c < 3; c++) {
    eq_copy(&o->combined[c].eqns, &o->latest[c].eqns);
    eq_copy(&o->combined[c].strength.eqns, &o->latest[c].strength.eqns);
    o->combined[c].strength.num_equations = o->latest[c].strength.num_equations;
    o->combined[c].num_observations = o->latest[c].num_observations;
    o->combined[c].ar_gain = o->latest[c].ar_gain;
  }
}

static int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* NoiseModel::get_grain_parameters (libaom aom_noise_model_get_grain_parameters) */
static void get_grain_parameters(const g1s_oracle *o, uint64_t start_ts, uint64_t end_ts, g1s_segment *seg) {
  memset(seg, 0, sizeof(*seg));
  seg->start_time = start_ts;
  seg->end_time = end_ts;
  seg->random_seed = start_ts == 0 ? DEFAULT_GRAIN_SEED : 0;
  seg->ar_coeff_lag = NOISE_MODEL_LAG;

  strength_lut sp[3];
  ss_fit_piecewise(&o->combined[0].strength, G1S_NUM_Y_POINTS, &sp[0]);
  ss_fit_piecewise(&o->combined[1].strength, G1S_NUM_UV_POINTS, &sp[1]);
  ss_fit_piecewise(&o->combined[2].strength, G1S_NUM_UV_POINTS, &sp[2]);

  const double strength_divisor = 1; /* 1 << (8 - 8) */
  double max_scaling_value = 1e-4;
  for (int c = 0; c < 3; ++c) {
    for (int i = 0; i < sp[c].n; ++i) {
      sp[c].pts[i][0] = fmin(255, sp[c].pts[i][0] / strength_divisor);
      sp[c].pts[i][1] = fmin(255, sp[c].pts[i][1] / strength_divisor);
      max_scaling_value = fmax(sp[c].pts[i][1], max_scaling_value);
    }
  }
  const int max_scaling_value_log2 = iclamp((int)floor(log2(max_scaling_value) + 1), 2, 5);
  seg->scaling_shift = (uint8_t)(5 + (8 - max_scaling_value_log2));
  const double scale_factor = (double)(1 << (8 - max_scaling_value_log2));
  seg->num_y_points = (uint8_t)sp[0].n;
  seg->num_cb_points = (uint8_t)sp[1].n;
  seg->num_cr_points = (uint8_t)sp[2].n;
  uint8_t(*dst[3])[2] = {seg->scaling_points_y, seg->scaling_points_cb, seg->scaling_points_cr};
  for (int c = 0; c < 3; c++) {
    for (int i = 0; i < sp[c].n; ++i) {
      dst[c][i][0] = (uint8_t)(int)(sp[c].pts[i][0] + 0.5);
      dst[c][i][1] = (uint8_t)iclamp((int)(scale_factor * sp[c].pts[i][1] + 0.5), 0, 255);
    }
  }

  const int n_coeff = o->combined[0].eqns.n;
  double max_coeff = 1e-4, min_coeff = -1e-4;
  double y_corr[2] = {0, 0};
  double avg_luma_strength = 0;
  for (int c = 0; c < 3; c++) {
    const eqsys *eqns = &o->combined[c].eqns;
    for (int i = 0; i < n_coeff; ++i) {
      max_coeff = fmax(max_coeff, eqns->x[i]);
      min_coeff = fmin(min_coeff, eqns->x[i]);
    }
    const strength_solver *solver = &o->combined[c].strength;
    double average_strength = 0, total_weight = 0;
    for (int i = 0; i < solver->eqns.n; ++i) {
      double w = 0;
      for (int j = 0; j < solver->eqns.n; ++j) w += solver->eqns.A[i * solver->eqns.n + j];
      w = sqrt(w);
      average_strength += solver->eqns.x[i] * w;
      total_weight += w;
    }
    if (total_weight == 0)
      average_strength = 1;
    else
      average_strength /= total_weight;
    if (c == 0) {
      avg_luma_strength = average_strength;
    } else {
      y_corr[c - 1] = avg_luma_strength * eqns->x[n_coeff] / average_strength;
      max_coeff = fmax(max_coeff, y_corr[c - 1]);
      min_coeff = fmin(min_coeff, y_corr[c - 1]);
    }
  }
  seg->ar_coeff_shift =
      (uint8_t)iclamp(7 - (int)fmax(1 + floor(log2(max_coeff)), ceil(log2(-min_coeff))), 6, 9);
  const double scale_ar_coeff = (double)(1 << seg->ar_coeff_shift);
  int8_t *ar[3] = {seg->ar_coeffs_y, seg->ar_coeffs_cb, seg->ar_coeffs_cr};
  for (int c = 0; c < 3; ++c) {
    const eqsys *eqns = &o->combined[c].eqns;
    for (int i = 0; i < n_coeff; ++i) ar[c][i] = (int8_t)iclamp((int)round(scale_ar_coeff * eqns->x[i]), -128, 127);
    if (c > 0) ar[c][n_coeff] = (int8_t)iclamp((int)round(scale_ar_coeff * y_corr[c - 1]), -128, 127);
  }
  seg->cb_mult = 128;
  seg->cb_luma_mult = 192;
  seg->cb_offset = 256;
  seg->cr_mult = 128;
  seg->cr_luma_mult = 192;
  seg->cr_offset = 256;
  seg->chroma_scaling_from_luma = 0;
  seg->grain_scale_shift = 0;
  seg->overlap_flag = 1;
}

static void push_segment(g1s_oracle *o, const g1s_segment *s) {
  if (o->nsegs == o->capsegs) {
    o->capsegs = o->capsegs ? o->capsegs * 2 : 8;
    o->segs = (g1s_segment *)realloc(o->segs, o->capsegs * sizeof(g1s_segment));
  }
  o->segs[o->nsegs++] = *s;
}

/* ------------------------------------------------------------- public (test) API */

/* DiffGenerator::new — call site src/main.rs:420-427 */
g1s_oracle *g1s_oracle_new(int64_t fps_num, int64_t fps_den, int src_bd, int den_bd, int gram_mode, int exp_mode) {
  g1s_oracle *o = (g1s_oracle *)calloc(1, sizeof(*o));
  o->fps_num = fps_num;
  o->fps_den = fps_den;
  o->src_bd = src_bd;
  o->den_bd = den_bd;
  o->gram_mode = gram_mode;
  o->exp_mode = exp_mode;
  fbf_init(o);
  o->n = 24;
  int i = 0;
  for (int y = -NOISE_MODEL_LAG; y <= 0; ++y) {
    const int max_x = y == 0 ? -1 : NOISE_MODEL_LAG;
    for (int x = -NOISE_MODEL_LAG; x <= max_x; ++x) {
      o->coords[i][0] = x;
      o->coords[i][1] = y;
      ++i;
    }
  }
  for (int c = 0; c < 3; ++c) {
    ns_init(&o->latest[c], c == 0 ? 24 : 25);
    ns_init(&o->combined[c], c == 0 ? 24 : 25);
  }
  return o;
}

void g1s_oracle_free(g1s_oracle *o) {
  if (!o) return;
  for (int c = 0; c < 3; ++c) {
    ns_free(&o->latest[c]);
    ns_free(&o->combined[c]);
  }
  free(o->fbA);
  free(o->segs);
  free(o->last_flat);
  free(o->last_scores);
  free(o->last_feat);
  free(o);
}

/* util.rs::frame_into_u8: 8-bit passes through, >8-bit is `v >> (bd - 8)` */
static uint8_t *plane_into_u8(const void *p, size_t stride_bytes, int w, int h, int bd) {
  uint8_t *out = (uint8_t *)malloc((size_t)w * h);
  for (int y = 0; y < h; ++y) {
    if (bd == 8) {
      memcpy(out + (size_t)y * w, (const uint8_t *)p + y * stride_bytes, w);
    } else {
      const uint16_t *row = (const uint16_t *)((const uint8_t *)p + y * stride_bytes);
      for (int x = 0; x < w; ++x) out[(size_t)y * w + x] = (uint8_t)(row[x] >> (bd - 8));
    }
  }
  return out;
}

/* DiffGenerator::diff_frame — call sites src/main.rs:442/462/482/502.
 * Returns 0 (Ok) or G1S_E_DIMS; NoiseStatus::Error is swallowed (the crate compares
 * the status only against DifferentType), *status_out reports it for tests. */
int g1s_oracle_diff_frame(g1s_oracle *o, const g1s_frame *src, int src_w, int src_h, const g1s_frame *den,
                          int den_w, int den_h, int ss_x, int ss_y, int monochrome) {
  if (src_w != den_w || src_h != den_h) {
    snprintf(o->err, sizeof(o->err), "Luma dimensions do not match: source %dx%d, denoised %dx%d", src_w, src_h,
             den_w, den_h);
    return G1S_E_DIMS;
  }
  const int w = src_w, h = src_h;
  const int cw = (w + ss_x) >> ss_x, ch = (h + ss_y) >> ss_y; /* plane storage size */
  uint8_t *s8[3] = {0, 0, 0}, *d8[3] = {0, 0, 0};
  int stride[3] = {w, cw, cw};
  s8[0] = plane_into_u8(src->plane[0], src->stride_bytes[0], w, h, o->src_bd);
  d8[0] = plane_into_u8(den->plane[0], den->stride_bytes[0], w, h, o->den_bd);
  if (!monochrome) {
    for (int c = 1; c < 3; ++c) {
      s8[c] = plane_into_u8(src->plane[c], src->stride_bytes[c], cw, ch, o->src_bd);
      d8[c] = plane_into_u8(den->plane[c], den->stride_bytes[c], cw, ch, o->den_bd);
    }
  }
  const int nbw = (w + BLOCK_SIZE - 1) / BLOCK_SIZE, nbh = (h + BLOCK_SIZE - 1) / BLOCK_SIZE;
  if (o->last_nb_w * o->last_nb_h != nbw * nbh) {
    free(o->last_flat);
    free(o->last_scores);
    free(o->last_feat);
    o->last_flat = (uint8_t *)malloc((size_t)nbw * nbh);
    o->last_scores = (float *)malloc(sizeof(float) * nbw * nbh);
    o->last_feat = (double *)malloc(sizeof(double) * 5 * nbw * nbh);
  }
  o->last_nb_w = nbw;
  o->last_nb_h = nbh;
  o->last_num_flat = fbf_run(o, s8[0], w, h, stride[0], o->last_flat);

  const int sub[2] = {ss_x, ss_y};
  const uint8_t *const cs[3] = {s8[0], s8[1], s8[2]};
  const uint8_t *const cd[3] = {d8[0], d8[1], d8[2]};
  const int status = noise_model_update(o, cs, cd, w, h, stride, sub, o->last_flat);
  o->last_status = status;
  if (status == ST_DIFFERENT) {
    const uint64_t cur_timestamp =
        (uint64_t)o->frame_count * 10000000ull * (uint64_t)o->fps_den / (uint64_t)o->fps_num;
    g1s_segment seg;
    get_grain_parameters(o, o->prev_timestamp, cur_timestamp, &seg);
    push_segment(o, &seg);
    noise_model_save_latest(o);
    o->prev_timestamp = cur_timestamp;
  }
  o->frame_count += 1;
  for (int c = 0; c < 3; ++c) {
    free(s8[c]);
    free(d8[c]);
  }
  return 0;
}

/* DiffGenerator::finish — call site src/main.rs:524 */
int g1s_oracle_finish(g1s_oracle *o, g1s_segment *out, size_t cap, size_t *n) {
  g1s_segment seg;
  get_grain_parameters(o, o->prev_timestamp, (uint64_t)INT64_MAX, &seg);
  push_segment(o, &seg);
  *n = o->nsegs;
  if (cap < o->nsegs) {
    o->nsegs--; /* allow a retry */
    return G1S_E_STATE;
  }
  memcpy(out, o->segs, o->nsegs * sizeof(g1s_segment));
  return 0;
}

const char *g1s_oracle_last_error(const g1s_oracle *o) { return o->err; }
int g1s_oracle_last_status(const g1s_oracle *o) { return o->last_status; }
int g1s_oracle_last_num_flat(const g1s_oracle *o) { return o->last_num_flat; }
int g1s_oracle_last_flat(const g1s_oracle *o, uint8_t *out, float *scores, double *feat) {
  const int nb = o->last_nb_w * o->last_nb_h;
  if (out) memcpy(out, o->last_flat, nb);
  if (scores) memcpy(scores, o->last_scores, sizeof(float) * nb);
  if (feat) memcpy(feat, o->last_feat, sizeof(double) * 5 * nb);
  return nb;
}
/* EXACT_INT mode only: the integer Gram of the most recent frame, [26][26] with
 * index 24 = luma tap scaled by 2^(ss_x+ss_y), 25 = centre sample. */
int64_t g1s_oracle_last_gram(const g1s_oracle *o, int c, int64_t *G) {
  memcpy(G, o->last_gram[c], sizeof(int64_t) * 26 * 26);
  return o->last_gram_nobs[c];
}
/* The normal equations add_block_observations left for the most recent frame: A [n][n] row-major and b [n]
 * (n = 24 luma, 25 chroma), in whatever accumulation mode the handle runs.  In G1SO_GRAM_REF_ORDER these are the
 * reference's own per-term f64 sums: what the engine's strict mode (gram_reforder_kernel) must reproduce bit for bit. */
int g1s_oracle_last_eqns(const g1s_oracle *o, int c, double *A, double *b) {
  const noise_state *s = &o->latest[c];
  const int n = s->eqns.n;
  memcpy(A, s->eqns.A, sizeof(double) * n * n);
  memcpy(b, s->eqns.b, sizeof(double) * n);
  return n;
}
/* state dump for debugging/tests: which = 0 latest, 1 combined */
void g1s_oracle_get_state(const g1s_oracle *o, int which, int c, double *ar_x, double *ar_gain, double *str_x,
                          int64_t *nobs) {
  const noise_state *s = which ? &o->combined[c] : &o->latest[c];
  memcpy(ar_x, s->eqns.x, sizeof(double) * s->eqns.n);
  *ar_gain = s->ar_gain;
  memcpy(str_x, s->strength.eqns.x, sizeof(double) * NUM_BINS);
  *nobs = s->num_observations;
}
double g1s_oracle_exp_fixed(double x) { return g1s_exp_fixed(x); }

/* Grain-table text — src/main.rs:525-530 ("filmgrn1") and :631-696
 * (write_film_grain_segment); note the trailing space after the sY count (:659). */
int g1s_oracle_write_table(const g1s_segment *segs, size_t n, const char *path) {
  FILE *f = fopen(path, "wb");
  if (!f) return G1S_E_IO;
  fprintf(f, "filmgrn1\n");
  for (size_t s = 0; s < n; ++s) {
    const g1s_segment *p = &segs[s];
    fprintf(f, "E %llu %llu 1 %u 1\n", (unsigned long long)p->start_time, (unsigned long long)p->end_time,
            (unsigned)p->random_seed);
    fprintf(f, "\tp %u %u %u %u %u %u %u %u %u %u %u %u\n", p->ar_coeff_lag, p->ar_coeff_shift,
            p->grain_scale_shift, p->scaling_shift, p->chroma_scaling_from_luma ? 1 : 0, p->overlap_flag ? 1 : 0,
            p->cb_mult, p->cb_luma_mult, p->cb_offset, p->cr_mult, p->cr_luma_mult, p->cr_offset);
    fprintf(f, "\tsY %u ", p->num_y_points);
    for (int i = 0; i < p->num_y_points; ++i) fprintf(f, " %u %u", p->scaling_points_y[i][0], p->scaling_points_y[i][1]);
    fprintf(f, "\n\tsCb %u", p->num_cb_points);
    for (int i = 0; i < p->num_cb_points; ++i) fprintf(f, " %u %u", p->scaling_points_cb[i][0], p->scaling_points_cb[i][1]);
    fprintf(f, "\n\tsCr %u", p->num_cr_points);
    for (int i = 0; i < p->num_cr_points; ++i) fprintf(f, " %u %u", p->scaling_points_cr[i][0], p->scaling_points_cr[i][1]);
    const int lag = p->ar_coeff_lag > 3 ? 3 : p->ar_coeff_lag;
    const int ny = 2 * lag * (lag + 1); /* coefficient count follows the lag */
    fprintf(f, "\n\tcY");
    for (int i = 0; i < ny; ++i) fprintf(f, " %d", p->ar_coeffs_y[i]);
    fprintf(f, "\n\tcCb");
    for (int i = 0; i < ny + 1; ++i) fprintf(f, " %d", p->ar_coeffs_cb[i]);
    fprintf(f, "\n\tcCr");
    for (int i = 0; i < ny + 1; ++i) fprintf(f, " %d", p->ar_coeffs_cr[i]);
    fprintf(f, "\n");
  }
  fclose(f);
  return 0;
}
