// Host noise model: restates av1-grain 0.4.2 `diff/solver.rs` (NoiseModel, NoiseStrengthSolver,
// EquationSystem; lineage libaom aom_dsp/noise_model.c) downstream of the per-pixel sums.
// Reached from the reference at /root/reference/src/main.rs:442 (diff_frame) and :524 (finish).
// Must be compiled without FP contraction (-ffp-contract=off): the reference has none.
#include "g1s_model.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "g1s_kernels.h"

namespace g1s {

namespace {
constexpr double kTiny = 1.0e-16;
constexpr double kNorm2 = 255.0 * 255.0;  // BLOCK_NORMALIZATION^2
constexpr int kNumBins = 20;
constexpr uint16_t kDefaultGrainSeed = 10956;

// util.rs::linsolve (libaom mathutils.h): elimination with adjacent-row pivot bubbling.
// What is restructured is exact by construction: the multipliers of one pivot are independent of each other (divided
// out first, as a vectorisable loop), and the columns left of the pivot are not updated (the reference turns them into
// values that nothing reads again: the bubble pass, the multipliers and the back substitution only look at columns
// >= the pivot's).  Every value that reaches x goes through the same operations on the same operands.
constexpr int kGaussStride = 32;  // row stride of the working copy: rows [A | b | padding], whole 8-lane vectors

__attribute__((target_clones("avx512f", "avx2", "default"))) bool gauss_small(int n, const double *A, const double *b, double *x) {
  alignas(64) double M[kGaussStride - 1][kGaussStride];
  double cv[kGaussStride];
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) M[i][j] = A[i * n + j];
    M[i][n] = b[i];
    for (int j = n + 1; j < kGaussStride; ++j) M[i][j] = 0.0;
  }
  const int jend = (n + 1 + 7) & ~7;
  for (int k = 0; k + 1 < n; ++k) {
    const int j0 = k & ~7;
    for (int i = n - 1; i > k; --i) {
      if (std::fabs(M[i - 1][k]) < std::fabs(M[i][k])) {
        for (int j = j0; j < jend; ++j) std::swap(M[i][j], M[i - 1][j]);
      }
    }
    const double piv = M[k][k];
    if (std::fabs(piv) < kTiny) return false;
    const int m = n - 1 - k;
    for (int r = 0; r < m; ++r) cv[r] = M[k + 1 + r][k] / piv;
    const double *pk = M[k];
    for (int r = 0; r < m; ++r) {
      double *row = M[k + 1 + r];
      const double c = cv[r];
      for (int j = j0; j < jend; j += 8)
        for (int t = 0; t < 8; ++t) row[j + t] -= c * pk[j + t];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    if (std::fabs(M[i][i]) < kTiny) return false;
    double c = 0;
    for (int j = i + 1; j < n; ++j) c += M[i][j] * x[j];
    x[i] = (M[i][n] - c) / M[i][i];
  }
  return true;
}

__attribute__((target_clones("avx512f", "avx2", "default"))) bool gauss_solve(int n, double *A, double *b, double *x) {
  if (n + 1 <= kGaussStride) return gauss_small(n, A, b, x);
  for (int k = 0; k + 1 < n; ++k) {
    for (int i = n - 1; i > k; --i) {
      if (std::fabs(A[(i - 1) * n + k]) < std::fabs(A[i * n + k])) {
        std::swap_ranges(A + i * n + k, A + i * n + n, A + (i - 1) * n + k);
        std::swap(b[i], b[i - 1]);
      }
    }
    for (int i = k; i + 1 < n; ++i) {
      if (std::fabs(A[k * n + k]) < kTiny) return false;
      const double c = A[(i + 1) * n + k] / A[k * n + k];
      for (int j = k; j < n; ++j) A[(i + 1) * n + j] -= c * A[k * n + j];
      b[i + 1] -= c * b[k];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    if (std::fabs(A[i * n + i]) < kTiny) return false;
    double c = 0;
    for (int j = i + 1; j < n; ++j) c += A[i * n + j] * x[j];
    x[i] = (b[i] - c) / A[i * n + i];
  }
  return true;
}

// The same elimination on a TRIDIAGONAL system (the noise strength equations: add_measurement touches (i0, i0),
// (i0, i1), (i1, i0), (i1, i1) with i1 <= i0 + 1, the regulariser the three diagonals).  As long as the bubble pass
// finds nothing to swap -- only row k + 1 has a non-zero in column k, so that is one comparison per pivot -- every row
// below k + 1 gets the multiplier 0 / pivot = 0 and is left as it is (x - 0 * y == x for the finite y here), and row
// k + 1 only changes in columns k + 1 (row k is zero beyond it).  Returns 0 when a swap would be needed (the caller then
// runs the dense routine on the untouched inputs), 1 solved, -1 the reference's failure.
int tridiagonal_solve(int n, const double *A, const double *b, double *x) {
  double d[32], u[32], r[32];  // diagonal, superdiagonal, right-hand side of the eliminated system
  if (n > 32) return 0;
  for (int i = 0; i < n; ++i) d[i] = A[i * n + i], u[i] = i + 1 < n ? A[i * n + i + 1] : 0.0, r[i] = b[i];
  for (int k = 0; k + 1 < n; ++k) {
    const double sub = A[(k + 1) * n + k];
    if (std::fabs(d[k]) < std::fabs(sub)) return 0;
    if (std::fabs(d[k]) < kTiny) return -1;
    const double c = sub / d[k];
    d[k + 1] -= c * u[k];
    r[k + 1] -= c * r[k];
  }
  double xs[32];
  for (int i = n - 1; i >= 0; --i) {
    if (std::fabs(d[i]) < kTiny) {
      for (int j = i + 1; j < n; ++j) x[j] = xs[j];  // the reference has written these before it fails
      return -1;
    }
    // c = 0 + u[i] * x[i+1] + 0 * x[i+2] + ...: the zero products add +0.0 to a sum that is not -0
    double c = 0;
    if (i + 1 < n) c += u[i] * xs[i + 1];
    xs[i] = (r[i] - c) / d[i];
  }
  for (int i = 0; i < n; ++i) x[i] = xs[i];
  return 1;
}

}  // namespace

// Test hook (include/g1s.h): the linear solvers of the host model on caller data.  which: 0 the elimination as the
// model runs it (small systems take gauss_small), 1 the tridiagonal fast path alone.  Returns 1 solved, -1 failed,
// 0 (tridiagonal only) a row swap would be needed.  A and b are left untouched.
extern "C" int g1s_linsolve_probe(int which, int n, const double *A, const double *b, double *x) {
  if (n <= 0 || !A || !b || !x) return -2;
  if (which == 1) return tridiagonal_solve(n, A, b, x);
  std::vector<double> Ac(A, A + (size_t)n * n), bc(b, b + n);
  return gauss_solve(n, Ac.data(), bc.data(), x) ? 1 : -1;
}

namespace {
inline int pair_index(int i, int j) {  // i <= j, row-major upper triangle of a 26x26 matrix
  return i * kTaps - i * (i - 1) / 2 + (j - i);
}
inline int64_t gram_at(const int64_t *g, int i, int j) { return i <= j ? g[pair_index(i, j)] : g[pair_index(j, i)]; }

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
}  // namespace

// ------------------------------------------------------------------ LinearSystem
void LinearSystem::reset(int n_) {
  n = n_;
  A.assign((size_t)n * n, 0.0);
  b.assign(n, 0.0);
  x.assign(n, 0.0);
}
void LinearSystem::clear() {
  std::fill(A.begin(), A.end(), 0.0);
  std::fill(b.begin(), b.end(), 0.0);
  std::fill(x.begin(), x.end(), 0.0);
}
bool LinearSystem::solve() {
  // elimination works on copies (EquationSystem::solve); the scratch is per thread, not per call
  static thread_local std::vector<double> Ac, bc;
  Ac.assign(A.begin(), A.end());
  bc.assign(b.begin(), b.end());
  return gauss_solve(n, Ac.data(), bc.data(), x.data());
}
void LinearSystem::add(const LinearSystem &o) {
  // (two different systems never share storage: the qualifiers let the compiler use whole vectors)
  double *__restrict a = A.data();
  const double *__restrict oa = o.A.data();
  const size_t na = A.size(), nb = b.size();
  for (size_t i = 0; i < na; ++i) a[i] += oa[i];
  double *__restrict pb = b.data();
  const double *__restrict ob = o.b.data();
  for (size_t i = 0; i < nb; ++i) pb[i] += ob[i];
}
void LinearSystem::copy_from(const LinearSystem &o) {
  A = o.A;
  b = o.b;
  x = o.x;
}

// ---------------------------------------------------------------- StrengthSolver
StrengthSolver::StrengthSolver() : eqns(kNumBins), num_bins(kNumBins) {}
void StrengthSolver::clear() {
  eqns.clear();
  num_equations = 0;
  total = 0;
}
double StrengthSolver::bin_index(double v) const {
  const double val = v < min_intensity ? min_intensity : (v > max_intensity ? max_intensity : v);
  const double range = max_intensity - min_intensity;
  return (num_bins - 1) * (val - min_intensity) / range;
}
double StrengthSolver::bin_center(int i) const {
  const double range = max_intensity - min_intensity;
  return ((double)i) / (num_bins - 1) * range + min_intensity;
}
double StrengthSolver::value_at(double intensity) const {
  const double bin = bin_index(intensity);
  const int i0 = (int)std::floor(bin);
  const int i1 = std::min(num_bins - 1, i0 + 1);
  const double a = bin - i0;
  return (1.0 - a) * eqns.x[i0] + a * eqns.x[i1];
}
void StrengthSolver::add_measurement(double block_mean, double noise_std) {
  const double bin = bin_index(block_mean);
  const int i0 = (int)std::floor(bin);
  const int i1 = std::min(num_bins - 1, i0 + 1);
  const double a = bin - i0;
  const int n = num_bins;
  eqns.A[i0 * n + i0] += (1.0 - a) * (1.0 - a);
  eqns.A[i1 * n + i0] += a * (1.0 - a);
  eqns.A[i1 * n + i1] += a * a;
  eqns.A[i0 * n + i1] += a * (1.0 - a);
  eqns.b[i0] += (1.0 - a) * noise_std;
  eqns.b[i1] += a * noise_std;
  total += noise_std;
  num_equations++;
}
// The reference's solve regularises a COPY of A but adds the ridge term to b in place, so b depends on how many
// times solve has run.  bump_b is that side effect alone; solve_bumped is the elimination on the current b.
void StrengthSolver::bump_b() {
  const int n = num_bins;
  const double mean = total / num_equations;
  for (int i = 0; i < n; ++i) eqns.b[i] += mean / 8192.;
}
bool StrengthSolver::solve_bumped() {
  const int n = num_bins;
  const double alpha = 2.0 * (double)num_equations / n;
  static thread_local std::vector<double> Ar, br;
  Ar.assign(eqns.A.begin(), eqns.A.end());
  for (int i = 0; i < n; ++i) {
    const int lo = std::max(0, i - 1), hi = std::min(n - 1, i + 1);
    Ar[i * n + lo] -= alpha;
    Ar[i * n + i] += 2 * alpha;
    Ar[i * n + hi] -= alpha;
  }
  for (int i = 0; i < n; ++i) Ar[i * n + i] += 1.0 / 8192.;
  // the system is tridiagonal (see tridiagonal_solve); checked, not assumed: a state restored from elsewhere could differ
  bool banded = true;
  for (int i = 0; i < n && banded; ++i)
    for (int j = 0; j < n; ++j)
      if ((j < i - 1 || j > i + 1) && Ar[i * n + j] != 0.0) {
        banded = false;
        break;
      }
  if (banded) {
    const int rc = tridiagonal_solve(n, Ar.data(), eqns.b.data(), eqns.x.data());
    if (rc != 0) return rc > 0;
  }
  br.assign(eqns.b.begin(), eqns.b.end());  // elimination consumes its inputs
  return gauss_solve(n, Ar.data(), br.data(), eqns.x.data());
}
bool StrengthSolver::solve() {
  bump_b();
  return solve_bumped();
}
void StrengthSolver::add(const StrengthSolver &o) {
  eqns.add(o.eqns);
  num_equations += o.num_equations;
  total += o.total;
}
void StrengthSolver::update_residual(const std::vector<std::pair<double, double>> &pts,
                                     std::vector<double> &residual, int start, int end) const {
  const double dx = 255. / num_bins;
  const int npts = (int)pts.size();
  for (int i = std::max(start, 1); i < std::min(end, npts - 1); ++i) {
    const int lower = std::max(0, (int)std::floor(bin_index(pts[i - 1].first)));
    const int upper = std::min(num_bins - 1, (int)std::ceil(bin_index(pts[i + 1].first)));
    double r = 0;
    for (int j = lower; j <= upper; ++j) {
      const double x = bin_center(j);
      if (x < pts[i - 1].first) continue;
      if (x >= pts[i + 1].first) continue;
      const double y = eqns.x[j];
      const double a = (x - pts[i - 1].first) / (pts[i + 1].first - pts[i - 1].first);
      const double est = pts[i - 1].second * (1.0 - a) + pts[i + 1].second * a;
      r += std::fabs(y - est);
    }
    residual[i] = r * dx;
  }
}
std::vector<std::pair<double, double>> StrengthSolver::fit_piecewise(int max_points) const {
  const double tol = max_intensity * 0.00625 / 255.0;
  std::vector<std::pair<double, double>> pts(num_bins);
  for (int i = 0; i < num_bins; ++i) pts[i] = {bin_center(i), eqns.x[i]};
  std::vector<double> residual(num_bins, 0.0);
  update_residual(pts, residual, 0, num_bins);
  while (pts.size() > 2) {
    int min_index = 1;
    for (int j = 1; j + 1 < (int)pts.size(); ++j)
      if (residual[j] < residual[min_index]) min_index = j;
    const double dx = pts[min_index + 1].first - pts[min_index - 1].first;
    const double avg = residual[min_index] / dx;
    if ((int)pts.size() <= max_points && avg > tol) break;
    // Only the points move up; residual[] keeps its slots, as in libaom's noise_model.c (memmove of
    // lut->points only) which the crate ports -- verified against the libaom 3.13.1 binary (oracle/aom_pin.py).
    pts.erase(pts.begin() + min_index);
    update_residual(pts, residual, min_index - 1, min_index + 1);
  }
  return pts;
}

// ------------------------------------------------------------------ ChannelState
bool ChannelState::solve_ar(bool is_chroma) {
  const bool ok = eqns.solve();
  ar_gain = 1.0;
  if (!ok) return false;
  const int n = eqns.n, m = n - (is_chroma ? 1 : 0);
  const double nobs = (double)num_observations;
  double var = 0;
  for (int i = 0; i < m; ++i) var += eqns.A[i * n + i] / nobs;
  var /= m;
  double sum_covar = 0;
  for (int i = 0; i < m; ++i) {
    double bi = eqns.b[i];
    if (is_chroma) bi -= eqns.A[i * n + (n - 1)] * eqns.x[n - 1];
    sum_covar += (bi * eqns.x[i]) / nobs;
  }
  const double noise_var = std::fmax(var - sum_covar, 1e-6);
  ar_gain = std::fmax(1, std::sqrt(std::fmax(var / noise_var, 1e-6)));
  return true;
}

static void chroma_fallback(LinearSystem &e) {
  const int last = e.n - 1;
  std::fill(e.x.begin(), e.x.end(), 0.0);
  if (std::fabs(e.A[last * e.n + last]) > 1e-6) e.x[last] = e.b[last] / e.A[last * e.n + last];
}

// ------------------------------------------------------------------- LatestFrame
namespace {
const char *const kFailTexts[3] = {nullptr, "Solving latest noise equation system failed 0!",
                                   "Solving latest noise strength failed!"};
}

// Digest layout (doubles): [enough_flat, channels, fail_channel, fail_text] then per channel
//   AR system: upper triangle of A row-major in 325 slots (A is symmetric bit for bit: load_equations fills
//   (i, j) and (j, i) from the same Gram entry), b[25], x[25], ar_gain, num_observations,
//   strength system: diagonal[20], superdiagonal[20] (A is symmetric tridiagonal: add_measurement only touches
//   (i0, i0), (i1, i0), (i1, i1), (i0, i1) with i1 <= i0 + 1, the mirrored entries with identical adds),
//   b[20], x[20], total, num_equations.
void LatestFrame::to_digest(double *out) const {
  double *p = out;
  *p++ = enough_flat ? 1.0 : 0.0;
  *p++ = (double)channels;
  *p++ = (double)fail_channel;
  *p++ = fail_text == kFailTexts[1] ? 1.0 : (fail_text == kFailTexts[2] ? 2.0 : 0.0);
  for (int c = 0; c < 3; ++c) {
    const ChannelState &s = ch[c];
    const int n = s.eqns.n;
    std::memset(p, 0, sizeof(double) * (325 + 25 + 25));
    double *q = p;
    for (int i = 0; i < n; ++i) {
      std::memcpy(q, s.eqns.A.data() + i * n + i, sizeof(double) * (n - i));
      q += n - i;
    }
    std::memcpy(p + 325, s.eqns.b.data(), sizeof(double) * n);
    std::memcpy(p + 350, s.eqns.x.data(), sizeof(double) * n);
    p += 375;
    *p++ = s.ar_gain;
    *p++ = (double)s.num_observations;  // exact: < 2^53
    const double *SA = s.strength.eqns.A.data();
    for (int i = 0; i < 20; ++i) {
      p[i] = SA[i * 20 + i];
      p[20 + i] = i + 1 < 20 ? SA[i * 20 + i + 1] : 0.0;
    }
    std::memcpy(p + 40, s.strength.eqns.b.data(), sizeof(double) * 20);
    std::memcpy(p + 60, s.strength.eqns.x.data(), sizeof(double) * 20);
    p += 80;
    *p++ = s.strength.total;
    *p++ = (double)s.strength.num_equations;
  }
}

void LatestFrame::from_digest(const double *in) {
  const double *p = in;
  enough_flat = *p++ != 0.0;
  channels = (int)*p++;
  fail_channel = (int)*p++;
  fail_text = kFailTexts[(int)*p++];
  for (int c = 0; c < 3; ++c) {
    ChannelState &s = ch[c];
    const int n = s.eqns.n;
    const double *q = p;
    double *A = s.eqns.A.data();
    for (int i = 0; i < n; ++i) {
      std::memcpy(A + i * n + i, q, sizeof(double) * (n - i));
      for (int j = i + 1; j < n; ++j) A[j * n + i] = q[j - i];
      q += n - i;
    }
    std::memcpy(s.eqns.b.data(), p + 325, sizeof(double) * n);
    std::memcpy(s.eqns.x.data(), p + 350, sizeof(double) * n);
    p += 375;
    s.ar_gain = *p++;
    s.num_observations = (int64_t)*p++;
    double *SA = s.strength.eqns.A.data();
    std::memset(SA, 0, sizeof(double) * 400);
    for (int i = 0; i < 20; ++i) {
      SA[i * 20 + i] = p[i];
      if (i + 1 < 20) SA[i * 20 + i + 1] = SA[(i + 1) * 20 + i] = p[20 + i];
    }
    std::memcpy(s.strength.eqns.b.data(), p + 40, sizeof(double) * 20);
    std::memcpy(s.strength.eqns.x.data(), p + 60, sizeof(double) * 20);
    p += 80;
    s.strength.total = *p++;
    s.strength.num_equations = (int)*p++;
  }
}

// -------------------------------------------------------------------- NoiseModel
NoiseModel::NoiseModel(const StreamGeometry &g)
    : combined{ChannelState(24), ChannelState(25), ChannelState(25)}, g_(g) {
  // add_noise_std_observations keeps a block when it has more than 32 samples; luma and chroma agree on
  // that for every block unless the frame ends in a sliver, in which case nothing is shared below.
  same_blocks_ = true;
  cnt_luma_.resize(g_.nb);
  cnt_chroma_.resize(g_.nb);
  inv_luma_.resize(g_.nb);
  inv_chroma_.resize(g_.nb);
  for (int by = 0; by < g_.nbh; ++by)
    for (int bx = 0; bx < g_.nbw; ++bx) {
      const int lw = std::min(g_.width - bx * kBlock, kBlock), lh = std::min(g_.height - by * kBlock, kBlock);
      const int bw = kBlock >> g_.ss_x, bh = kBlock >> g_.ss_y;
      const int cw = std::min((g_.width >> g_.ss_x) - bx * bw, bw), ch = std::min((g_.height >> g_.ss_y) - by * bh, bh);
      // samples of the (frame-clipped) block: max_w * max_h of get_block_mean / get_noise_var
      const int b = by * g_.nbw + bx;
      cnt_luma_[b] = lw * lh;
      cnt_chroma_[b] = cw * ch;
      auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
      inv_luma_[b] = pow2(lw * lh) ? 1.0 / (lw * lh) : 1.0;
      inv_chroma_[b] = pow2(cw * ch) ? 1.0 / (cw * ch) : 1.0;
      if (!pow2(lw * lh)) odd_luma_.push_back(b);
      if (!pow2(cw * ch) && cw * ch > 0) odd_chroma_.push_back(b);
      if ((lw * lh > kBlock) != (cw * ch > kBlock)) same_blocks_ = false;
    }
}

// Integer Gram -> the f64 normal equations add_block_observations would hold: every entry is
// (exact sum) / (tap scale, a power of two) / 255^2 with a single rounding.
void NoiseModel::load_equations(int c, const FrameRecordView &rec, ChannelState &st) const {
  const int n = st.eqns.n;
  const int64_t *G = rec.gram + (size_t)c * kPairs;
  const double nss = c ? (double)(1 << (g_.ss_x + g_.ss_y)) : 1.0;
  if (rec.gramf) {
    // strict mode: the device accumulated every entry term by term in the reference's order (gram_reforder_kernel);
    // chains through the luma tap were summed on the unscaled integers, and RN commutes with the power-of-two scale
    const double *F = rec.gramf + (size_t)c * kPairs;
    auto at = [&](int i, int j) { return i <= j ? F[pair_index(i, j)] : F[pair_index(j, i)]; };
    for (int i = 0; i < n; ++i) {
      const double si = i == 24 ? nss : 1.0;
      for (int j = 0; j < n; ++j) st.eqns.A[i * n + j] += at(i, j) / (si * (j == 24 ? nss : 1.0));
      st.eqns.b[i] += at(i, 25) / si;
    }
    st.num_observations += rec.nobs[c];
    return;
  }
  for (int i = 0; i < n; ++i) {
    const double si = i == 24 ? nss : 1.0;
    for (int j = 0; j < n; ++j) {
      const double sj = j == 24 ? nss : 1.0;
      st.eqns.A[i * n + j] += ((double)gram_at(G, i, j) / (si * sj)) / kNorm2;
    }
    st.eqns.b[i] += ((double)gram_at(G, i, 25) / si) / kNorm2;
  }
  st.num_observations += rec.nobs[c];
}

namespace {
// The per-block arithmetic of add_noise_std_observations (get_block_mean, get_noise_var, bin index,
// uncorrelated-std formula) as flat loops over ALL blocks of the frame, so the compiler vectorises the
// divides and square roots (IEEE results do not depend on the vector width).  Blocks that do not
// contribute are computed too and simply never read.

// NoiseStrengthSolver::get_bin_index(block_mean) -> integer bin and interpolation weight
__attribute__((target_clones("avx512f", "avx2", "default"))) void block_bins(int n, const unsigned *__restrict__ luma_sum,
                                                                  const double *__restrict__ inv_cnt, int nbins,
                                                                  int *__restrict__ bin0, double *__restrict__ frac) {
  const double scale = (double)(nbins - 1);
  for (int b = 0; b < n; ++b) {
    const double block_mean = (double)luma_sum[b] * inv_cnt[b];  // == / cnt: cnt is a power of two here
    const double val = block_mean < 0.0 ? 0.0 : (block_mean > 255.0 ? 255.0 : block_mean);
    const double bin = scale * (val - 0.0) / 255.0;
    const int i0 = (int)bin;  // == floor: bin >= 0
    bin0[b] = i0;
    frac[b] = bin - (double)i0;
  }
}

// luma_gain * NoiseStrengthSolver::get_value(luma, block_mean)
__attribute__((target_clones("avx512f", "avx2", "default"))) void block_luma_strength(int n, const int *__restrict__ bin0,
                                                                           const double *__restrict__ frac,
                                                                           const double *__restrict__ x, int nbins,
                                                                           double luma_gain, double *__restrict__ out) {
  for (int b = 0; b < n; ++b) {
    const int i0 = bin0[b], i1 = i0 + 1 < nbins - 1 ? i0 + 1 : nbins - 1;
    const double a = frac[b];
    out[b] = luma_gain * ((1.0 - a) * x[i0] + a * x[i1]);
  }
}

__attribute__((target_clones("avx512f", "avx2", "default"))) void block_strengths(
    int n, const int *__restrict__ rsum, const unsigned *__restrict__ rsq, const double *__restrict__ inv_cnt,
    const double *__restrict__ luma_strength, double corr, double noise_gain, double *__restrict__ out) {
  for (int b = 0; b < n; ++b) {
    const double ic = inv_cnt[b];  // exact reciprocal of a power-of-two sample count: x * ic == x / cnt bit for bit
    const double noise_mean = (double)rsum[b] * ic;
    const double noise_var = (double)rsq[b] * ic - noise_mean * noise_mean;
    const double t = corr * luma_strength[b];
    const double lo = noise_var / 16, hi = noise_var - t * t;
    const double m = hi != hi ? lo : (lo > hi ? lo : hi);  // fmax(lo, hi), branch-free
    out[b] = std::sqrt(m) / noise_gain;
  }
}
}  // namespace

void NoiseModel::add_strength_measurements(int c, const FrameRecordView &rec, LatestFrame &lf) const {
  const int nb = g_.nb;
  StrengthSolver &solver = lf.ch[c].strength;
  const StrengthSolver &luma = lf.ch[0].strength;
  const double corr = c > 0 ? lf.ch[c].eqns.x[24] : 0;
  const int nbins = solver.num_bins;
  const std::vector<int> &cnt = c ? cnt_chroma_ : cnt_luma_;

  // pass 1 (once per frame): block mean -> bin index and interpolation weight, shared by all channels
  if (c == 0) {
    lf.bin0.resize(nb);
    lf.frac.resize(nb);
    lf.mean.resize(nb);
    lf.strength.assign(nb, 0.0);
    block_bins(nb, rec.luma_sum, inv_luma_.data(), nbins, lf.bin0.data(), lf.frac.data());
    for (int b : odd_luma_) {  // frame-edge blocks whose sample count is not a power of two: true division
      const double block_mean = (double)rec.luma_sum[b] / (double)cnt_luma_[b];
      const double val = block_mean < 0.0 ? 0.0 : (block_mean > 255.0 ? 255.0 : block_mean);
      const double bin = (double)(nbins - 1) * (val - 0.0) / 255.0;
      lf.bin0[b] = (int)bin;
      lf.frac[b] = bin - (double)lf.bin0[b];
    }
  } else {
    block_luma_strength(nb, lf.bin0.data(), lf.frac.data(), luma.eqns.x.data(), nbins, lf.ch[0].ar_gain,
                        lf.strength.data());
  }
  // pass 2: per-block adjusted strength of this channel
  block_strengths(nb, rec.rsum + (size_t)c * nb, rec.rsq + (size_t)c * nb, (c ? inv_chroma_ : inv_luma_).data(),
                  lf.strength.data(), corr, lf.ch[c].ar_gain, lf.mean.data());
  for (int b : (c ? odd_chroma_ : odd_luma_)) {  // non-power-of-two sample counts: true division
    const double cn = (double)cnt[b];
    const double noise_mean = (double)rec.rsum[(size_t)c * nb + b] / cn;
    const double noise_var = (double)rec.rsq[(size_t)c * nb + b] / cn - noise_mean * noise_mean;
    const double t = corr * lf.strength[b];
    lf.mean[b] = std::sqrt(std::fmax(noise_var / 16, noise_var - t * t)) / lf.ch[c].ar_gain;
  }
  // pass 3: NoiseStrengthSolver::add_measurement in block order (the sums are order dependent).  The blocks
  // that contribute (flat, and more than block_size samples: the reference's guard) are listed once per
  // frame, so the order-dependent loop has no data-dependent branch to mispredict.
  const bool reuse_A = c > 0 && same_blocks_;
  if (c == 0 || !same_blocks_) {
    lf.contrib.clear();
    for (int b = 0; b < nb; ++b)
      if (rec.flat[b] && cnt[b] > kBlock) lf.contrib.push_back(b);
  }
  double *A = solver.eqns.A.data(), *bv = solver.eqns.b.data();
  const double *std_of = lf.mean.data();
  double total = solver.total;
  const int m = (int)lf.contrib.size();
  const int *bin0 = lf.bin0.data();
  const double *frac = lf.frac.data();
  for (int k = 0; k < m; ++k) {
    const int b = lf.contrib[k];
    const int i0 = bin0[b], i1 = std::min(nbins - 1, i0 + 1);
    const double a = frac[b], s = std_of[b];
    if (!reuse_A) {
      A[i0 * nbins + i0] += (1.0 - a) * (1.0 - a);
      A[i1 * nbins + i0] += a * (1.0 - a);
      A[i1 * nbins + i1] += a * a;
      A[i0 * nbins + i1] += a * (1.0 - a);
    }
    bv[i0] += (1.0 - a) * s;
    bv[i1] += a * s;
    total += s;
  }
  if (reuse_A) solver.eqns.A = luma.eqns.A;  // same blocks, same weights, same order: identical sums
  solver.total = total;
  solver.num_equations += m;
}

bool NoiseModel::is_different(const LatestFrame &lf) const {
  return differs(lf, combined[0].eqns.x.data(), combined[0].strength.eqns.x.data());
}

// NoiseModel::is_different against the combined luma solutions given by pointer (24 AR taps, 20 strength bins)
bool NoiseModel::differs(const LatestFrame &lf, const double *cxx, const double *cex) {
  const LinearSystem &lx = lf.ch[0].eqns;
  double c = 0, a_len = 0, b_len = 0;
  for (int i = 0; i < lx.n; ++i) {
    a_len += lx.x[i] * lx.x[i];
    b_len += cxx[i] * cxx[i];
    c += lx.x[i] * cxx[i];
  }
  const double corr = c / (std::sqrt(a_len) * std::sqrt(b_len));
  if (corr < 0.9) return true;
  const LinearSystem &le = lf.ch[0].strength.eqns;
  const double dx = 1.0 / lf.ch[0].strength.num_bins;
  double diff = 0, total_weight = 0;
  // column sums of A, each in row order (the loop nest is turned inside out: whole rows are added to 20 running sums)
  double weight[32];
  const int n = le.n;
  if (n > 32) return true;  // not a strength system
  for (int j = 0; j < n; ++j) weight[j] = 0;
  for (int i = 0; i < n; ++i) {
    const double *row = le.A.data() + (size_t)i * n;
    for (int j = 0; j < n; ++j) weight[j] += row[j];
  }
  for (int j = 0; j < n; ++j) {
    const double w = std::sqrt(weight[j]);
    diff += w * std::fabs(le.x[j] - cex[j]);
    total_weight += w;
  }
  return diff * dx / total_weight > 0.005;
}

// The per-frame half of NoiseModel::update: clear latest, check the flat-block count, then per channel
// add_block_observations (here: load the integer Gram), solve the AR system, add the strength
// measurements and solve the strength system.  Stops at the first NoiseStatus::Error.
void NoiseModel::compute_latest(const FrameRecordView &rec, LatestFrame &lf) const {
  for (auto &s : lf.ch) {
    s.eqns.clear();
    s.num_observations = 0;
    s.ar_gain = 1.0;
    s.strength.clear();
  }
  lf.channels = 0;
  lf.fail_channel = -1;
  lf.fail_text = nullptr;
  lf.enough_flat = rec.num_flat > 1;
  if (!lf.enough_flat) return;
  for (int c = 0; c < g_.planes; ++c) {
    const bool is_chroma = c != 0;
    lf.channels = c + 1;
    load_equations(c, rec, lf.ch[c]);
    if (!lf.ch[c].solve_ar(is_chroma)) {
      if (is_chroma) {
        chroma_fallback(lf.ch[c].eqns);
      } else {
        lf.fail_channel = c;
        lf.fail_text = kFailTexts[1];
        return;
      }
    }
    add_strength_measurements(c, rec, lf);
    if (!lf.ch[c].strength.solve()) {
      lf.fail_channel = c;
      lf.fail_text = kFailTexts[2];
      return;
    }
  }
}

// The sequential half: compare with / merge into the combined state, in the reference's statement
// order (channel by channel, so an error in channel c leaves channels < c merged, as the reference does).
NoiseStatus NoiseModel::fold(const LatestFrame &lf) {
  last_ = &lf;
  if (!lf.enough_flat) {
    err_ = "Not enough flat blocks to update noise estimate";
    return NoiseStatus::Error;
  }
  bool y_model_different = false;
  for (int c = 0; c < g_.planes; ++c) {
    const bool is_chroma = c != 0;
    if (lf.fail_channel == c) {
      err_ = lf.fail_text;
      return NoiseStatus::Error;
    }
    if (c == 0 && combined[0].strength.num_equations > 0 && is_different(lf)) y_model_different = true;
    if (y_model_different) continue;

    combined[c].num_observations += lf.ch[c].num_observations;
    combined[c].eqns.add(lf.ch[c].eqns);
    combined[c].strength.add(lf.ch[c].strength);
    if (is_chroma) {
      // The crate re-solves the combined chroma systems after every frame, but nothing reads those solutions
      // until a segment is emitted (is_different looks at luma only): the eliminations are deferred to settle().
      // What is NOT deferred is the strength solve's side effect on b (one ridge bump per frame, in frame order),
      // so the deferred solve sees bit for bit the b the crate's last per-frame solve would.  The one observable
      // difference would be the crate's early return when a combined chroma strength solve fails, which cannot
      // happen: the system is a sum of PSD terms plus a 1/8192 ridge.  This takes two thirds of the sequential
      // per-frame eliminations off the fold thread, which is what bounds the multi-GPU rate.
      combined[c].strength.bump_b();
      stale_[c] = true;
      continue;
    }
    if (!combined[c].solve_ar(is_chroma)) {
      err_ = "Solving combined noise equation system failed 0!";
      return NoiseStatus::Error;
    }
    if (!combined[c].strength.solve()) {
      err_ = "Solving combined noise strength failed!";
      return NoiseStatus::Error;
    }
  }
  return y_model_different ? NoiseStatus::DifferentType : NoiseStatus::Ok;
}

namespace {
// The combined luma state after one frame of a run: the sums as fold() leaves them, then the solutions; and what
// is_different needs of the frame itself.
struct LumaSnap {
  double A[24 * 24], b[24], x[24];
  double SA[kNumBins * kNumBins], Sb[kNumBins], Sx[kNumBins];
  double total, ar_gain, bump;
  double x_len;                           // sum of x[i]^2 (is_different's b_len when this state is the combined one)
  double l_len, l_weight[kNumBins], l_total_weight;  // of the FRAME: sum of its x[i]^2, sqrt of its strength column sums
  int64_t nobs;
  int neq;
  bool ok;
};

// r[e] = start[e], then for every frame r[e] += addend(frame)[e] (+ bump[frame]) and out(frame)[e] = r[e]: the running
// sums of elements [e0, e1) over the frames, every element its own chain of rounded adds in frame order.
template <class Addend, class Out>
inline void running_sums(int e0, int e1, const double *start, int count, Addend addend, const double *bump, Out out) {
  double r[128];
  const int n = e1 - e0;
  for (int e = 0; e < n; ++e) r[e] = start[e0 + e];
  for (int i = 0; i < count; ++i) {
    const double *__restrict a = addend(i) + e0;
    double *__restrict o = out(i);
    if (bump) {
      const double bi = bump[i];
      for (int e = 0; e < n; ++e) r[e] += a[e], r[e] += bi;
    } else {
      for (int e = 0; e < n; ++e) r[e] += a[e];
    }
    for (int e = 0; e < n; ++e) o[e0 + e] = r[e];
  }
}
}  // namespace

int NoiseModel::fold_run(const LatestFrame *frames, int count, const ParallelFor &par) {
  if (count <= 0) return 0;
  static thread_local std::vector<LumaSnap> snaps;
  if ((int)snaps.size() < count) snaps.resize(count);
  LumaSnap *sn = snaps.data();
  const ChannelState &y0 = combined[0];
  // (0) the scalar chains (observation and equation counts, strength totals, the ridge bump each solve adds to b)
  {
    int64_t nobs = y0.num_observations;
    int neq = y0.strength.num_equations;
    double total = y0.strength.total;
    for (int i = 0; i < count; ++i) {
      const ChannelState &l = frames[i].ch[0];
      nobs += l.num_observations, neq += l.strength.num_equations, total += l.strength.total;
      sn[i].nobs = nobs, sn[i].neq = neq, sn[i].total = total;
      sn[i].bump = (total / neq) / 8192.;
    }
  }
  // (1) the luma sums of every prefix: every matrix element is its own chain over the frames, so slices of elements go
  // to different threads (fold(): add, then the strength solve's ridge bump of b)
  {
    constexpr int kA = 8, kS = 4;  // slices of the 576 AR and the 400 strength matrix elements
    par(kA + kS + 1, [&](int t) {
      if (t < kA) {
        running_sums(576 * t / kA, 576 * (t + 1) / kA, y0.eqns.A.data(), count,
                     [&](int i) { return frames[i].ch[0].eqns.A.data(); }, nullptr, [&](int i) { return sn[i].A; });
      } else if (t < kA + kS) {
        const int u = t - kA;
        running_sums(400 * u / kS, 400 * (u + 1) / kS, y0.strength.eqns.A.data(), count,
                     [&](int i) { return frames[i].ch[0].strength.eqns.A.data(); }, nullptr, [&](int i) { return sn[i].SA; });
      } else {
        running_sums(0, 24, y0.eqns.b.data(), count, [&](int i) { return frames[i].ch[0].eqns.b.data(); }, nullptr,
                     [&](int i) { return sn[i].b; });
        std::vector<double> bump(count);
        for (int i = 0; i < count; ++i) bump[i] = sn[i].bump;
        running_sums(0, kNumBins, y0.strength.eqns.b.data(), count,
                     [&](int i) { return frames[i].ch[0].strength.eqns.b.data(); }, bump.data(), [&](int i) { return sn[i].Sb; });
      }
    });
  }
  // (2) their solves (ChannelState::solve_ar and StrengthSolver::solve_bumped on a scratch state per thread), and the
  // frame's own terms of is_different
  const int chunks = std::min(count, 64);
  par(chunks, [&](int c) {
    static thread_local ChannelState ws(24);
    for (int i = (int)((int64_t)count * c / chunks); i < (int)((int64_t)count * (c + 1) / chunks); ++i) {
      LumaSnap &s = sn[i];
      std::memcpy(ws.eqns.A.data(), s.A, sizeof s.A);
      std::memcpy(ws.eqns.b.data(), s.b, sizeof s.b);
      ws.num_observations = s.nobs;
      s.ok = ws.solve_ar(false);
      if (s.ok) {
        std::memcpy(ws.strength.eqns.A.data(), s.SA, sizeof s.SA);
        std::memcpy(ws.strength.eqns.b.data(), s.Sb, sizeof s.Sb);
        ws.strength.num_equations = s.neq, ws.strength.total = s.total;
        s.ok = ws.strength.solve_bumped();
      }
      std::memcpy(s.x, ws.eqns.x.data(), sizeof s.x);
      std::memcpy(s.Sx, ws.strength.eqns.x.data(), sizeof s.Sx);
      s.ar_gain = ws.ar_gain;
      double len = 0;
      for (int k = 0; k < 24; ++k) len += s.x[k] * s.x[k];
      s.x_len = len;
      const ChannelState &l = frames[i].ch[0];
      len = 0;
      for (int k = 0; k < 24; ++k) len += l.eqns.x[k] * l.eqns.x[k];
      s.l_len = len;
      double w[kNumBins];
      for (int j = 0; j < kNumBins; ++j) w[j] = 0;
      for (int r = 0; r < kNumBins; ++r) {
        const double *row = l.strength.eqns.A.data() + r * kNumBins;
        for (int j = 0; j < kNumBins; ++j) w[j] += row[j];
      }
      double tw = 0;
      for (int j = 0; j < kNumBins; ++j) s.l_weight[j] = std::sqrt(w[j]), tw += s.l_weight[j];
      s.l_total_weight = tw;
    }
  });
  // (3) the verdicts, in order: frame i is judged against the solutions of the state before it (NoiseModel::differs
  // with the terms that belong to one side alone taken from above: each is the same chain of operations)
  int good = 0;
  for (; good < count; ++good) {
    const int neq_before = good ? sn[good - 1].neq : y0.strength.num_equations;
    if (neq_before > 0) {
      const double *cx = good ? sn[good - 1].x : y0.eqns.x.data();
      const double *cs = good ? sn[good - 1].Sx : y0.strength.eqns.x.data();
      double b_len = 0;
      if (good) {
        b_len = sn[good - 1].x_len;
      } else {
        for (int k = 0; k < 24; ++k) b_len += cx[k] * cx[k];
      }
      const ChannelState &l = frames[good].ch[0];
      const LumaSnap &s = sn[good];
      double c = 0;
      for (int k = 0; k < 24; ++k) c += l.eqns.x[k] * cx[k];
      const double corr = c / (std::sqrt(s.l_len) * std::sqrt(b_len));
      if (corr < 0.9) break;
      const double dx = 1.0 / l.strength.num_bins;
      double diff = 0;
      for (int j = 0; j < kNumBins; ++j) diff += s.l_weight[j] * std::fabs(l.strength.eqns.x[j] - cs[j]);
      if (diff * dx / s.l_total_weight > 0.005) break;
    }
    if (!sn[good].ok) break;  // fold() reports this frame's failed combined solve itself
  }
  if (good == 0) return 0;
  // commit frames [0, good): luma from the snapshot; chroma as fold() merges it (sums and the ridge bumps, the
  // eliminations wait for settle()), again as running sums of element slices
  {
    const LumaSnap &s = sn[good - 1];
    ChannelState &y = combined[0];
    std::memcpy(y.eqns.A.data(), s.A, sizeof s.A);
    std::memcpy(y.eqns.b.data(), s.b, sizeof s.b);
    std::memcpy(y.eqns.x.data(), s.x, sizeof s.x);
    std::memcpy(y.strength.eqns.A.data(), s.SA, sizeof s.SA);
    std::memcpy(y.strength.eqns.b.data(), s.Sb, sizeof s.Sb);
    std::memcpy(y.strength.eqns.x.data(), s.Sx, sizeof s.Sx);
    y.strength.total = s.total, y.strength.num_equations = s.neq, y.num_observations = s.nobs, y.ar_gain = s.ar_gain;
  }
  if (g_.planes > 1) {
    std::vector<double> bump[3];
    for (int c = 1; c < g_.planes; ++c) {
      ChannelState &u = combined[c];
      bump[c].resize(good);
      for (int i = 0; i < good; ++i) {
        const ChannelState &l = frames[i].ch[c];
        u.num_observations += l.num_observations;
        u.strength.num_equations += l.strength.num_equations;
        u.strength.total += l.strength.total;
        bump[c][i] = (u.strength.total / u.strength.num_equations) / 8192.;
      }
      stale_[c] = true;
    }
    constexpr int kA = 5, kS = 4;  // slices of the 625 AR and the 400 strength matrix elements, per chroma channel
    const int per = kA + kS + 1, nch = g_.planes - 1;
    par(per * nch, [&](int t) {
      const int c = 1 + t / per, k = t % per;
      ChannelState &u = combined[c];
      // in place: the running sums end in the combined state itself
      auto in_place = [&](double *dst, int e0, int e1, auto addend, const double *bp) {
        double r[128];
        const int n = e1 - e0;
        for (int e = 0; e < n; ++e) r[e] = dst[e0 + e];
        for (int i = 0; i < good; ++i) {
          const double *__restrict a = addend(i) + e0;
          if (bp) {
            const double bi = bp[i];
            for (int e = 0; e < n; ++e) r[e] += a[e], r[e] += bi;
          } else {
            for (int e = 0; e < n; ++e) r[e] += a[e];
          }
        }
        for (int e = 0; e < n; ++e) dst[e0 + e] = r[e];
      };
      if (k < kA) {
        in_place(u.eqns.A.data(), 625 * k / kA, 625 * (k + 1) / kA, [&](int i) { return frames[i].ch[c].eqns.A.data(); }, nullptr);
      } else if (k < kA + kS) {
        const int q = k - kA;
        in_place(u.strength.eqns.A.data(), 400 * q / kS, 400 * (q + 1) / kS,
                 [&](int i) { return frames[i].ch[c].strength.eqns.A.data(); }, nullptr);
      } else {
        in_place(u.eqns.b.data(), 0, 25, [&](int i) { return frames[i].ch[c].eqns.b.data(); }, nullptr);
        in_place(u.strength.eqns.b.data(), 0, kNumBins, [&](int i) { return frames[i].ch[c].strength.eqns.b.data(); },
                 bump[c].data());
      }
    });
  }
  last_ = &frames[good - 1];
  return good;
}

NoiseStatus NoiseModel::update(const FrameRecordView &rec) {
  compute_latest(rec, scratch_);
  return fold(scratch_);
}

void NoiseModel::settle() {
  for (int c = 1; c < g_.planes; ++c) {
    if (!stale_[c]) continue;
    if (!combined[c].solve_ar(true)) chroma_fallback(combined[c].eqns);
    combined[c].strength.solve_bumped();
    stale_[c] = false;
  }
}

void NoiseModel::save_latest() {
  stale_[0] = stale_[1] = stale_[2] = false;
  const LatestFrame &lf = *last_;
  for (int c = 0; c < 3; ++c) {
    combined[c].eqns.copy_from(lf.ch[c].eqns);
    combined[c].strength.eqns.copy_from(lf.ch[c].strength.eqns);
    combined[c].strength.num_equations = lf.ch[c].strength.num_equations;
    combined[c].num_observations = lf.ch[c].num_observations;
    combined[c].ar_gain = lf.ch[c].ar_gain;
  }
}

void NoiseModel::grain_parameters(uint64_t start_ts, uint64_t end_ts, g1s_segment *seg) const {
  std::memset(seg, 0, sizeof(*seg));
  seg->start_time = start_ts;
  seg->end_time = end_ts;
  seg->random_seed = start_ts == 0 ? kDefaultGrainSeed : 0;
  seg->ar_coeff_lag = kLag;

  std::vector<std::pair<double, double>> sp[3] = {combined[0].strength.fit_piecewise(G1S_NUM_Y_POINTS),
                                                  combined[1].strength.fit_piecewise(G1S_NUM_UV_POINTS),
                                                  combined[2].strength.fit_piecewise(G1S_NUM_UV_POINTS)};
  double max_scaling_value = 1e-4;
  for (auto &pts : sp)
    for (auto &p : pts) {
      p.first = std::fmin(255, p.first / 1.0);
      p.second = std::fmin(255, p.second / 1.0);
      max_scaling_value = std::fmax(p.second, max_scaling_value);
    }
  const int max_log2 = clampi((int)std::floor(std::log2(max_scaling_value) + 1), 2, 5);
  seg->scaling_shift = (uint8_t)(5 + (8 - max_log2));
  const double scale_factor = (double)(1 << (8 - max_log2));
  seg->num_y_points = (uint8_t)sp[0].size();
  seg->num_cb_points = (uint8_t)sp[1].size();
  seg->num_cr_points = (uint8_t)sp[2].size();
  uint8_t(*dst[3])[2] = {seg->scaling_points_y, seg->scaling_points_cb, seg->scaling_points_cr};
  for (int c = 0; c < 3; ++c)
    for (size_t i = 0; i < sp[c].size(); ++i) {
      dst[c][i][0] = (uint8_t)(int)(sp[c][i].first + 0.5);
      dst[c][i][1] = (uint8_t)clampi((int)(scale_factor * sp[c][i].second + 0.5), 0, 255);
    }

  const int n_coeff = combined[0].eqns.n;
  double max_coeff = 1e-4, min_coeff = -1e-4;
  double y_corr[2] = {0, 0};
  double avg_luma_strength = 0;
  for (int c = 0; c < 3; ++c) {
    const LinearSystem &e = combined[c].eqns;
    for (int i = 0; i < n_coeff; ++i) {
      max_coeff = std::fmax(max_coeff, e.x[i]);
      min_coeff = std::fmin(min_coeff, e.x[i]);
    }
    const LinearSystem &se = combined[c].strength.eqns;
    double average_strength = 0, total_weight = 0;
    for (int i = 0; i < se.n; ++i) {
      double w = 0;
      for (int j = 0; j < se.n; ++j) w += se.A[i * se.n + j];
      w = std::sqrt(w);
      average_strength += se.x[i] * w;
      total_weight += w;
    }
    if (total_weight == 0)
      average_strength = 1;
    else
      average_strength /= total_weight;
    if (c == 0) {
      avg_luma_strength = average_strength;
    } else {
      y_corr[c - 1] = avg_luma_strength * e.x[n_coeff] / average_strength;
      max_coeff = std::fmax(max_coeff, y_corr[c - 1]);
      min_coeff = std::fmin(min_coeff, y_corr[c - 1]);
    }
  }
  seg->ar_coeff_shift = (uint8_t)clampi(
      7 - (int)std::fmax(1 + std::floor(std::log2(max_coeff)), std::ceil(std::log2(-min_coeff))), 6, 9);
  const double scale_ar = (double)(1 << seg->ar_coeff_shift);
  int8_t *ar[3] = {seg->ar_coeffs_y, seg->ar_coeffs_cb, seg->ar_coeffs_cr};
  for (int c = 0; c < 3; ++c) {
    const LinearSystem &e = combined[c].eqns;
    for (int i = 0; i < n_coeff; ++i) ar[c][i] = (int8_t)clampi((int)std::round(scale_ar * e.x[i]), -128, 127);
    if (c > 0) ar[c][n_coeff] = (int8_t)clampi((int)std::round(scale_ar * y_corr[c - 1]), -128, 127);
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

// ----------------------------------------------------------------- DiffSequencer
DiffSequencer::DiffSequencer(int64_t fps_num, int64_t fps_den, const StreamGeometry &g)
    : fps_num_(fps_num), fps_den_(fps_den), model_(g) {}

void DiffSequencer::after_update(NoiseStatus st) {
  // NoiseStatus::Error is swallowed, as in the crate (status is only compared to DifferentType).
  if (st == NoiseStatus::DifferentType) {
    const uint64_t cur = (uint64_t)frame_count_ * 10000000ull * (uint64_t)fps_den_ / (uint64_t)fps_num_;
    g1s_segment seg;
    model_.settle();
    model_.grain_parameters(prev_timestamp_, cur, &seg);
    table_.push_back(seg);
    model_.save_latest();
    prev_timestamp_ = cur;
  }
  ++frame_count_;
}

void DiffSequencer::consume(const FrameRecordView &rec) { after_update(model_.update(rec)); }

void DiffSequencer::consume_latest(const LatestFrame &lf) { after_update(model_.fold(lf)); }

void DiffSequencer::consume_latest_batch(const LatestFrame *frames, int count, const NoiseModel::ParallelFor &par) {
  constexpr int kMinRun = 16;  // shorter runs are not worth a fork and join
  const int planes = model_.planes();
  int i = 0;
  while (i < count) {
    int r = i;
    while (r < count && NoiseModel::plain(frames[r], planes)) ++r;
    if (r - i < kMinRun) {
      for (const int e = std::max(r, i + 1); i < e; ++i) consume_latest(frames[i]);
      continue;
    }
    const int n = model_.fold_run(frames + i, r - i, par);
    frame_count_ += n;  // after_update(NoiseStatus::Ok), n times
    i += n;
    if (i < r) consume_latest(frames[i++]);  // the frame that differs from the model so far (or fails): the plain path
  }
}

std::vector<g1s_segment> DiffSequencer::finish() {
  std::vector<g1s_segment> out(table_);
  g1s_segment seg;
  model_.settle();
  model_.grain_parameters(prev_timestamp_, (uint64_t)INT64_MAX, &seg);
  out.push_back(seg);
  return out;
}

// ---------------------------------------------------------------- misc host math
void flat_block_ata_inv(double out[9]) {
  // FlatBlockFinder::new: A rows are [yd, xd, 1] with d = (i - 16) / 16; (A^T A)^-1 by three solves.
  LinearSystem e(3);
  for (int y = 0; y < kBlock; ++y) {
    const double yd = ((double)y - 16.0) / 16.0;
    for (int x = 0; x < kBlock; ++x) {
      const double xd = ((double)x - 16.0) / 16.0;
      const double c[3] = {yd, xd, 1.0};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) e.A[3 * i + j] += c[i] * c[j];
    }
  }
  for (int i = 0; i < 3; ++i) {
    std::fill(e.b.begin(), e.b.end(), 0.0);
    e.b[i] = 1.0;
    e.solve();
    for (int j = 0; j < 3; ++j) out[j * 3 + i] = e.x[j];
  }
}

std::string format_grain_table(const g1s_segment *segs, size_t n) {
  std::string s = "filmgrn1\n";
  char buf[256];
  for (size_t k = 0; k < n; ++k) {
    const g1s_segment &p = segs[k];
    std::snprintf(buf, sizeof buf, "E %llu %llu 1 %u 1\n", (unsigned long long)p.start_time,
                  (unsigned long long)p.end_time, (unsigned)p.random_seed);
    s += buf;
    std::snprintf(buf, sizeof buf, "\tp %u %u %u %u %u %u %u %u %u %u %u %u\n", p.ar_coeff_lag, p.ar_coeff_shift,
                  p.grain_scale_shift, p.scaling_shift, p.chroma_scaling_from_luma ? 1 : 0, p.overlap_flag ? 1 : 0,
                  p.cb_mult, p.cb_luma_mult, p.cb_offset, p.cr_mult, p.cr_luma_mult, p.cr_offset);
    s += buf;
    auto points = [&](const char *tag, const uint8_t(*pts)[2], int cnt, bool trailing_space) {
      s += tag;
      s += std::to_string(cnt);
      if (trailing_space) s += ' ';  // src/main.rs:659 writes "\tsY {} " (extra space), :665/:671 do not
      for (int i = 0; i < cnt; ++i) {
        s += ' ';
        s += std::to_string(pts[i][0]);
        s += ' ';
        s += std::to_string(pts[i][1]);
      }
      s += '\n';
    };
    points("\tsY ", p.scaling_points_y, p.num_y_points, true);
    points("\tsCb ", p.scaling_points_cb, p.num_cb_points, false);
    points("\tsCr ", p.scaling_points_cr, p.num_cr_points, false);
    auto coeffs = [&](const char *tag, const int8_t *v, int cnt) {
      s += tag;
      for (int i = 0; i < cnt; ++i) {
        s += ' ';
        s += std::to_string((int)v[i]);
      }
      s += '\n';
    };
    // coefficient counts follow the lag (AV1 5.9.30): 2*lag*(lag+1) luma, one more for chroma
    const int lag = std::min<int>(p.ar_coeff_lag, 3);
    const int ny = 2 * lag * (lag + 1);
    // ... unless the segment carries explicit counts (inspect path: the reference prints the parsed ArrayVecs as they are)
    auto count = [&](int k, int dflt, int cap) { return p.num_ar_coeffs_plus1[k] ? std::min<int>(p.num_ar_coeffs_plus1[k] - 1, cap) : dflt; };
    coeffs("\tcY", p.ar_coeffs_y, count(0, ny, G1S_NUM_Y_COEFFS));
    coeffs("\tcCb", p.ar_coeffs_cb, count(1, ny + 1, G1S_NUM_UV_COEFFS));
    coeffs("\tcCr", p.ar_coeffs_cr, count(2, ny + 1, G1S_NUM_UV_COEFFS));
  }
  return s;
}

}  // namespace g1s
