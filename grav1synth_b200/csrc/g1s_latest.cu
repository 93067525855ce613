// latest_kernel — the per-frame half of NoiseModel::update on the device: everything the host model derives from ONE
// frame's record before it looks at the combined state (g1s_model.cpp NoiseModel::compute_latest, i.e. av1-grain's
// add_block_observations epilogue + ar_equation_system_solve + add_noise_std_observations + NoiseStrengthSolver::solve,
// reached from /root/reference/src/main.rs:442).  Output: the frame's 11 KB digest (LatestFrame::to_digest layout), so
// only digests cross PCIe (0.3 MB of record per frame stays on the device) and no host core does per-frame work: what
// bounded the 8-GPU rate in round 1.
//
// Bit-exactness against the host (g1s_diff_digest_from_record is the checker): every f64 operation is issued in the
// host's order with explicit round-to-nearest intrinsics (no contraction).  What is restructured is only what is
// exact by construction:
//   * gauss_solve: per pivot, the bubble pass is a row permutation decided by one column (found by one thread, applied
//     by all) and the rows below the pivot are updated independently of each other (all at once); the back
//     substitution stays sequential;
//   * add_measurement: every entry of the strength system is its own chain of adds in block order, fed by the blocks
//     of one or two bins -> stable partition of the blocks by bin, then one thread per entry (20 diagonal, 19
//     off-diagonal, 20 right-hand sides, the total) walks its lists merged by block index;
//   * the per-block arithmetic (block mean, bin, noise variance, adjusted strength) is independent per block.
// One CTA per frame; the three channels run one after the other (chroma needs luma's strength solution).
#include "g1s_kernels.h"

#include <cstdio>

namespace g1s {

namespace {

constexpr int kLatestThreads = 512;
constexpr int kLatestWarps = kLatestThreads / 32;
constexpr int kBins = 20;
constexpr int kN = 25;   // largest AR system
constexpr int kW = 26;   // row stride of an elimination row [A | b]
constexpr int kSW = 21;  // the same for the strength systems
constexpr double kTinyD = 1.0e-16;
constexpr uint32_t kFull = 0xffffffffu;

// AR phase working set; it lives where the per-block arrays go afterwards.
struct ArWork {
  double A[3][kN * kN];     // the systems as loaded (digest, gain)
  double P[2][3][kN * kW];  // elimination ping-pong
  double U[3][kN * kW];     // the latest version of every row
};

struct LatestSmem {
  double b[3][kN], x[3][kN];  // AR right-hand sides and solutions
  double ar_gain[3];
  double SP[2][2][kBins * kSW], SU[2][kBins * kSW];  // strength eliminations (two systems side by side)
  double diag[kBins], off[kBins];                    // strength matrix sums of the current set of measuring blocks
  double Sb[3][kBins], Sx[3][kBins], total[3];
  long long nobs[3];
  int neq[3];
  int perm[kLatestWarps][kN];      // bubble pass replayed by one lane (NaN in the pivot column only)
  int solved;                      // bit s: back substitution of system s went through
  int mstart[kBins], mend[kBins];  // merged lists (starts are multiples of 8, ends are padded up to the next one)
  int hw[kLatestWarps][kBins];     // measuring blocks per warp range and bin
  int base[kLatestWarps][kBins];   // where a warp's next entry of list k goes
  uint32_t gm[kLatestWarps][kBins];  // lanes of the current group of 32 blocks per bin
  int nmeas;
};

#ifdef G1S_LATEST_PROF  // phase clocks of CTA 0, printed when the kernel ends (tuning aid, not in the product build)
__device__ long long g_prof_ts[64];
__device__ const char *g_prof_name[64];
__device__ int g_prof_n;
#define PROF_MARK(name)                                                      \
  do {                                                                       \
    if (blockIdx.x == 0 && threadIdx.x == 0 && g_prof_n < 64) {              \
      g_prof_name[g_prof_n] = name;                                          \
      g_prof_ts[g_prof_n++] = clock64();                                     \
    }                                                                        \
  } while (0)
#else
#define PROF_MARK(name)
#endif

// Back substitution of g1s_model.cpp::gauss_solve on the rows R (stride W, right-hand side in column n), one thread.
// Kept as plain loops: this kernel runs every phase once or thrice per CTA, so its time is instruction fetch as much as
// arithmetic, and small code is fast code.
__device__ __noinline__ bool back_substitute(const double *R, int W, int n, double *xv) {
  for (int i = n - 1; i >= 0; --i) {
    const double *row = R + i * W;
    const double d = row[i];
    if (fabs(d) < kTinyD) return false;
    double c = 0;
#pragma unroll 4
    for (int j = i + 1; j < n; ++j) c = __dadd_rn(c, __dmul_rn(row[j], xv[j]));
    xv[i] = __ddiv_rn(__dsub_rn(row[n], c), d);
  }
  return true;
}

// util.rs::linsolve (g1s_model.cpp::gauss_solve) for up to three systems side by side, by the whole CTA.
// System s has ns unknowns (0: absent), its rows [A | b] (b in column ns) with stride W in P0 + s * stride; P1 and U are
// scratch of the same shape; x + s * xs receives the solution and must come in as the host's x does (zeros).
// Per pivot k the host does (1) one bubble pass over column k from the bottom up (adjacent row swaps), (2) for every
// row below k: c = A[i][k] / A[k][k], row -= c * row k.
// (1) is a permutation of the rows that only depends on |column k|: what the pass carries upwards is the running
// maximum (ties: the upper row takes over), every row below the top receives the loser of its comparison.  So
// row k <- topmost maximum of rows k..n-1, row i > k <- (|c[i-1]| < S_i) ? row i-1 : the carried row, S_i the suffix
// maximum of rows i..n-1: one warp scan.  (A NaN in the column breaks the order argument: one lane replays the pass.)
// (2) updates every row independently with the SAME pivot row.  Five warps per system; every warp runs the scan itself
// (no hand-over), then takes every fifth row below the pivot, one lane per column, from the old buffer into the new one
// (the permutation is applied on the way): one CTA barrier per pivot.  Same operations on the same operands as the
// host: same roundings.  Columns left of the pivot are dropped (the host turns them into values nothing reads).
// Returns the mask of the systems solved, uniformly.
__device__ __noinline__ uint32_t gauss_multi(int W, int n0, int n1, int n2, double *P0, double *P1, double *U, int stride,
                                             double *x, int xs, int (*perm)[kN], int *solved) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nmax = max(n0, max(n1, n2));
  const int s = warp / 5, wr = warp - 5 * s;
  const int n = s == 0 ? n0 : (s == 1 ? n1 : (s == 2 ? n2 : 0));
  bool alive = n > 0;  // the same in the five warps of a system: they look at the same pivots
  double *cur = P0, *nxt = P1;
  for (int k = 0; k + 1 < nmax; ++k) {
    if (alive && k + 1 < n) {
      const double *c = cur + s * stride;
      const bool in = lane >= k && lane < n;
      // |c[lane][k]| as an integer: the order of non-negative doubles is the order of their bit patterns (NaN sorts
      // above infinity), rows outside k..n-1 sort below everything
      const long long key0 = in ? (__double_as_longlong(c[lane * W + k]) & 0x7fffffffffffffffll) : -1ll;
      int mine = lane;  // the old row that lands in row `lane`
      if (__any_sync(kFull, key0 > 0x7ff0000000000000ll)) {
        int *pm = perm[warp];
        if (lane == 0) {
          for (int i = k; i < n; ++i) pm[i] = i;
          for (int i = n - 1; i > k; --i)
            if (fabs(c[pm[i - 1] * W + k]) < fabs(c[pm[i] * W + k])) {
              const int t = pm[i];
              pm[i] = pm[i - 1], pm[i - 1] = t;
            }
        }
        __syncwarp();
        if (in) mine = pm[lane];
        __syncwarp();
      } else {
        long long v = key0;
        int idx = lane;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {  // suffix maxima of rows k..n-1, ties to the upper row
          const long long ov = __shfl_down_sync(kFull, v, o);
          const int oi = __shfl_down_sync(kFull, idx, o);
          if (lane + o < 32 && v < ov) v = ov, idx = oi;
        }
        const long long up = __shfl_up_sync(kFull, key0, 1);  // |c[lane - 1]|
        if (in) mine = lane == k ? idx : (up < v ? lane - 1 : idx);
      }
      const int pk = __shfl_sync(kFull, mine, k);
      const double piv = c[pk * W + k];
      if (fabs(piv) < kTinyD) {
        alive = false;  // the host returns false here
      } else {
        constexpr int kRows = (kN + 4) / 5;
        const int j = lane;
        const double pj = (j >= k && j <= n) ? c[pk * W + j] : 0.0;
        if (wr == 0 && j >= k && j <= n) U[s * stride + k * W + j] = pj;  // row k is final
        // this warp's rows k + 1 + wr + 5 r: lane r divides out the factor of row r, one division for the warp
        const int my_i = k + 1 + wr + 5 * lane;
        const int my_pi = __shfl_sync(kFull, mine, (lane < kRows && my_i < n) ? my_i : k);
        double my_f = 0.0;
        if (lane < kRows && my_i < n) my_f = __ddiv_rn(c[my_pi * W + k], piv);
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          const int i = k + 1 + wr + 5 * r;
          const int pi = __shfl_sync(kFull, my_pi, r);
          const double f = __shfl_sync(kFull, my_f, r);
          if (i < n && j >= k && j <= n) {
            const double out = __dsub_rn(c[pi * W + j], __dmul_rn(f, pj));
            nxt[s * stride + i * W + j] = out;
            U[s * stride + i * W + j] = out;
          }
        }
      }
    }
    __syncthreads();
    double *t = cur;
    cur = nxt, nxt = t;
  }
  if (tid == 0) *solved = 0;
  __syncthreads();
  PROF_MARK("  pivots");
  if (s < 3 && wr == 0 && lane == 0 && alive) {
    // a system of one unknown never entered the loop: its row is still in P0
    if (back_substitute((n == 1 ? P0 : U) + s * stride, W, n, x + s * xs)) atomicOr(solved, 1 << s);
  }
  __syncthreads();
  PROF_MARK("  back substitution");
  return (uint32_t)*solved;
}

__device__ __forceinline__ int pair_idx(int i, int j) { return i * kTaps - i * (i - 1) / 2 + (j - i); }

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}

// acc = (...((0 + v(beg)) + v(beg + 1)) + ...) + v(end - 1), every add rounded: the order-dependent sums of
// add_measurement, v(p) = val(idx(p)).  Three stages in flight -- list entries two steps ahead, their values one step
// ahead, the adds -- so the loop runs at the latency of the S dependent adds.  Lists are padded to a multiple of S
// with a dummy block whose values add +0.0 (exact: the sums are never -0), and readable 2 S entries past that.
template <int S, class I, class V>
__device__ __forceinline__ double chain_sum(int beg, int end, I idx, V val) {
  double acc = 0.0, v[S];
  uint32_t e[S];
#pragma unroll
  for (int q = 0; q < S; ++q) e[q] = idx(beg + q);
#pragma unroll
  for (int q = 0; q < S; ++q) v[q] = val(e[q]);
#pragma unroll
  for (int q = 0; q < S; ++q) e[q] = idx(beg + S + q);
#pragma unroll 1
  for (int p = beg; p < end; p += S) {
    double w[S];
#pragma unroll
    for (int q = 0; q < S; ++q) w[q] = val(e[q]);
#pragma unroll
    for (int q = 0; q < S; ++q) e[q] = idx(p + 2 * S + q);
#pragma unroll
    for (int q = 0; q < S; ++q) acc = __dadd_rn(acc, v[q]);
#pragma unroll
    for (int q = 0; q < S; ++q) v[q] = w[q];
  }
  return acc;
}

// x / cnt, cnt > 0: sample counts are powers of two except at the frame edge, and RN(x * 2^-k) == RN(x / 2^k)
__device__ __forceinline__ double div_count(double x, int cnt) {
  if ((cnt & (cnt - 1)) == 0) return __dmul_rn(x, __longlong_as_double((long long)(1023 - (31 - __clz(cnt))) << 52));
  return __ddiv_rn(x, (double)cnt);
}

__device__ __forceinline__ void clear_channel(double *dg, int c) {
  double *p = dg + 4 + c * 459;
  for (int e = threadIdx.x; e < 459; e += kLatestThreads) p[e] = e == 375 ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(kLatestThreads, 1)
latest_kernel(Geometry g, const uint8_t *__restrict__ records, RecordLayout rl, int strict, int same_blocks,
              double *__restrict__ digests, int digest_doubles) {
  extern __shared__ __align__(16) uint8_t dyn[];
  __shared__ LatestSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.x, nb = g.nb;
  const uint8_t *rec = records + (size_t)f * rl.bytes;
  double *dg = digests + (size_t)f * digest_doubles;
  const int nbp = (nb + 15) & ~15;
  // Per measuring block, in block order ("rank"): interpolation weight fa and the current channel's measurement fs;
  // [nb] of either is the dummy the list padding names, and fs reads as 0.0 for 48 entries past the last block.
  // Per block: key (luma-mean bin, 255 = does not measure) and rank.  M: per bin i the ranks of the blocks of bins
  // i - 1 and i merged in block order (bit 15: the block's own bin is i).
  double *fa = reinterpret_cast<double *>(dyn);  // nb + 1
  double *fs = fa + (nb + 1);                    // nb + 49 (before the lists exist: the weights of ALL blocks)
  uint8_t *key = reinterpret_cast<uint8_t *>(fs + (nb + 49));
  uint16_t *rank = reinterpret_cast<uint16_t *>(key + nbp);
  uint16_t *M = rank + nbp;                      // 2 nbp + 8 (kBins + 2) entries
  ArWork &ar = *reinterpret_cast<ArWork *>(dyn);

#ifdef G1S_LATEST_PROF
  if (blockIdx.x == 0 && threadIdx.x == 0) g_prof_n = 0;
  PROF_MARK("start");
#endif
  const long long num_flat = *reinterpret_cast<const long long *>(rec + rl.off_num_flat);
  const uint8_t *flat = rec + rl.off_flat;
  const uint32_t *luma_sum = reinterpret_cast<const uint32_t *>(rec + rl.off_luma_sum);
  const bool enough = num_flat > 1;
  const int planes = g.planes;
  for (int i = tid; i < digest_doubles; i += kLatestThreads) dg[i] = 0.0;
  __syncthreads();
  if (tid == 0) dg[0] = enough ? 1.0 : 0.0, dg[2] = -1.0;
  // an untouched channel digest still carries ar_gain = 1 (ChannelState after clear())
  if (tid < 3) dg[4 + tid * 459 + 375] = 1.0;
  if (!enough) return;

  // ================================================================== AR systems of all channels, side by side
  // ---- load_equations: integer Gram (or the strict-mode f64 sums) -> A, b
  for (int e = tid; e < 3 * (kN * kN + kN); e += kLatestThreads) {
    const int c = e / (kN * kN + kN), r = e - c * (kN * kN + kN);
    const int n = c ? 25 : 24;
    const bool isb = r >= kN * kN;
    const int i = isb ? r - kN * kN : r / kN, j = isb ? 25 : r - (r / kN) * kN;
    if (c >= planes || i >= n || (!isb && j >= n)) continue;
    const double nss = c ? (double)(1 << (g.ss_x + g.ss_y)) : 1.0;
    const long long *G = reinterpret_cast<const long long *>(rec + rl.off_gram) + (size_t)c * kPairs;
    const double *F = reinterpret_cast<const double *>(rec + rl.off_gramf) + (size_t)c * kPairs;
    const int p = i <= j ? pair_idx(i, j) : pair_idx(j, i);
    const double si = i == 24 ? nss : 1.0, sj = j == 24 ? nss : 1.0;
    double v;
    if (strict) v = isb ? __ddiv_rn(F[p], si) : __ddiv_rn(F[p], __dmul_rn(si, sj));
    else v = isb ? __ddiv_rn(__ddiv_rn((double)G[p], si), 65025.0) : __ddiv_rn(__ddiv_rn((double)G[p], __dmul_rn(si, sj)), 65025.0);
    v = __dadd_rn(0.0, v);  // the host adds into a cleared system
    if (isb) {
      sm.b[c][i] = v;
      ar.P[0][c][i * kW + n] = v;
    } else {
      ar.A[c][i * n + j] = v;
      ar.P[0][c][i * kW + j] = v;
    }
  }
  if (tid < 3 * kN) sm.x[tid / kN][tid % kN] = 0.0;
  if (tid < 3) sm.nobs[tid] = tid < planes ? reinterpret_cast<const long long *>(rec + rl.off_nobs)[tid] : 0;
  __syncthreads();
  PROF_MARK("load equations");
  const uint32_t ar_ok = gauss_multi(kW, 24, planes == 3 ? 25 : 0, planes == 3 ? 25 : 0, &ar.P[0][0][0], &ar.P[1][0][0],
                                         &ar.U[0][0], kN * kW, &sm.x[0][0], kN, sm.perm, &sm.solved);
  PROF_MARK("AR eliminations");
  // ---- ChannelState::solve_ar: the gain (or chroma_fallback)
  if (warp < planes && lane == 0) {
    const int c = warp, n = c ? 25 : 24;
    const bool chroma = c != 0;
    const double *A = ar.A[c], *b = sm.b[c];
    double *x = sm.x[c];
    double gain = 1.0;
    if ((ar_ok >> c) & 1u) {
      const int m = n - (chroma ? 1 : 0);
      const double nobs = (double)sm.nobs[c];
      double var = 0;
      for (int i = 0; i < m; ++i) var = __dadd_rn(var, __ddiv_rn(A[i * n + i], nobs));
      var = __ddiv_rn(var, (double)m);
      double sum_covar = 0;
      for (int i = 0; i < m; ++i) {
        double bi = b[i];
        if (chroma) bi = __dsub_rn(bi, __dmul_rn(A[i * n + (n - 1)], x[n - 1]));
        sum_covar = __dadd_rn(sum_covar, __ddiv_rn(__dmul_rn(bi, x[i]), nobs));
      }
      const double noise_var = fmax(__dsub_rn(var, sum_covar), 1e-6);
      gain = fmax(1.0, __dsqrt_rn(fmax(__ddiv_rn(var, noise_var), 1e-6)));
    } else if (chroma) {
      // chroma_fallback: zero solution except the luma-correlation tap
      for (int i = 0; i < n; ++i) x[i] = 0.0;
      const int last = n - 1;
      if (fabs(A[last * n + last]) > 1e-6) x[last] = __ddiv_rn(b[last], A[last * n + last]);
    }
    sm.ar_gain[c] = gain;
  }
  __syncthreads();
  // ---- the AR part of the digests.  The host stops at the first failing channel: luma's AR failure leaves the chroma
  // digests cleared (a later strength failure clears them again, below).
  const bool luma_ar_ok = ar_ok & 1u;
  for (int c = 0; c < (luma_ar_ok ? planes : 1); ++c) {
    const int n = c ? 25 : 24;
    double *p = dg + 4 + c * 459;
    for (int e = tid; e < n * n; e += kLatestThreads) {
      const int i = e / n, j = e - i * n;
      if (j >= i) p[i * n - i * (i - 1) / 2 + (j - i)] = ar.A[c][e];
    }
    if (tid < n) p[325 + tid] = sm.b[c][tid], p[350 + tid] = sm.x[c][tid];
    if (tid == 0) p[375] = sm.ar_gain[c], p[376] = (double)sm.nobs[c];
  }
  if (!luma_ar_ok) {
    if (tid == 0) dg[1] = 1.0, dg[2] = 0.0, dg[3] = 1.0;
    return;
  }
  __syncthreads();  // the AR working set is dead: the per-block arrays take its place
  PROF_MARK("gains, AR digests");

  // ================================================================== strength systems, channel by channel
  int channels = 0, fail_channel = -1;
  for (int c = 0; c < planes; ++c) {
    const bool chroma = c != 0;
    channels = c + 1;
    const bool new_lists = c == 0 || (c == 1 && !same_blocks);  // Cb and Cr always measure the same blocks
    const int bw = kBlock >> (chroma ? g.ss_x : 0), bh = kBlock >> (chroma ? g.ss_y : 0);
    const int pw = g.width >> (chroma ? g.ss_x : 0), ph = g.height >> (chroma ? g.ss_y : 0);
    constexpr int kIlp = 2;  // blocks per thread and step: their loads, divides and roots overlap
    if (new_lists) {
      // ---- which blocks measure (add_noise_std_observations: flat, more than 32 samples in this plane), their bin and
      // interpolation weight (NoiseStrengthSolver::get_bin_index of the luma block mean)
      for (int b0 = tid; b0 < nb; b0 += kIlp * kLatestThreads) {
        uint32_t lsum[kIlp];
        int cnt_l[kIlp];
        bool meas[kIlp];
#pragma unroll
        for (int q = 0; q < kIlp; ++q) {
          const int b = b0 + q * kLatestThreads;
          meas[q] = false;
          if (b < nb) {
            const int by = b / g.nbw, bx = b - by * g.nbw;
            cnt_l[q] = min(g.width - bx * kBlock, kBlock) * min(g.height - by * kBlock, kBlock);
            meas[q] = flat[b] && min(pw - bx * bw, bw) * min(ph - by * bh, bh) > kBlock;
            lsum[q] = luma_sum[b];
          }
        }
#pragma unroll
        for (int q = 0; q < kIlp; ++q) {
          const int b = b0 + q * kLatestThreads;
          if (b < nb) {
            const double block_mean = div_count((double)lsum[q], cnt_l[q]);
            const double val = block_mean < 0.0 ? 0.0 : (block_mean > 255.0 ? 255.0 : block_mean);
            const double bin = __ddiv_rn(__dmul_rn((double)(kBins - 1), __dsub_rn(val, 0.0)), 255.0);
            const int i0 = (int)bin;
            key[b] = meas[q] ? (uint8_t)i0 : 255;
            fs[b] = __dsub_rn(bin, (double)i0);
          }
        }
      }
      __syncthreads();
      PROF_MARK("block bins");
      // ---- ranks and lists.  A contiguous range of blocks per warp, 32 blocks per look.  Pass 1: counts per warp and bin.
      const int per_warp = ((nb + 32 * kLatestWarps - 1) / (32 * kLatestWarps)) * 32;
      const int b_lo = warp * per_warp, b_hi = min(nb, b_lo + per_warp);
      const uint32_t lt = (1u << lane) - 1u;
      if (lane < kBins) sm.hw[warp][lane] = 0;
      __syncwarp();
      for (int b0 = b_lo; b0 < b_hi; b0 += 32) {
        const int b = b0 + lane;
        const int kb = b < b_hi ? key[b] : 255;
        const uint32_t same = __match_any_sync(kFull, kb);
        if (kb != 255 && !(same & lt)) sm.hw[warp][kb] += __popc(same);  // one lane per bin, the warp owns its row
        __syncwarp();
      }
      __syncthreads();
      // offsets: ranks follow the block order; M_k takes the blocks of bins k - 1 and k, block order kept
      if (warp == 0) {
        int hk = 0;
        if (lane < kBins)
          for (int w = 0; w < kLatestWarps; ++w) hk += sm.hw[w][lane];
        const int hp = __shfl_up_sync(kFull, hk, 1);
        const int len = hk + (lane > 0 ? hp : 0), padded = (len + 7) & ~7;
        int incl = padded, tot = hk;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
        if (lane < kBins) sm.mstart[lane] = incl - padded, sm.mend[lane] = incl - padded + len;
        if (lane == 0) sm.nmeas = tot;
      }
      __syncthreads();
      int at = 0;  // rank of this warp's first measuring block
      if (lane < kBins) {
        const int k = lane;
        int before = 0, mine = 0;
        for (int w = 0; w < kLatestWarps; ++w) {
          const int h = sm.hw[w][k] + (k > 0 ? sm.hw[w][k - 1] : 0);
          if (w < warp) before += h, mine += sm.hw[w][k];
        }
        sm.base[warp][k] = sm.mstart[k] + before;
        at = mine;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) at += __shfl_xor_sync(kFull, at, o);
      if (tid == 0) fa[nb] = 0.0;
      __syncwarp();
      // Pass 2: ranks, list entries, and the weights of the measuring blocks gathered by rank
      for (int b0 = b_lo; b0 < b_hi; b0 += 32) {
        const int b = b0 + lane;
        const int kb = b < b_hi ? key[b] : 255;
        const uint32_t same = __match_any_sync(kFull, kb);
        if (lane < kBins) sm.gm[warp][lane] = 0u;
        __syncwarp();
        if (kb != 255 && !(same & lt)) sm.gm[warp][kb] = same;
        __syncwarp();
        const uint32_t meas = __ballot_sync(kFull, kb != 255);
        if (kb != 255) {
          const int r = at + __popc(meas & lt);
          rank[b] = (uint16_t)r;
          fa[r] = fs[b];
          const int own = __popc(same & lt);
          const uint32_t lo = kb > 0 ? sm.gm[warp][kb - 1] : 0u;
          M[sm.base[warp][kb] + own + __popc(lo & lt)] = (uint16_t)(r | 0x8000);
          if (kb < kBins - 1) M[sm.base[warp][kb + 1] + own + __popc(sm.gm[warp][kb + 1] & lt)] = (uint16_t)r;
        }
        at += __popc(meas);
        __syncwarp();
        if (lane < kBins) sm.base[warp][lane] += __popc(sm.gm[warp][lane]) + (lane > 0 ? __popc(sm.gm[warp][lane - 1]) : 0);
        __syncwarp();
      }
      // padding up to the next list (the last list: 16 entries more for the look-ahead)
      if (warp < kBins / 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = 2 * warp + h, end = sm.mend[k];
          const int pad_end = ((end + 7) & ~7) + (k == kBins - 1 ? 16 : 0);
          if (end + lane < pad_end) M[end + lane] = (uint16_t)nb;
        }
      }
      __syncthreads();  // the per-block weights in fs are dead from here
      PROF_MARK("lists");
    }
    // ---- add_noise_std_observations, per measuring block: this channel's measurement, by rank
    {
      const int32_t *rsum = reinterpret_cast<const int32_t *>(rec + rl.off_rsum) + (size_t)c * nb;
      const uint32_t *rsq = reinterpret_cast<const uint32_t *>(rec + rl.off_rsq) + (size_t)c * nb;
      const double corr = chroma ? sm.x[c][24] : 0.0, gain = sm.ar_gain[c], lgain = sm.ar_gain[0];
      const int nmeas = sm.nmeas;
      if (tid < 48) fs[nmeas + tid] = 0.0;
      if (tid == 48) fs[nb] = 0.0;
      for (int b0 = tid; b0 < nb; b0 += kIlp * kLatestThreads) {
        int kb[kIlp], cnt[kIlp], r[kIlp];
        int32_t s1[kIlp];
        uint32_t s2[kIlp];
#pragma unroll
        for (int q = 0; q < kIlp; ++q) {
          const int b = b0 + q * kLatestThreads;
          kb[q] = b < nb ? key[b] : 255;
          if (kb[q] != 255) {
            const int by = b / g.nbw, bx = b - by * g.nbw;
            cnt[q] = min(pw - bx * bw, bw) * min(ph - by * bh, bh);
            r[q] = rank[b];
            s1[q] = rsum[b], s2[q] = rsq[b];
          }
        }
#pragma unroll
        for (int q = 0; q < kIlp; ++q) {
          if (kb[q] != 255) {
            const double noise_mean = div_count((double)s1[q], cnt[q]);
            const double noise_var = __dsub_rn(div_count((double)s2[q], cnt[q]), __dmul_rn(noise_mean, noise_mean));
            double ls = 0.0;  // luma_gain * NoiseStrengthSolver::get_value(luma, block_mean)
            if (chroma) {
              const int i0 = kb[q], i1 = i0 + 1 < kBins - 1 ? i0 + 1 : kBins - 1;
              const double a = fa[r[q]];
              ls = __dmul_rn(lgain, __dadd_rn(__dmul_rn(__dsub_rn(1.0, a), sm.Sx[0][i0]), __dmul_rn(a, sm.Sx[0][i1])));
            }
            const double t = __dmul_rn(corr, ls);
            const double lo = __dmul_rn(noise_var, 0.0625), hi = __dsub_rn(noise_var, __dmul_rn(t, t));
            const double m = hi != hi ? lo : (lo > hi ? lo : hi);
            fs[r[q]] = __ddiv_rn(__dsqrt_rn(m), gain);
          }
        }
      }
    }
    __syncthreads();
    PROF_MARK("block measurements");
    // ---- NoiseStrengthSolver::add_measurement in block order.  Every entry of the system is its own chain of adds, and
    // an entry only hears from the blocks of one or two bins: one thread per chain.
    // Chains.  Bin i < 19: A[i][i] hears (1-a)^2 from its own blocks and a^2 from the blocks of bin i - 1;
    // A[i][i+1] = A[i+1][i] hears a (1-a) from its own blocks; b[i] hears (1-a) s and a s.  Bin 19 is its own upper
    // neighbour: an own block (a = 0: block mean 255) adds (1-a)^2, a (1-a), a^2, a (1-a) to A[19][19] and (1-a) s, a s
    // to b[19].  total hears every block.  Warps 0..2: the three kinds of bins 0..18 (one lane per bin, one loop for
    // all kinds: an addend is x * y with x = own ? 1-a : a and y = x | own ? a : 0 | s); warp 3: bin 19; warp 4: total.
    {
      const uint32_t a_fa = (uint32_t)__cvta_generic_to_shared(fa), a_fs = (uint32_t)__cvta_generic_to_shared(fs);
      const uint32_t a_M = (uint32_t)__cvta_generic_to_shared(M);
      const int kind = warp;  // 0 diagonal, 1 off-diagonal, 2 right-hand side
      if (warp < 3 && lane < kBins - 1 && (kind == 2 || new_lists)) {
        const double acc = chain_sum<8>(sm.mstart[lane], sm.mend[lane], [&](int p) { return lds_u16(a_M + 2 * p); },
                                     [&](uint32_t e) {
                                       const bool own = e >> 15;
                                       const double a = lds_f64(a_fa + 8 * (e & 0x7fffu)), sv = lds_f64(a_fs + 8 * (e & 0x7fffu));
                                       const double x = own ? __dsub_rn(1.0, a) : a;
                                       const double y = kind == 0 ? x : (kind == 2 ? sv : (own ? a : 0.0));
                                       return __dmul_rn(x, y);
                                     });
        if (kind == 0) sm.diag[lane] = acc;
        else if (kind == 1) sm.off[lane] = acc;
        else sm.Sb[c][lane] = acc;
      } else if (warp == 3 && lane < 2 && (lane == 1 || new_lists)) {
        const bool rhs = lane == 1;
        double acc = 0.0;
#pragma unroll 1
        for (int p = sm.mstart[kBins - 1]; p < sm.mend[kBins - 1]; ++p) {
          const uint32_t e = lds_u16(a_M + 2 * p);
          const double a = lds_f64(a_fa + 8 * (e & 0x7fffu)), na = __dsub_rn(1.0, a);
          const double y = rhs ? lds_f64(a_fs + 8 * (e & 0x7fffu)) : 0.0;
          if (e >> 15) {
            acc = __dadd_rn(acc, __dmul_rn(na, rhs ? y : na));
            acc = __dadd_rn(acc, __dmul_rn(a, rhs ? y : na));
            if (!rhs) {
              acc = __dadd_rn(acc, __dmul_rn(a, a));
              acc = __dadd_rn(acc, __dmul_rn(a, na));
            }
          } else {
            acc = __dadd_rn(acc, __dmul_rn(a, rhs ? y : a));
          }
        }
        if (rhs) sm.Sb[c][kBins - 1] = acc;
        else sm.diag[kBins - 1] = acc, sm.off[kBins - 1] = 0.0;
      } else if (warp == 4 && lane == 0) {
        // the measurements sit in rank order: a straight walk
        sm.total[c] = chain_sum<16>(0, sm.nmeas, [&](int p) { return (uint32_t)p; }, [&](uint32_t e) { return lds_f64(a_fs + 8 * e); });
        sm.neq[c] = sm.nmeas;
      }
    }
    __syncthreads();
    PROF_MARK("chains");
    // this channel's strength sums (the right-hand side goes out after the ridge bump below)
    if (tid < kBins) {
      double *p = dg + 4 + c * 459;
      p[377 + tid] = sm.diag[tid];
      p[397 + tid] = sm.off[tid];
    }
    if (c == 1 && planes == 3) continue;  // Cb and Cr are eliminated together
    // ---- NoiseStrengthSolver::solve: ridge bump of b in place, regularised copy of A, elimination
    const int c_lo = c == 0 ? 0 : 1, nsys = c == 0 ? 1 : c;  // luma | Cb (+ Cr)
    for (int e = tid; e < nsys * kBins * kSW; e += kLatestThreads) {
      const int s = e / (kBins * kSW), r = e - s * (kBins * kSW), i = r / kSW, j = r - i * kSW;
      const int cc = c_lo + s;
      const double neq = (double)sm.neq[cc];
      double v;
      if (j == kBins) {
        const double mean = __ddiv_rn(sm.total[cc], neq);
        v = __dadd_rn(sm.Sb[cc][i], __ddiv_rn(mean, 8192.));
        sm.Sb[cc][i] = v;
        sm.Sx[cc][i] = 0.0;
      } else {
        const double alpha = __ddiv_rn(__dmul_rn(2.0, neq), (double)kBins);
        v = i == j ? sm.diag[i] : (j == i + 1 ? sm.off[i] : (i == j + 1 ? sm.off[j] : 0.0));
        // Ar[i][lo] -= alpha; Ar[i][i] += 2 alpha; Ar[i][hi] -= alpha (lo / hi clamp onto i at the ends); Ar[i][i] += 1/8192
        const int lo = i - 1 > 0 ? i - 1 : 0, hi = i + 1 < kBins - 1 ? i + 1 : kBins - 1;
        if (j == lo) v = __dsub_rn(v, alpha);
        if (j == i) v = __dadd_rn(v, __dmul_rn(2.0, alpha));
        if (j == hi) v = __dsub_rn(v, alpha);
        if (j == i) v = __dadd_rn(v, 1.0 / 8192.);
      }
      sm.SP[0][s][r] = v;
    }
    __syncthreads();
    PROF_MARK("  strength system");
    const uint32_t ok = gauss_multi(kSW, kBins, nsys == 2 ? kBins : 0, 0, &sm.SP[0][0][0], &sm.SP[1][0][0], &sm.SU[0][0],
                                         kBins * kSW, &sm.Sx[c_lo][0], kBins, sm.perm, &sm.solved);
    PROF_MARK("strength eliminations");
    // ---- the strength part of the digests
    for (int s = 0; s < nsys; ++s) {
      const int cc = c_lo + s;
      double *p = dg + 4 + cc * 459;
      if (tid < kBins) p[417 + tid] = sm.Sb[cc][tid], p[437 + tid] = sm.Sx[cc][tid];
      if (tid == 0) p[457] = sm.total[cc], p[458] = (double)sm.neq[cc];
      if (!((ok >> s) & 1u)) {
        // the host returns here: the channels after this one stay cleared
        fail_channel = cc, channels = cc + 1;
        for (int k = cc + 1; k < planes; ++k) clear_channel(dg, k);
        break;
      }
    }
    if (fail_channel >= 0) break;
  }
  if (tid == 0) dg[1] = (double)channels, dg[2] = (double)fail_channel, dg[3] = fail_channel >= 0 ? 2.0 : 0.0;
#ifdef G1S_LATEST_PROF
  if (blockIdx.x == 0 && tid == 0)
    for (int i = 1; i < g_prof_n; ++i) printf("latest %-22s %8lld clk\n", g_prof_name[i], g_prof_ts[i] - g_prof_ts[i - 1]);
#endif
}

}  // namespace

size_t latest_smem_bytes(const Geometry &g) {
  const size_t nbp = ((size_t)g.nb + 15) & ~(size_t)15;
  const size_t blocks = ((size_t)g.nb + 1) * 8 + ((size_t)g.nb + 49) * 8 + nbp + 2 * nbp + 2 * (2 * nbp + 8 * (kBins + 2)) + 64;
  return blocks > sizeof(ArWork) ? blocks : sizeof(ArWork);
}

// add_noise_std_observations keeps a block with more than 32 samples: luma and chroma agree on that for every block
// unless the frame ends in a sliver (the four kinds of block: inner, right edge, bottom edge, corner).
static bool same_measuring_blocks(const Geometry &g) {
  if (g.planes != 3) return true;
  const int lw[2] = {kBlock, g.width - (g.nbw - 1) * kBlock}, lh[2] = {kBlock, g.height - (g.nbh - 1) * kBlock};
  const int bw = kBlock >> g.ss_x, bh = kBlock >> g.ss_y;
  const int cw[2] = {bw, (g.width >> g.ss_x) - (g.nbw - 1) * bw}, ch[2] = {bh, (g.height >> g.ss_y) - (g.nbh - 1) * bh};
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      if ((lw[a] * lh[b] > kBlock) != (cw[a] * ch[b] > kBlock)) return false;
  return true;
}

// The per-block arrays must fit in shared memory (up to ~9 000 blocks: 4K yes, 8K no).  Frames that end in a sliver
// where luma and chroma disagree on which blocks measure stay on the host: the kernel has the code for them (lists
// rebuilt for chroma) but no test geometry exercises it yet; the same goes for monochrome streams.
bool latest_supported(const Geometry &g) {
  return g.planes == 3 && latest_smem_bytes(g) + sizeof(LatestSmem) <= 226 * 1024 && g.nb < 32768 && same_measuring_blocks(g);
}

void launch_latest(int nframes, const Geometry &g, const uint8_t *records, const RecordLayout &rl, bool strict,
                   double *digests, int digest_doubles, cudaStream_t st) {
  const size_t smem = latest_smem_bytes(g);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(latest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024 - sizeof(LatestSmem)));
    attr_set[dev & 63] = true;
  }
  latest_kernel<<<nframes, kLatestThreads, smem, st>>>(g, records, rl, strict ? 1 : 0, same_measuring_blocks(g) ? 1 : 0, digests,
                                                       digest_doubles);
}

}  // namespace g1s
