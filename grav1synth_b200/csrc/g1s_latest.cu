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

namespace g1s {

namespace {

constexpr int kLatestThreads = 256;
constexpr int kBins = 20;
constexpr int kN = 25;  // largest AR system
constexpr double kTinyD = 1.0e-16;

struct LatestSmem {
  double A[kN * kN], b[kN], x[kN];        // AR system of the current channel
  double Ac[kN * kN], bc[kN];             // elimination scratch
  double SA[kBins * kBins], Sb[kBins], Sx[kBins];  // strength system of the current channel
  double lumaSx[kBins];                   // luma's strength solution (chroma's luma-strength LUT)
  double diag[kBins], off[kBins], total;  // chains of add_measurement
  double ar_gain, luma_gain;
  double tmp[kN * kN + kN], cvec[kN];     // gauss_solve_cta scratch
  long long nobs;
  int neq, ok, flag, perm[kN];
  int bin_start[kBins + 1], bin_cnt[kBins];
};

// Gaussian elimination of g1s_model.cpp::gauss_solve on (A, b) in shared memory, n <= 25, by the whole CTA.
// Per pivot k the host does (1) one bubble pass over column k from the bottom up (adjacent row swaps), (2) for every row
// below k: c = A[i+1][k] / A[k][k], row -= c * row k.  (1) is a permutation of the rows that only depends on column k:
// thread 0 finds it on a register copy of the column, then every element moves at once.  (2) touches each row
// independently with the SAME row k, so all rows are updated at once.  Same operations, same operands, same roundings.
// tmp: n * n + n doubles of scratch; perm / cvec: n entries.  Returns (uniformly) whether the solve succeeded.
__device__ bool gauss_solve_cta(int n, double *A, double *b, double *x, double *tmp, int *perm, double *cvec, int *flag) {
  const int tid = threadIdx.x;
  for (int k = 0; k + 1 < n; ++k) {
    if (tid == 0) {
      double col[kN];
      int idx[kN];
#pragma unroll
      for (int i = 0; i < kN; ++i) {
        col[i] = i < n ? fabs(A[i * n + k]) : 0.0;
        idx[i] = i;
      }
#pragma unroll
      for (int i = kN - 1; i > 0; --i) {
        if (i < n && i > k && col[i - 1] < col[i]) {
          const double t = col[i];
          col[i] = col[i - 1], col[i - 1] = t;
          const int u = idx[i];
          idx[i] = idx[i - 1], idx[i - 1] = u;
        }
      }
#pragma unroll
      for (int i = 0; i < kN; ++i)
        if (i < n) perm[i] = idx[i];
    }
    __syncthreads();
    for (int e = tid; e < (n - k) * (n + 1); e += blockDim.x) {
      const int i = k + e / (n + 1), j = e - (i - k) * (n + 1);
      tmp[e] = j < n ? A[perm[i] * n + j] : b[perm[i]];
    }
    __syncthreads();
    for (int e = tid; e < (n - k) * (n + 1); e += blockDim.x) {
      const int i = k + e / (n + 1), j = e - (i - k) * (n + 1);
      if (j < n) A[i * n + j] = tmp[e];
      else b[i] = tmp[e];
    }
    __syncthreads();
    const double piv = A[k * n + k];
    if (fabs(piv) < kTinyD) return false;  // uniform: every thread reads the same pivot
    if (tid < n - 1 - k) cvec[tid] = __ddiv_rn(A[(k + 1 + tid) * n + k], piv);
    __syncthreads();
    for (int e = tid; e < (n - 1 - k) * (n + 1); e += blockDim.x) {
      const int r = e / (n + 1), j = e - r * (n + 1), i = k + 1 + r;
      const double c = cvec[r];
      if (j < n) A[i * n + j] = __dsub_rn(A[i * n + j], __dmul_rn(c, A[k * n + j]));
      else b[i] = __dsub_rn(b[i], __dmul_rn(c, b[k]));
    }
    __syncthreads();
  }
  if (tid == 0) {
    int ok = 1;
    for (int i = n - 1; i >= 0; --i) {
      if (fabs(A[i * n + i]) < kTinyD) {
        ok = 0;
        break;
      }
      double c = 0;
      for (int j = i + 1; j < n; ++j) c = __dadd_rn(c, __dmul_rn(A[i * n + j], x[j]));
      x[i] = __ddiv_rn(__dsub_rn(b[i], c), A[i * n + i]);
    }
    *flag = ok;
  }
  __syncthreads();
  return *flag != 0;
}

__device__ __forceinline__ int pair_idx(int i, int j) { return i * kTaps - i * (i - 1) / 2 + (j - i); }

__global__ void __launch_bounds__(kLatestThreads, 1)
latest_kernel(Geometry g, const uint8_t *__restrict__ records, RecordLayout rl, int strict, double *__restrict__ digests,
              int digest_doubles) {
  extern __shared__ __align__(16) uint8_t dyn[];
  __shared__ LatestSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.x, nb = g.nb;
  const uint8_t *rec = records + (size_t)f * rl.bytes;
  double *dg = digests + (size_t)f * digest_doubles;
  // per-block arrays of the current channel: interpolation weight, measurement, and the key (bin, 255 = no measurement)
  double *fa = reinterpret_cast<double *>(dyn);
  double *fs = fa + nb;
  uint8_t *key = reinterpret_cast<uint8_t *>(fs + nb);
  uint8_t *bin0 = key + ((nb + 15) & ~15);  // luma block-mean bin of every block (shared by the channels)
  uint16_t *lst = reinterpret_cast<uint16_t *>(bin0 + ((nb + 15) & ~15));  // measuring blocks grouped by bin, block order

  const long long num_flat = *reinterpret_cast<const long long *>(rec + rl.off_num_flat);
  const uint8_t *flat = rec + rl.off_flat;
  const uint32_t *luma_sum = reinterpret_cast<const uint32_t *>(rec + rl.off_luma_sum);
  const bool enough = num_flat > 1;
  for (int i = tid; i < digest_doubles; i += kLatestThreads) dg[i] = 0.0;
  __syncthreads();
  if (tid == 0) {
    dg[0] = enough ? 1.0 : 0.0;
    dg[2] = -1.0;
    sm.luma_gain = 1.0;
  }
  // an untouched channel digest still carries ar_gain = 1 (ChannelState after clear())
  if (tid < 3) dg[4 + tid * 459 + 375] = 1.0;
  if (!enough) return;

  // pass 1 (once per frame): block mean -> bin index and interpolation weight (NoiseStrengthSolver::get_bin_index)
  for (int b = tid; b < nb; b += kLatestThreads) {
    const int by = b / g.nbw, bx = b - by * g.nbw;
    const int lw = min(g.width - bx * kBlock, kBlock), lh = min(g.height - by * kBlock, kBlock);
    const double block_mean = __ddiv_rn((double)luma_sum[b], (double)(lw * lh));
    const double val = block_mean < 0.0 ? 0.0 : (block_mean > 255.0 ? 255.0 : block_mean);
    const double bin = __ddiv_rn(__dmul_rn((double)(kBins - 1), __dsub_rn(val, 0.0)), 255.0);
    const int i0 = (int)bin;
    bin0[b] = (uint8_t)i0;
    fa[b] = __dsub_rn(bin, (double)i0);
  }
  __syncthreads();

  int fail_channel = -1, fail_text = 0, channels = 0;
  for (int c = 0; c < g.planes; ++c) {
    const bool chroma = c != 0;
    const int n = chroma ? 25 : 24;
    channels = c + 1;
    // ---- load_equations: integer Gram (or the strict-mode f64 sums) -> A, b
    const double nss = chroma ? (double)(1 << (g.ss_x + g.ss_y)) : 1.0;
    const long long *G = reinterpret_cast<const long long *>(rec + rl.off_gram) + (size_t)c * kPairs;
    const double *F = reinterpret_cast<const double *>(rec + rl.off_gramf) + (size_t)c * kPairs;
    for (int e = tid; e < n * n + n; e += kLatestThreads) {
      const bool isb = e >= n * n;
      const int i = isb ? e - n * n : e / n, j = isb ? 25 : e - (e / n) * n;
      const int p = i <= j ? pair_idx(i, j) : pair_idx(j, i);
      const double si = i == 24 ? nss : 1.0, sj = j == 24 ? nss : 1.0;
      double v;
      if (strict) v = isb ? __ddiv_rn(F[p], si) : __ddiv_rn(F[p], __dmul_rn(si, sj));
      else v = isb ? __ddiv_rn(__ddiv_rn((double)G[p], si), 65025.0) : __ddiv_rn(__ddiv_rn((double)G[p], __dmul_rn(si, sj)), 65025.0);
      v = __dadd_rn(0.0, v);  // the host adds into a cleared system
      if (isb) sm.b[i] = v;
      else sm.A[e] = v;
    }
    if (tid < kN) sm.x[tid] = 0.0;
    if (tid == 0) sm.nobs = reinterpret_cast<const long long *>(rec + rl.off_nobs)[c];
    __syncthreads();
    // ---- ChannelState::solve_ar: solve on copies, then the gain
    for (int e = tid; e < n * n; e += kLatestThreads) sm.Ac[e] = sm.A[e];
    if (tid < n) sm.bc[tid] = sm.b[tid];
    __syncthreads();
    {
      const bool ok = gauss_solve_cta(n, sm.Ac, sm.bc, sm.x, sm.tmp, sm.perm, sm.cvec, &sm.flag);
      if (tid == 0) {
        sm.ok = ok;
        double gain = 1.0;
        if (ok) {
          const int m = n - (chroma ? 1 : 0);
          const double nobs = (double)sm.nobs;
          double var = 0;
          for (int i = 0; i < m; ++i) var = __dadd_rn(var, __ddiv_rn(sm.A[i * n + i], nobs));
          var = __ddiv_rn(var, (double)m);
          double sum_covar = 0;
          for (int i = 0; i < m; ++i) {
            double bi = sm.b[i];
            if (chroma) bi = __dsub_rn(bi, __dmul_rn(sm.A[i * n + (n - 1)], sm.x[n - 1]));
            sum_covar = __dadd_rn(sum_covar, __ddiv_rn(__dmul_rn(bi, sm.x[i]), nobs));
          }
          const double noise_var = fmax(__dsub_rn(var, sum_covar), 1e-6);
          gain = fmax(1.0, __dsqrt_rn(fmax(__ddiv_rn(var, noise_var), 1e-6)));
        } else if (chroma) {
          // chroma_fallback: zero solution except the luma-correlation tap
          for (int i = 0; i < n; ++i) sm.x[i] = 0.0;
          const int last = n - 1;
          if (fabs(sm.A[last * n + last]) > 1e-6) sm.x[last] = __ddiv_rn(sm.b[last], sm.A[last * n + last]);
        }
        sm.ar_gain = gain;
        if (!chroma) sm.luma_gain = gain;
      }
    }
    __syncthreads();
    // this channel's digest, AR part (written even when the solve failed: the host's state holds it too)
    double *p = dg + 4 + c * 459;
    for (int e = tid; e < n * n; e += kLatestThreads) {
      const int i = e / n, j = e - i * n;
      if (j >= i) p[i * n - i * (i - 1) / 2 + (j - i)] = sm.A[e];
    }
    if (tid < n) p[325 + tid] = sm.b[tid], p[350 + tid] = sm.x[tid];
    if (tid == 0) p[375] = sm.ar_gain, p[376] = (double)sm.nobs;
    if (!sm.ok && !chroma) {
      fail_channel = c, fail_text = 1;
      break;
    }
    // ---- add_noise_std_observations, per block: key + measurement of this channel
    {
      const int bw = kBlock >> (chroma ? g.ss_x : 0), bh = kBlock >> (chroma ? g.ss_y : 0);
      const int pw = g.width >> (chroma ? g.ss_x : 0), ph = g.height >> (chroma ? g.ss_y : 0);
      const int32_t *rsum = reinterpret_cast<const int32_t *>(rec + rl.off_rsum) + (size_t)c * nb;
      const uint32_t *rsq = reinterpret_cast<const uint32_t *>(rec + rl.off_rsq) + (size_t)c * nb;
      const double corr = chroma ? sm.x[24] : 0.0, gain = sm.ar_gain, lgain = sm.luma_gain;
      for (int b = tid; b < nb; b += kLatestThreads) {
        const int by = b / g.nbw, bx = b - by * g.nbw;
        const int cw = min(pw - bx * bw, bw), ch = min(ph - by * bh, bh);
        const int cnt = cw * ch;
        const bool contributes = flat[b] && cnt > kBlock;
        key[b] = contributes ? bin0[b] : 255;
        if (!contributes) continue;
        const double cn = (double)cnt;
        const double noise_mean = __ddiv_rn((double)rsum[b], cn);
        const double noise_var = __dsub_rn(__ddiv_rn((double)rsq[b], cn), __dmul_rn(noise_mean, noise_mean));
        double ls = 0.0;  // luma_gain * NoiseStrengthSolver::get_value(luma, block_mean)
        if (chroma) {
          const int i0 = bin0[b], i1 = i0 + 1 < kBins - 1 ? i0 + 1 : kBins - 1;
          const double a = fa[b];
          ls = __dmul_rn(lgain, __dadd_rn(__dmul_rn(__dsub_rn(1.0, a), sm.lumaSx[i0]), __dmul_rn(a, sm.lumaSx[i1])));
        }
        const double t = __dmul_rn(corr, ls);
        const double lo = __ddiv_rn(noise_var, 16.0), hi = __dsub_rn(noise_var, __dmul_rn(t, t));
        const double m = hi != hi ? lo : (lo > hi ? lo : hi);
        fs[b] = __ddiv_rn(__dsqrt_rn(m), gain);
      }
    }
    __syncthreads();
    // ---- NoiseStrengthSolver::add_measurement in block order.  Every entry of the system is its own chain of adds, and
    // an entry only hears from the blocks of one or two bins: (a) stable partition of the measuring blocks by bin
    // (warp ballots, block order kept), (b) one thread per entry walks its bin list(s), merged by block index.
    for (int k = warp; k < kBins; k += kLatestThreads / 32) {  // (a1) count
      int cnt = 0;
      for (int b0 = 0; b0 < nb; b0 += 32) {
        const int b = b0 + lane;
        cnt += __popc(__ballot_sync(0xffffffffu, b < nb && key[b] == k));
      }
      if (lane == 0) sm.bin_cnt[k] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
      int at = 0;
      for (int k = 0; k < kBins; ++k) sm.bin_start[k] = at, at += sm.bin_cnt[k];
      sm.bin_start[kBins] = at;
    }
    __syncthreads();
    for (int k = warp; k < kBins; k += kLatestThreads / 32) {  // (a2) fill, in block order
      int at = sm.bin_start[k];
      for (int b0 = 0; b0 < nb; b0 += 32) {
        const int b = b0 + lane;
        const bool mine = b < nb && key[b] == k;
        const uint32_t m = __ballot_sync(0xffffffffu, mine);
        if (mine) lst[at + __popc(m & ((1u << lane) - 1u))] = (uint16_t)b;
        at += __popc(m);
      }
    }
    __syncthreads();
    if (tid < 3 * kBins + 1) {  // (b)
      const int kind = tid / kBins, i = tid - kind * kBins;  // 0 diagonal, 1 off-diagonal (i, i+1), 2 rhs, 3 total
      double acc = 0.0;
      if (kind == 3) {
        // total += noise_std of every measuring block, in block order: the longest chain of the frame.  Eight blocks per
        // step so that the loads run ahead of the dependent adds.
        int neq = 0, b = 0;
        for (; b + 8 <= nb; b += 8) {
          const uint2 kk = *reinterpret_cast<const uint2 *>(key + b);
          double v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = fs[b + q];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t kb = ((q < 4 ? kk.x : kk.y) >> (8 * (q & 3))) & 0xFF;
            if (kb != 255) acc = __dadd_rn(acc, v[q]), ++neq;
          }
        }
        for (; b < nb; ++b)
          if (key[b] != 255) acc = __dadd_rn(acc, fs[b]), ++neq;
        sm.total = acc, sm.neq = neq;
      } else if (kind == 1) {  // A[i0][i1] += a (1 - a) for the blocks of bin i (bin 19 is its own i1: diagonal)
        if (i + 1 < kBins)
          for (int p = sm.bin_start[i]; p < sm.bin_start[i + 1]; ++p) {
            const double a = fa[lst[p]];
            acc = __dadd_rn(acc, __dmul_rn(a, __dsub_rn(1.0, a)));
          }
        sm.off[i] = acc;
      } else {
        // blocks of bin i (i0 == i) and of bin i - 1 (their i1 == i), in block order
        int pa = sm.bin_start[i], ea = sm.bin_start[i + 1];
        int pb = i > 0 ? sm.bin_start[i - 1] : 0, eb = i > 0 ? sm.bin_start[i] : 0;
        const bool own_i1 = i == kBins - 1;  // i1 = min(19, i0 + 1)
        while (pa < ea || pb < eb) {
          const int ba = pa < ea ? lst[pa] : 0x7fffffff, bb = pb < eb ? lst[pb] : 0x7fffffff;
          if (ba < bb) {
            const double a = fa[ba], na = __dsub_rn(1.0, a);
            if (kind == 0) {  // A[i0][i0] += (1-a)^2; then, when i1 == i0: A[i1][i0] += a(1-a); A[i1][i1] += a^2; A[i0][i1] += a(1-a)
              acc = __dadd_rn(acc, __dmul_rn(na, na));
              if (own_i1) {
                acc = __dadd_rn(acc, __dmul_rn(a, na));
                acc = __dadd_rn(acc, __dmul_rn(a, a));
                acc = __dadd_rn(acc, __dmul_rn(a, na));
              }
            } else {  // b[i0] += (1-a) s; then, when i1 == i0: b[i1] += a s
              const double sv = fs[ba];
              acc = __dadd_rn(acc, __dmul_rn(na, sv));
              if (own_i1) acc = __dadd_rn(acc, __dmul_rn(a, sv));
            }
            ++pa;
          } else {
            const double a = fa[bb];
            if (kind == 0) acc = __dadd_rn(acc, __dmul_rn(a, a));           // A[i1][i1] += a^2
            else acc = __dadd_rn(acc, __dmul_rn(a, fs[bb]));                 // b[i1] += a s
            ++pb;
          }
        }
        if (kind == 0) sm.diag[i] = acc;
        else sm.Sb[i] = acc;
      }
    }
    __syncthreads();
    // ---- NoiseStrengthSolver::solve: ridge bump of b in place, regularised copy of A, elimination
    const double mean = __ddiv_rn(sm.total, (double)sm.neq);
    const double alpha = __ddiv_rn(__dmul_rn(2.0, (double)sm.neq), (double)kBins);
    for (int e = tid; e < kBins * kBins; e += kLatestThreads) {
      const int i = e / kBins, j = e - i * kBins;
      double v = i == j ? sm.diag[i] : (j == i + 1 ? sm.off[i] : (i == j + 1 ? sm.off[j] : 0.0));
      // Ar[i][lo] -= alpha; Ar[i][i] += 2 alpha; Ar[i][hi] -= alpha (lo / hi clamp onto i at the ends); Ar[i][i] += 1/8192
      const int lo = i - 1 > 0 ? i - 1 : 0, hi = i + 1 < kBins - 1 ? i + 1 : kBins - 1;
      if (j == lo) v = __dsub_rn(v, alpha);
      if (j == i) v = __dadd_rn(v, __dmul_rn(2.0, alpha));
      if (j == hi) v = __dsub_rn(v, alpha);
      if (j == i) v = __dadd_rn(v, 1.0 / 8192.);
      sm.SA[e] = v;
    }
    if (tid < kBins) {
      sm.Sb[tid] = __dadd_rn(sm.Sb[tid], __ddiv_rn(mean, 8192.));
      sm.bc[tid] = sm.Sb[tid];
      sm.Sx[tid] = 0.0;
    }
    __syncthreads();
    {
      const bool ok = gauss_solve_cta(kBins, sm.SA, sm.bc, sm.Sx, sm.tmp, sm.perm, sm.cvec, &sm.flag);
      if (tid == 0) sm.ok = ok;
    }
    __syncthreads();
    // ---- this channel's digest, strength part
    if (tid < kBins) {
      p[377 + tid] = sm.diag[tid];
      p[397 + tid] = tid + 1 < kBins ? sm.off[tid] : 0.0;
      p[417 + tid] = sm.Sb[tid];
      p[437 + tid] = sm.Sx[tid];
      if (!chroma) sm.lumaSx[tid] = sm.Sx[tid];
    }
    if (tid == 0) p[457] = sm.total, p[458] = (double)sm.neq;
    __syncthreads();
    if (!sm.ok) {
      fail_channel = c, fail_text = 2;
      break;
    }
  }
  if (tid == 0) dg[1] = (double)channels, dg[2] = (double)fail_channel, dg[3] = (double)fail_text;
}

}  // namespace

size_t latest_smem_bytes(const Geometry &g) {
  return (size_t)g.nb * 16 + 2 * (((size_t)g.nb + 15) & ~(size_t)15) + 2 * (size_t)g.nb + 64;
}

bool latest_supported(const Geometry &g) { return latest_smem_bytes(g) <= 190 * 1024 && g.nb < 65535; }

void launch_latest(int nframes, const Geometry &g, const uint8_t *records, const RecordLayout &rl, bool strict,
                   double *digests, int digest_doubles, cudaStream_t st) {
  const size_t smem = latest_smem_bytes(g);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(latest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 190 * 1024);
    attr_set[dev & 63] = true;
  }
  latest_kernel<<<nframes, kLatestThreads, smem, st>>>(g, records, rl, strict ? 1 : 0, digests, digest_doubles);
}

}  // namespace g1s
