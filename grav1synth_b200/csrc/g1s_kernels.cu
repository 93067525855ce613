// sm_100a kernels of the grain-estimation hot path (grav1synth `diff`).
//
//   flat_features_kernel   FlatBlockFinder::run per 32x32 source-luma block, f64 in the
//                          reference's operation order (av1-grain diff/solver.rs; call
//                          site /root/reference/src/main.rs:442)
//   flat_select_kernel     90th-percentile score threshold + final flat flags per frame
//   gram_generic_kernel    fused residual (source - denoised) + AR normal-equation
//                          (Gram) accumulation + per-block noise statistics, exact
//                          integers, any residual magnitude (NoiseModel::
//                          add_block_observations / get_block_mean / get_noise_var)
//
// Everything that the reference computes from integers is kept in integers here
// (bit-exact, order independent); the flat-block features are genuinely f64 and are
// evaluated with explicit round-to-nearest intrinsics in the reference's summation
// order so no FMA contraction or reassociation can change a bit.
#include "g1s_kernels.h"

#include <stdio.h>

namespace g1s {

RecordLayout RecordLayout::make(int nb) {
  RecordLayout r;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o += (bytes + 7) & ~size_t(7);
    return at;
  };
  r.off_gram = take(sizeof(int64_t) * 3 * kPairs);
  r.off_nobs = take(sizeof(int64_t) * 3);
  r.off_num_flat = take(sizeof(int64_t));
  r.off_luma_sum = take(sizeof(uint32_t) * nb);
  r.off_rsum = take(sizeof(int32_t) * 3 * nb);
  r.off_rsq = take(sizeof(uint32_t) * 3 * nb);
  r.off_score = take(sizeof(float) * nb);
  r.off_flat = take(nb);
  r.off_ovf_count = take(sizeof(int64_t));
  r.off_ovf = take(3 * (size_t)nb);
  r.off_gramf = take(sizeof(double) * 3 * kPairs);
  r.bytes = (o + 15) & ~size_t(15);
  return r;
}

// ----------------------------------------------------------------------------- helpers

__device__ __forceinline__ int load_sample8(const void *base, uint32_t stride, int y, int x, int bytes, int shift) {
  const uint8_t *row = reinterpret_cast<const uint8_t *>(base) + (size_t)y * stride;
  if (bytes == 1) return row[x];
  const int v = reinterpret_cast<const uint16_t *>(row)[x];
  return (v >> shift) & 0xFF;  // util.rs::frame_into_u8: truncating shift, then `as u8`
}

// pix / 255.0 correctly rounded without a divide: q = p*(1/255); r = fma(-q,255,p); q += r*(1/255).
// Verified exhaustively for p = 0..255 (tests/test_host_logic.py::test_div255_sequence).
__device__ __forceinline__ double div255(int p) {
  const double inv = 1.0 / 255.0;
  const double pd = (double)p;
  const double q = __dmul_rn(pd, inv);
  const double r = __fma_rn(-q, 255.0, pd);
  return __fma_rn(r, inv, q);
}

// Same fixed sequence as oracle/g1s_oracle.c::g1s_exp_fixed.
__device__ __forceinline__ double exp_fixed(double x) {
  const double inv_ln2 = 1.4426950408889634074;
  const double ln2_hi = 6.93147180369123816490e-01;
  const double ln2_lo = 1.90821492927058770002e-10;
  const double shift = 6755399441055744.0;
  const double kd = __dsub_rn(__dadd_rn(__dmul_rn(x, inv_ln2), shift), shift);
  const int k = __double2int_rz(kd);
  double r = __fma_rn(-kd, ln2_hi, x);
  r = __fma_rn(-kd, ln2_lo, r);
  double p = 1.0 / 6227020800.0;
  p = __fma_rn(p, r, 1.0 / 479001600.0);
  p = __fma_rn(p, r, 1.0 / 39916800.0);
  p = __fma_rn(p, r, 1.0 / 3628800.0);
  p = __fma_rn(p, r, 1.0 / 362880.0);
  p = __fma_rn(p, r, 1.0 / 40320.0);
  p = __fma_rn(p, r, 1.0 / 5040.0);
  p = __fma_rn(p, r, 1.0 / 720.0);
  p = __fma_rn(p, r, 1.0 / 120.0);
  p = __fma_rn(p, r, 1.0 / 24.0);
  p = __fma_rn(p, r, 1.0 / 6.0);
  p = __fma_rn(p, r, 0.5);
  p = __fma_rn(p, r, 1.0);
  p = __fma_rn(p, r, 1.0);
  const double sc = __longlong_as_double((long long)(1023 + k) << 52);
  return __dmul_rn(p, sc);
}

// ------------------------------------------------------------------- flat_features_kernel
//
// One thread per (frame, block).  The reference's sums are sequential f64 chains in row-major pixel
// order, so a block is walked by one thread; parallelism comes from the 8160 blocks per 4K frame times
// the frames of a batch.  What is exact-by-construction and therefore free to restructure:
//   * pix / 255.0 comes from a 256-entry table (built with the IEEE divide) in shared memory;
//   * gx = (r - l) / 2 is an exact scaling, so sum (gx*gx) == 0.25 * sum ((r-l)*(r-l)) with identical
//     roundings: the halvings are dropped and the three gradient sums are scaled by 0.25 at the end;
//   * the plane-fit residual rows needed by the central differences live in a 2-row shared-memory
//     ring laid out [row][x][thread] (conflict-free: consecutive lanes touch consecutive doubles);
//     row y+1 is produced on the fly and overwrites row y-1 in the same step that reads it.

// Shared-memory bandwidth bounds this kernel next to the FP64 pipe (per sample: two table lookups, two ring loads, one
// ring store, all 64-bit), and latency next to that (12 warps per SM: 8 measured 21.2 us per frame, 12 15.7).  A single
// 256-entry table costs ~6 wavefronts per lookup (32 lanes, random entries, 16 bank pairs); kLutCopies = 16 interleaved
// copies give every lane of a half-warp its own bank pair: 2 wavefronts, the minimum.  One CTA of 384 threads per SM
// keeps the 12 warps per SM of the earlier 3 x 128 layout with one table instead of three: 192 KB ring + 32 KB table
// (measured: 17.7 -> 15.7 us per frame; 8 copies 17.0).  Keeping the two residual rows in registers instead of the ring
// was tried and lost (128 registers of rows: spills, or 8 warps per SM: 20 us and more).
#ifndef G1S_FLAT_THREADS
#define G1S_FLAT_THREADS 384
#endif
#ifndef G1S_FLAT_LUT
#define G1S_FLAT_LUT 16
#endif
constexpr int kFlatThreads = G1S_FLAT_THREADS;
constexpr int kLutCopies = G1S_FLAT_LUT;

#ifndef G1S_FLAT_PF
#define G1S_FLAT_PF 0
#endif
// G1S_FLAT_PF: 0 = L2 prefetch four rows ahead; 1 = L1 prefetch two rows ahead; 2 = both
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Eight consecutive source-luma samples reduced to 8 bit; coordinates clamped to the frame
// (FlatBlockFinder::extract_block clamps, it does not pad).
template <int SB>
__device__ __forceinline__ void load8(const uint8_t *__restrict__ row, int x0, int w, int shift, bool vec_ok,
                                      int (&p)[8]) {
  if (vec_ok && x0 + 8 <= w) {
    if (SB == 2) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row + 2 * x0));
      const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p[2 * i] = ((q[i] & 0xFFFFu) >> shift) & 0xFF;
        p[2 * i + 1] = ((q[i] >> 16) >> shift) & 0xFF;
      }
    } else {
      const uint2 v = __ldg(reinterpret_cast<const uint2 *>(row + x0));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p[i] = (v.x >> (8 * i)) & 0xFF;
        p[4 + i] = (v.y >> (8 * i)) & 0xFF;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int x = min(x0 + i, w - 1);
      p[i] = SB == 2 ? ((reinterpret_cast<const uint16_t *>(row)[x] >> shift) & 0xFF) : row[x];
    }
  }
}

// One 8-sample chunk as loaded (16 bytes of u16 or 8 bytes of u8); blocks that lie fully inside an
// aligned frame run one block row (four chunks) ahead of the arithmetic.
template <int SB>
struct RawChunk {
  uint32_t q[SB == 2 ? 4 : 2];
};
template <int SB>
__device__ __forceinline__ RawChunk<SB> fetch_chunk(const uint8_t *__restrict__ p) {
  RawChunk<SB> r;
  if (SB == 2) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    r.q[0] = v.x, r.q[1] = v.y, r.q[2] = v.z, r.q[3] = v.w;
  } else {
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
    r.q[0] = v.x, r.q[1] = v.y;
  }
  return r;
}
template <int SB>
__device__ __forceinline__ void unpack_chunk(const RawChunk<SB> &r, int shift, int (&p)[8]) {
  if (SB == 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      p[2 * i] = ((r.q[i] & 0xFFFFu) >> shift) & 0xFF;
      p[2 * i + 1] = ((r.q[i] >> 16) >> shift) & 0xFF;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      p[i] = (r.q[0] >> (8 * i)) & 0xFF;
      p[4 + i] = (r.q[1] >> (8 * i)) & 0xFF;
    }
  }
}

template <int SB>
__global__ void __launch_bounds__(kFlatThreads)
flat_features_kernel(const FrameDesc *__restrict__ frames, int nframes, Geometry g, FlatConsts fc,
                     uint8_t *__restrict__ records, RecordLayout rl, int aligned, const uint8_t *__restrict__ y8,
                     size_t y8_frame_bytes, uint32_t y8_pitch) {
  extern __shared__ double fsm[];  // lut[256][kLutCopies] | ring[2][32][kFlatThreads]
  double *ring = fsm + 256 * kLutCopies;
  for (int i = threadIdx.x; i < 256 * kLutCopies; i += kFlatThreads) fsm[i] = __ddiv_rn((double)(i / kLutCopies), 255.0);
  __syncthreads();
  const double *lut = fsm + (threadIdx.x % kLutCopies);  // this lane's copy: entry p at lut[p * kLutCopies]

  const int gid = blockIdx.x * kFlatThreads + threadIdx.x;
  const int total = nframes * g.nb;
  if (gid >= total) return;
  const int f = gid / g.nb;
  const int b = gid - f * g.nb;
  const int by = b / g.nbw, bx = b - by * g.nbw;
  // y8: the 8-bit source luma plane residual_kernel wrote for this batch (tensor-core path): half the bytes of a
  // 10-bit frame, always aligned, and still in L2 when this kernel runs right behind it
  const uint8_t *src = y8 ? y8 + (size_t)f * y8_frame_bytes : static_cast<const uint8_t *>(frames[f].src[0]);
  const uint32_t stride = y8 ? y8_pitch : frames[f].src_stride[0];
  const int w = g.width, h = g.height;
  const int x0 = bx * kBlock, y0 = by * kBlock;
  const int tid = threadIdx.x;
  const bool vec_ok = aligned != 0;
  const int shift = y8 ? 0 : g.src_shift;

  // --- A^T * block (multiply_mat(block, A, ., 1, 1024, 3)): three sequential chains
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  const bool fastblk = vec_ok && x0 + kBlock <= w;  // every chunk is one aligned vector load
  const uint8_t *blk0 = src + (size_t)x0 * SB;       // first sample of the block's columns in row 0 of the frame
  // samples run one block row ahead of the arithmetic (the 8-bit plane of a batch is larger than L2: these loads come
  // from DRAM): chunk c of the next row is requested when chunk c of this row is taken
  RawChunk<SB> ahead[4];
  if (fastblk) {
#pragma unroll
    for (int c = 0; c < 4; ++c) ahead[c] = fetch_chunk<SB>(blk0 + (size_t)min(y0, h - 1) * stride + (size_t)(8 * c) * SB);
  }
#pragma unroll 1
  for (int yi = 0; yi < kBlock; ++yi) {
    const double yd = (double)(yi - 16) * 0.0625;
    const uint8_t *row = src + (size_t)min(y0 + yi, h - 1) * stride;
    const uint8_t *rown = blk0 + (size_t)min(y0 + min(yi + 1, kBlock - 1), h - 1) * stride;
    if (G1S_FLAT_PF != 1) prefetch_l2(src + (size_t)min(y0 + yi + 4, h - 1) * stride + (size_t)min(x0, w - 1) * SB);
    if (G1S_FLAT_PF >= 1) prefetch_l1(src + (size_t)min(y0 + yi + 2, h - 1) * stride + (size_t)min(x0, w - 1) * SB);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int p[8];
      if (fastblk) {
        const RawChunk<SB> cur = ahead[c];
        ahead[c] = fetch_chunk<SB>(rown + (size_t)(8 * c) * SB);
        unpack_chunk<SB>(cur, shift, p);
      } else {
        load8<SB>(row, x0 + 8 * c, w, shift, vec_ok, p);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double xd = (double)(8 * c + i - 16) * 0.0625;
        const double v = lut[p[i] * kLutCopies];
        s0 = __dadd_rn(s0, __dmul_rn(v, yd));
        s1 = __dadd_rn(s1, __dmul_rn(v, xd));
        s2 = __dadd_rn(s2, v);  // v * 1.0 is exact
      }
    }
  }
  // --- (A^T A)^-1 * (A^T block)
  double pc[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double s = __dadd_rn(0.0, __dmul_rn(fc.ata_inv[r * 3 + 0], s0));
    s = __dadd_rn(s, __dmul_rn(fc.ata_inv[r * 3 + 1], s1));
    s = __dadd_rn(s, __dmul_rn(fc.ata_inv[r * 3 + 2], s2));
    pc[r] = s;
  }
  // block[i] - (A * plane_coords)[i] for one sample
  auto resid = [&](int pv, double ty, int xi) -> double {
    const double xd = (double)(xi - 16) * 0.0625;
    double fit = __dadd_rn(ty, __dmul_rn(xd, pc[1]));
    fit = __dadd_rn(fit, pc[2]);  // 1.0 * pc[2] is exact
    return __dsub_rn(lut[pv * kLutCopies], fit);
  };
  auto row_ty = [&](int yi) -> double {
    const double yd = (double)(yi - 16) * 0.0625;
    return __dadd_rn(0.0, __dmul_rn(yd, pc[0]));
  };
  // rows 0 and 1 into the ring
#pragma unroll 1
  for (int yi = 0; yi < 2; ++yi) {
    const double ty = row_ty(yi);
    const uint8_t *row = src + (size_t)min(y0 + yi, h - 1) * stride;
    double *dst = ring + ((size_t)yi * kBlock) * kFlatThreads + tid;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int p[8];
      load8<SB>(row, x0 + 8 * c, w, shift, vec_ok, p);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[(size_t)(8 * c + i) * kFlatThreads] = resid(p[i], ty, 8 * c + i);
    }
  }
  // sums of (r-l)^2, (r-l)(d-u), (d-u)^2: the reference's Gxx, Gxy, Gyy are exactly 0.25 times these
  double Dxx = 0, Dxy = 0, Dyy = 0, var = 0, mean = 0;
  if (fastblk) {
#pragma unroll
    for (int c = 0; c < 4; ++c) ahead[c] = fetch_chunk<SB>(blk0 + (size_t)min(y0 + 2, h - 1) * stride + (size_t)(8 * c) * SB);
  }
#pragma unroll 1
  for (int yi = 1; yi < kBlock - 1; ++yi) {
    const double ty = row_ty(yi + 1);
    const uint8_t *row = src + (size_t)min(y0 + yi + 1, h - 1) * stride;
    const uint8_t *rown = blk0 + (size_t)min(y0 + min(yi + 2, kBlock - 1), h - 1) * stride;
    if (G1S_FLAT_PF != 1) prefetch_l2(src + (size_t)min(y0 + yi + 5, h - 1) * stride + (size_t)min(x0, w - 1) * SB);
    if (G1S_FLAT_PF >= 1) prefetch_l1(src + (size_t)min(y0 + yi + 3, h - 1) * stride + (size_t)min(x0, w - 1) * SB);
    const double *cur = ring + ((size_t)(yi & 1) * kBlock) * kFlatThreads + tid;
    double *oth = ring + ((size_t)((yi + 1) & 1) * kBlock) * kFlatThreads + tid;  // row yi-1, becomes row yi+1
    double left = cur[0], mid = cur[kFlatThreads];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int p[8];
      if (fastblk) {
        const RawChunk<SB> curc = ahead[c];
        ahead[c] = fetch_chunk<SB>(rown + (size_t)(8 * c) * SB);
        unpack_chunk<SB>(curc, shift, p);
      } else {
        load8<SB>(row, x0 + 8 * c, w, shift, vec_ok, p);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int xi = 8 * c + i;
        const double dn = resid(p[i], ty, xi);
        if (xi >= 1 && xi <= kBlock - 2) {
          const double right = cur[(size_t)(xi + 1) * kFlatThreads];
          const double up = oth[(size_t)xi * kFlatThreads];
          const double dx = __dsub_rn(right, left);
          const double dy = __dsub_rn(dn, up);
          Dxx = __dadd_rn(Dxx, __dmul_rn(dx, dx));
          Dxy = __dadd_rn(Dxy, __dmul_rn(dx, dy));
          Dyy = __dadd_rn(Dyy, __dmul_rn(dy, dy));
          mean = __dadd_rn(mean, mid);
          var = __dadd_rn(var, __dmul_rn(mid, mid));
          left = mid;
          mid = right;
        }
        oth[(size_t)xi * kFlatThreads] = dn;
      }
    }
  }

  const double nf = 900.0;  // (BLOCK_SIZE - 2)^2
  mean = __ddiv_rn(mean, nf);
  const double Gxx = __ddiv_rn(__dmul_rn(Dxx, 0.25), nf);
  const double Gxy = __ddiv_rn(__dmul_rn(Dxy, 0.25), nf);
  const double Gyy = __ddiv_rn(__dmul_rn(Dyy, 0.25), nf);
  var = __dsub_rn(__ddiv_rn(var, nf), __dmul_rn(mean, mean));

  const double trace = __dadd_rn(Gxx, Gyy);
  const double det = __dsub_rn(__dmul_rn(Gxx, Gyy), __dmul_rn(Gxy, Gxy));
  double disc = __dsub_rn(__dmul_rn(trace, trace), __dmul_rn(4.0, det));
  disc = disc > 0.0 ? disc : 0.0;  // f64::max(x, 0.)
  const double e_sub = __dsqrt_rn(disc);
  const double e1 = __dmul_rn(__dadd_rn(trace, e_sub), 0.5);
  const double e2 = __dmul_rn(__dsub_rn(trace, e_sub), 0.5);
  const double norm = e1;
  const double ratio = __ddiv_rn(e1, e2 > 1e-6 ? e2 : 1e-6);

  const double kTrace = 0.15 / 1024.0, kRatio = 1.25, kNorm = 0.08 / 1024.0, kVar = 0.005 / 1024.0;
  const bool is_flat = (trace < kTrace) && (ratio < kRatio) && (norm < kNorm) && (var > kVar);
  double sw = __fma_rn(-6682.0, var,
                       __fma_rn(-0.2056, ratio, __fma_rn(13087.0, trace, __fma_rn(-12434.0, norm, 2.5694))));
  sw = sw < -25.0 ? -25.0 : (sw > 100.0 ? 100.0 : sw);
  const double e = exp_fixed(-sw);
  const float score = __double2float_rn(__ddiv_rn(1.0, __dadd_rn(1.0, e)));

  uint8_t *rec = records + (size_t)f * rl.bytes;
  reinterpret_cast<float *>(rec + rl.off_score)[b] = var > kVar ? score : 0.0f;
  (rec + rl.off_flat)[b] = is_flat ? 255 : 0;
}

void launch_flat_features(const FrameDesc *frames, int nframes, const Geometry &g, const FlatConsts &fc,
                          uint8_t *records, const RecordLayout &rl, bool aligned, cudaStream_t st, const uint8_t *y8,
                          size_t y8_frame_bytes, uint32_t y8_pitch) {
  const int total = nframes * g.nb;
  const int grid = (total + kFlatThreads - 1) / kFlatThreads;
  const size_t smem = sizeof(double) * (256 * kLutCopies + 2 * kBlock * kFlatThreads);
  static bool attr_set[64] = {false};  // the attribute is a per-device property of the function
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(flat_features_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(flat_features_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set[dev & 63] = true;
  }
  if (y8)
    flat_features_kernel<1><<<grid, kFlatThreads, smem, st>>>(frames, nframes, g, fc, records, rl, 1, y8, y8_frame_bytes, y8_pitch);
  else if (g.src_bytes == 2)
    flat_features_kernel<2><<<grid, kFlatThreads, smem, st>>>(frames, nframes, g, fc, records, rl, aligned ? 1 : 0, nullptr, 0, 0);
  else
    flat_features_kernel<1><<<grid, kFlatThreads, smem, st>>>(frames, nframes, g, fc, records, rl, aligned ? 1 : 0, nullptr, 0, 0);
}

// --------------------------------------------------------------------- flat_select_kernel
//
// scores.sort(); thr = scores[nb * 90 / 100]; every block with score >= thr gets flat |= 1.
// Scores are non-negative floats, so their bit patterns order like unsigned integers and
// the k-th smallest is found by a 4-pass MSB-first radix select; one CTA per frame.

__global__ void __launch_bounds__(1024)
flat_select_kernel(int nframes, Geometry g, uint8_t *__restrict__ records, RecordLayout rl) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_k;
  __shared__ unsigned s_count;
  const int f = blockIdx.x;
  uint8_t *rec = records + (size_t)f * rl.bytes;
  const unsigned *scores = reinterpret_cast<const unsigned *>(rec + rl.off_score);
  uint8_t *flat = rec + rl.off_flat;
  const int nb = g.nb;
  if (threadIdx.x == 0) {
    s_prefix = 0;
    s_k = (unsigned)(nb * 90 / 100);
    s_count = 0;
  }
  for (int pass = 3; pass >= 0; --pass) {
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned himask = pass == 3 ? 0u : (0xFFFFFFFFu << (8 * (pass + 1)));
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
      const unsigned v = scores[i];
      if ((v & himask) == prefix) atomicAdd(&hist[(v >> (8 * pass)) & 0xFF], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned k = s_k, acc = 0;
      int d = 0;
      for (; d < 256; ++d) {
        if (acc + hist[d] > k) break;
        acc += hist[d];
      }
      s_k = k - acc;
      s_prefix = prefix | ((unsigned)d << (8 * pass));
    }
    __syncthreads();
  }
  const unsigned thr = s_prefix;
  unsigned local = 0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    uint8_t v = flat[i];
    if (scores[i] >= thr) v |= 1;
    flat[i] = v;
    local += v != 0;
  }
  atomicAdd(&s_count, local);
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<int64_t *>(rec + rl.off_num_flat) = (int64_t)s_count;
}

void launch_flat_select(int nframes, const Geometry &g, uint8_t *records, const RecordLayout &rl, cudaStream_t st) {
  flat_select_kernel<<<nframes, 1024, 0, st>>>(nframes, g, records, rl);
}

// --------------------------------------------------------------------- gram_generic_kernel
//
// One CTA per (frame, plane, block row); it walks the blocks of the row, and for every
// flat block stages the residual tile (with the 3-sample halo the lag-3 taps reach) in
// shared memory as int16, then each thread owns up to two of the 351 tap pairs and sums
// tap_i * tap_j over the block's observation rectangle.  int32 per block is exact
// (<= 1024 * 1020^2 would overflow only for the luma-tap square, which is bounded by
// 256 chroma samples * 1020^2 = 2.7e8), int64 across blocks; one atomicAdd per pair per CTA.

constexpr int kGramThreads = 256;
constexpr int kTilePitch = 40;             // >= 32 + 2*3, even
constexpr int kTileRows = kBlock + kLag;   // 35
constexpr int kTileElems = kTilePitch * kTileRows;

__global__ void __launch_bounds__(kGramThreads)
gram_generic_kernel(const FrameDesc *__restrict__ frames, int nframes, Geometry g, uint8_t *__restrict__ records,
                    RecordLayout rl, int only_overflow) {
  __shared__ int16_t tile[2 * kTileElems];  // [0]: residual with halo, [1]: luma tap (same geometry)
  __shared__ int red_i[kGramThreads / 32];
  __shared__ unsigned red_u[kGramThreads / 32];
  __shared__ unsigned red_l[kGramThreads / 32];

  const int by = blockIdx.x;
  const int c = blockIdx.y;
  const int f = blockIdx.z;
  const int tid = threadIdx.x;
  const FrameDesc fd = frames[f];
  uint8_t *rec = records + (size_t)f * rl.bytes;
  const uint8_t *flat = rec + rl.off_flat;
  const uint8_t *ovf = rec + rl.off_ovf + (size_t)c * g.nb;
  if (only_overflow && *reinterpret_cast<const int64_t *>(rec + rl.off_ovf_count) == 0) return;

  const int sx = c ? g.ss_x : 0, sy = c ? g.ss_y : 0;
  const int bw = kBlock >> sx, bh = kBlock >> sy;
  const int pw = g.width >> sx, ph = g.height >> sy;  // loop extents use the floor (reference: w >> sub_log2)
  const int sw = (g.width + sx) >> sx, sh = (g.height + sy) >> sy;  // plane storage size
  const void *sp = fd.src[c], *dp = fd.den[c];
  const uint32_t ss = fd.src_stride[c], ds = fd.den_stride[c];

  // pair -> (i, j), i <= j, row-major over the upper triangle
  int pi[2], pj[2], off_i[2], off_j[2];
  long long acc64[2] = {0, 0};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    int p = tid + q * kGramThreads;
    int i = 0, j = 0;
    if (p < kPairs) {
      int rem = p;
      i = 0;
      while (rem >= kTaps - i) {
        rem -= kTaps - i;
        ++i;
      }
      j = i + rem;
    } else {
      i = j = -1;
    }
    pi[q] = i;
    pj[q] = j;
    auto tap_off = [&](int t) -> int {
      if (t < 0) return 0;
      if (t < 24) {
        const int cy = t / 7 - 3, cx = t % 7 - 3;
        return cy * kTilePitch + cx;
      }
      if (t == 24) return kTileElems;  // luma tap plane
      return 0;                         // centre sample
    };
    off_i[q] = tap_off(i);
    off_j[q] = tap_off(j);
  }
  const bool use24 = c > 0;
  long long nobs = 0;

  for (int bx = 0; bx < g.nbw; ++bx) {
    const int bidx = by * g.nbw + bx;
    if (!flat[bidx]) continue;
    if (only_overflow && !ovf[bidx]) continue;
    const int x_o = bx * bw, y_o = by * bh;
    // ---- stage residual tile: tile(ty, tx) <-> plane (y_o - 3 + ty, x_o - 3 + tx)
    for (int e = tid; e < kTileRows * (bw + 2 * kLag); e += kGramThreads) {
      const int ty = e / (bw + 2 * kLag), tx = e - ty * (bw + 2 * kLag);
      const int y = y_o - kLag + ty, x = x_o - kLag + tx;
      int r = 0;
      if (ty < bh + kLag && y >= 0 && y < sh && x >= 0 && x < sw)
        r = load_sample8(sp, ss, y, x, g.src_bytes, g.src_shift) - load_sample8(dp, ds, y, x, g.den_bytes, g.den_shift);
      tile[ty * kTilePitch + tx] = (int16_t)r;
    }
    if (use24) {
      // luma tap: sum over the co-sited luma samples of (source - denoised); the reference divides
      // by the sample count (a power of two), the host undoes the scale exactly.
      for (int e = tid; e < bh * bw; e += kGramThreads) {
        const int yy = e / bw, xx = e - yy * bw;
        const int y = y_o + yy, x = x_o + xx;
        int l = 0;
        if (y < ph && x < pw) {
          for (int dy = 0; dy < (1 << sy); ++dy)
            for (int dx = 0; dx < (1 << sx); ++dx) {
              const int ly = (y << sy) + dy, lx = (x << sx) + dx;
              l += load_sample8(fd.src[0], fd.src_stride[0], ly, lx, g.src_bytes, g.src_shift) -
                   load_sample8(fd.den[0], fd.den_stride[0], ly, lx, g.den_bytes, g.den_shift);
            }
        }
        tile[kTileElems + (yy + kLag) * kTilePitch + (xx + kLag)] = (int16_t)l;
      }
    }
    __syncthreads();

    // ---- per-block noise statistics over the frame-clipped block (get_block_mean / get_noise_var);
    // also in overflow mode: the tensor-core kernel's dp4a statistics assume int8 residuals
    {
      const int max_w = min(pw - x_o, bw), max_h = min(ph - y_o, bh);
      int rs = 0;
      unsigned rq = 0;
      for (int e = tid; e < max_w * max_h; e += kGramThreads) {
        const int yy = e / max_w, xx = e - yy * max_w;
        const int r = tile[(yy + kLag) * kTilePitch + xx + kLag];
        rs += r;
        rq += (unsigned)(r * r);
      }
      unsigned ls = 0;
      if (c == 0) {
        const int lw = min(g.width - x_o, kBlock), lh = min(g.height - y_o, kBlock);
        for (int e = tid; e < lw * lh; e += kGramThreads) {
          const int yy = e / lw, xx = e - yy * lw;
          ls += (unsigned)load_sample8(sp, ss, y_o + yy, x_o + xx, g.src_bytes, g.src_shift);
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        rs += __shfl_xor_sync(0xffffffffu, rs, o);
        rq += __shfl_xor_sync(0xffffffffu, rq, o);
        ls += __shfl_xor_sync(0xffffffffu, ls, o);
      }
      if ((tid & 31) == 0) {
        red_i[tid >> 5] = rs;
        red_u[tid >> 5] = rq;
        red_l[tid >> 5] = ls;
      }
      __syncthreads();
      if (tid == 0) {
        int a = 0;
        unsigned bq = 0, cl = 0;
        for (int k = 0; k < kGramThreads / 32; ++k) {
          a += red_i[k];
          bq += red_u[k];
          cl += red_l[k];
        }
        reinterpret_cast<int32_t *>(rec + rl.off_rsum)[c * g.nb + bidx] = a;
        reinterpret_cast<uint32_t *>(rec + rl.off_rsq)[c * g.nb + bidx] = bq;
        if (c == 0) reinterpret_cast<uint32_t *>(rec + rl.off_luma_sum)[bidx] = cl;
      }
    }

    // ---- observation rectangle (add_block_observations)
    const int y_start = (by > 0 && flat[bidx - g.nbw]) ? 0 : kLag;
    const int x_start = (bx > 0 && flat[bidx - 1]) ? 0 : kLag;
    const int y_end = min(ph - y_o, bh);
    const int x_end = min(pw - x_o - kLag, (bx + 1 < g.nbw && flat[bidx + 1]) ? bw : bw - kLag);
    if (y_end > y_start && x_end > x_start) {
      nobs += (long long)(y_end - y_start) * (x_end - x_start);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (pi[q] < 0) continue;
        if (!use24 && (pi[q] == 24 || pj[q] == 24)) continue;
        int acc = 0;
        for (int y = y_start; y < y_end; ++y) {
          const int16_t *row = tile + (y + kLag) * kTilePitch + kLag;
          for (int x = x_start; x < x_end; ++x) acc += (int)row[x + off_i[q]] * (int)row[x + off_j[q]];
        }
        acc64[q] += acc;
      }
    }
    __syncthreads();
  }

  unsigned long long *gram = reinterpret_cast<unsigned long long *>(rec + rl.off_gram) + (size_t)c * kPairs;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int p = tid + q * kGramThreads;
    if (p < kPairs && acc64[q] != 0) atomicAdd(&gram[p], (unsigned long long)acc64[q]);
  }
  if (tid == 0 && nobs)
    atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_nobs) + c, (unsigned long long)nobs);
}

void launch_gram_generic(const FrameDesc *frames, int nframes, const Geometry &g, uint8_t *records,
                         const RecordLayout &rl, bool only_overflow, cudaStream_t st) {
  dim3 grid(g.nbh, g.planes, nframes);
  gram_generic_kernel<<<grid, kGramThreads, 0, st>>>(frames, nframes, g, records, rl, only_overflow ? 1 : 0);
}

}  // namespace g1s
