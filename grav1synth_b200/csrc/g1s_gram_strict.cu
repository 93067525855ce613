// gram_reforder_kernel — the AR normal equations accumulated in the REFERENCE's order and rounding
// (g1s_diff_config.gram_order = G1S_GRAM_REF_ORDER, the "strict" mode).
//
// av1-grain's NoiseModel::add_block_observations (reached from /root/reference/src/main.rs:442; libaom
// noise_model.c add_block_observations) does, for every observed pixel in raster order of the flat blocks and
// raster order inside a block,
//     A[i][j] += (buffer[i] * buffer[j]) / (255 * 255);   b[i] += (buffer[i] * val) / (255 * 255);
// in f64.  Every entry of A and b is therefore its own serial chain of correctly rounded additions, up to
// 8.3 M terms long at 4K, and its last bits depend on the order.  The exact-integer Gram of the fast path
// (gram_imma_kernel) lands within 1e-13 of it, which is enough to flip the greedy fit_piecewise of the
// scaling points in ~5 % of streams.  This kernel reproduces the chains bit for bit instead:
//   * one thread per chain (tap pair i <= j, or tap x centre sample for b); A is symmetric and
//     buffer[i]*buffer[j] commutes, so A[j][i] is the same chain and is mirrored by the host;
//   * a chain walks the flat blocks in the reference's order; all chains of a (frame, plane) walk the same
//     pixels, so a CTA stages each block's residual tile (with the 3-sample halo) once in shared memory, as
//     f64, double buffered, and its threads read their two taps from it;
//   * the term: p = buffer[i]*buffer[j] is an exact integer-valued double; p / 65025.0 correctly rounded is
//     q0 = p*y, r = fma(-q0, 65025, p), q = fma(r, y, q0) with y = RN(1/65025): verified exhaustively for
//     every |p| <= 1020^2 (tests/test_host_model.py::test_div65025_sequence); then acc = RN(acc + q).
//   * chroma's luma tap is buffer[24] = (sum of the co-sited luma residuals) / nss with nss a power of two:
//     the chains that contain it are accumulated on the unscaled integer sums and scaled by the host
//     (RN commutes with a power-of-two scale; no overflow or underflow in range).
// Parallelism = chains x frames x planes (about 1000 chains per 4:2:0 frame): the kernel is latency- and
// FP64-pipe-bound (5 f64 operations per term, ~3.6 G terms per 4K frame), not memory-bound.
// Any subsampling; reads the caller's planes directly (no s8 store, no int8 range limit).
#include "g1s_kernels.h"

namespace g1s {

namespace {

constexpr int kStrictThreads = 128;
constexpr int kStrictGroups = (kPairs + kStrictThreads - 1) / kStrictThreads;  // 3
constexpr int kSPitch = 40;                 // >= 32 + 2*3
constexpr int kSRows = kBlock + kLag;       // 35
constexpr int kSElems = kSPitch * kSRows;   // residual tile; the luma-tap tile follows with the same geometry

__device__ __forceinline__ int sample8(const void *base, uint32_t stride, int y, int x, int bytes, int shift) {
  const uint8_t *row = reinterpret_cast<const uint8_t *>(base) + (size_t)y * stride;
  if (bytes == 1) return row[x];
  return (reinterpret_cast<const uint16_t *>(row)[x] >> shift) & 0xFF;  // util.rs::frame_into_u8
}

// Eight consecutive terms of a chain, without the adds: q[k] = RN((a[k]*b[k]) / 65025), every operation rounded to
// nearest, nothing contracted (p is an exact integer-valued double; the three-operation quotient is exact for every
// |p| <= 1020^2: oracle/div65025_check.c).
__device__ __forceinline__ void term_chunk(const double *__restrict__ row, int off_i, int off_j, double (&q)[8]) {
  const double y = 1.0 / 65025.0;
  const double *pa = row + off_i, *pb = row + off_j;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double p = __dmul_rn(pa[k], pb[k]);
    const double q0 = __dmul_rn(p, y);
    const double r = __fma_rn(-q0, 65025.0, p);
    q[k] = __fma_rn(r, y, q0);
  }
}

__global__ void __launch_bounds__(kStrictThreads)
gram_reforder_kernel(const FrameDesc *__restrict__ frames, Geometry g, uint8_t *__restrict__ records, RecordLayout rl) {
  __shared__ double tiles[2][2 * kSElems + 8];  // [buffer][residual | luma tap | slack for a short chunk's surplus reads]

  const int tid = threadIdx.x;
  const int c = blockIdx.y;
  const int f = blockIdx.z;
  const FrameDesc &fd = frames[f];
  uint8_t *rec = records + (size_t)f * rl.bytes;
  const uint8_t *flat = rec + rl.off_flat;

  const int sx = c ? g.ss_x : 0, sy = c ? g.ss_y : 0;
  const int bw = kBlock >> sx, bh = kBlock >> sy;
  const int pw = g.width >> sx, ph = g.height >> sy;                      // loop extents (reference: w >> sub_log2)
  const int sw = (g.width + sx) >> sx, sh = (g.height + sy) >> sy;        // plane storage size
  const void *sp = fd.src[c], *dp = fd.den[c];
  const uint32_t ss = fd.src_stride[c], ds = fd.den_stride[c];
  const bool use24 = c > 0;

  // this thread's chain: pair index -> (i, j), i <= j, row-major over the upper triangle of the 26 taps
  const int pair = blockIdx.x * kStrictThreads + tid;
  int ti = 0, tj = 0;
  bool live = pair < kPairs;
  if (live) {
    int rem = pair;
    while (rem >= kTaps - ti) {
      rem -= kTaps - ti;
      ++ti;
    }
    tj = ti + rem;
    if (ti == 25) live = false;                            // centre x centre is not part of the equations
    if (!use24 && (ti == 24 || tj == 24)) live = false;     // luma has no luma tap
  }
  auto tap_off = [&](int t) -> int {
    if (t < 24) return (t / 7 - 3) * kSPitch + (t % 7 - 3);
    if (t == 24) return kSElems;  // luma-tap tile
    return 0;                      // centre sample
  };
  const int off_i = live ? tap_off(ti) : 0, off_j = live ? tap_off(tj) : 0;

  // stage tile(ty, tx) <-> plane (y_o - 3 + ty, x_o - 3 + tx), and the luma-tap tile of the block
  auto stage = [&](int bidx, double *buf) {
    const int by = bidx / g.nbw, bx = bidx - by * g.nbw;
    const int x_o = bx * bw, y_o = by * bh;
    const int tw = bw + 2 * kLag;
    for (int e = tid; e < (bh + kLag) * tw; e += kStrictThreads) {
      const int ty = e / tw, tx = e - ty * tw;
      const int y = y_o - kLag + ty, x = x_o - kLag + tx;
      int r = 0;
      if (y >= 0 && y < sh && x >= 0 && x < sw)
        r = sample8(sp, ss, y, x, g.src_bytes, g.src_shift) - sample8(dp, ds, y, x, g.den_bytes, g.den_shift);
      buf[ty * kSPitch + tx] = (double)r;
    }
    if (use24) {
      for (int e = tid; e < bh * bw; e += kStrictThreads) {
        const int yy = e / bw, xx = e - yy * bw;
        const int y = y_o + yy, x = x_o + xx;
        int l = 0;
        if (y < ph && x < pw) {
          for (int dy = 0; dy < (1 << sy); ++dy)
            for (int dx = 0; dx < (1 << sx); ++dx) {
              const int ly = (y << sy) + dy, lx = (x << sx) + dx;
              l += sample8(fd.src[0], fd.src_stride[0], ly, lx, g.src_bytes, g.src_shift) -
                   sample8(fd.den[0], fd.den_stride[0], ly, lx, g.den_bytes, g.den_shift);
            }
        }
        buf[kSElems + (yy + kLag) * kSPitch + (xx + kLag)] = (double)l;
      }
    }
  };
  auto next_flat = [&](int from) {
    while (from < g.nb && !flat[from]) ++from;
    return from;
  };

  double acc = 0.0;
  int cur = next_flat(0), b = 0;
  if (cur < g.nb) stage(cur, tiles[0]);
  __syncthreads();
  while (cur < g.nb) {
    const int nxt = next_flat(cur + 1);
    if (nxt < g.nb) stage(nxt, tiles[b ^ 1]);
    const int by = cur / g.nbw, bx = cur - by * g.nbw;
    const int x_o = bx * bw, y_o = by * bh;
    // observation rectangle (add_block_observations)
    const int y_start = (by > 0 && flat[cur - g.nbw]) ? 0 : kLag;
    const int x_start = (bx > 0 && flat[cur - 1]) ? 0 : kLag;
    const int y_end = min(ph - y_o, bh);
    const int x_end = min(pw - x_o - kLag, (bx + 1 < g.nbw && flat[cur + 1]) ? bw : bw - kLag);
    if (live && y_end > y_start && x_end > x_start) {
      // The chain itself is serial (one dependent DADD per term); everything before it is not.  Terms are taken in
      // chunks of 8 along a row: the quotients of chunk k+1 are computed while the adds of chunk k retire, so the
      // f64 pipe sees independent work between two dependent adds.  A row's last chunk may be short: its surplus
      // lanes read on into the tile (valid shared memory, never added).
      const double *t0 = tiles[b] + kLag * kSPitch + kLag;
      const int w = x_end - x_start, cpr = (w + 7) >> 3, nchunks = (y_end - y_start) * cpr;
      // two chunk buffers in ping-pong (no register copies); only a row's last chunk can be short
      double qa[8], qb[8];
      int cy = y_start, cx = 0;  // position of the chunk held in qa
      auto next_pos = [&](int &y, int &x) {
        if (++x == cpr) x = 0, ++y;
      };
      auto add_chunk = [&](const double (&q)[8], int x) {
        const int cnt = w - 8 * x;
        if (cnt >= 8) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc = __dadd_rn(acc, q[k]);
        } else {
#pragma unroll
          for (int k = 0; k < 7; ++k)
            if (k < cnt) acc = __dadd_rn(acc, q[k]);
        }
      };
      term_chunk(t0 + cy * kSPitch + x_start, off_i, off_j, qa);
      int ci = 0;
#pragma unroll 1
      for (; ci + 2 <= nchunks; ci += 2) {
        int y1 = cy, x1 = cx;
        next_pos(y1, x1);
        term_chunk(t0 + y1 * kSPitch + x_start + 8 * x1, off_i, off_j, qb);
        add_chunk(qa, cx);
        int y2 = y1, x2 = x1;
        next_pos(y2, x2);
        if (ci + 2 < nchunks) term_chunk(t0 + y2 * kSPitch + x_start + 8 * x2, off_i, off_j, qa);
        add_chunk(qb, x1);
        cy = y2, cx = x2;
      }
      if (ci < nchunks) add_chunk(qa, cx);
    }
    __syncthreads();  // tile b is free again, tile b^1 is complete
    cur = nxt;
    b ^= 1;
  }
  if (live) reinterpret_cast<double *>(rec + rl.off_gramf)[(size_t)c * kPairs + pair] = acc;
}

}  // namespace

void launch_gram_strict(const FrameDesc *frames, int nframes, const Geometry &g, uint8_t *records, const RecordLayout &rl,
                        cudaStream_t st) {
  dim3 grid(kStrictGroups, g.planes, nframes);
  gram_reforder_kernel<<<grid, kStrictThreads, 0, st>>>(frames, g, records, rl);
}

}  // namespace g1s
