// gram_imma_kernel — the fused residual + AR normal-equation (Gram) kernel on the int8
// tensor-core path (mma.sync.m16n8k32.s8, SASS IMMA.16832.S8.S8), 4:2:0 and monochrome.
//
// Replaces NoiseModel::add_block_observations + get_block_mean + get_noise_var of av1-grain's
// diff module (reached from /root/reference/src/main.rs:442) for every flat block whose
// residuals fit in int8; the rare block that does not is flagged and redone exactly by
// gram_generic_kernel.  All sums are integers, so the result is bit-identical to the oracle
// whatever the summation order.
//
// Work unit ("super-unit"): two horizontally adjacent 32x32 luma blocks plus their co-sited
// 16x16 Cb and Cr blocks, so every sample of both frames is read from HBM once here: the luma
// residual tile also provides the chroma "luma tap" (sum of the co-sited 2x2 luma residuals).
// A CTA of 6 warps walks a run of super-units of one block row:
//   staging  all warps: 64-bit loads of source and denoised, >> (bd-8), subtract, pack to s8
//            words in shared memory (tile origin 4 samples left of the unit so loads are aligned),
//            per-block sum r / sum r^2 / sum luma while the values are in registers;
//   Gram     warps 0-3: the two luma blocks (half the rows each); warp 4: Cb pair; warp 5: Cb pair.
//            One k-step = 32 pixels of one row.  X[k][a] = residual at pixel k shifted by tap a;
//            D += X^T X over the upper block-triangle (6 m16n8k32 MMAs).  Tap a = 8q+g with
//            g = cx+3 (lane group), q = cy+3: a thread's four taps are the SAME column offset on
//            four consecutive rows, so its operand words slide down by one row per k-step and
//            only one new 32-bit window (two LDS + funnel shift) per half is fetched.
//            The observation mask (block margins, frame clipping) is a byte mask on k applied
//            to the operand words (mask^2 = mask, so masking both A and B is exact).
//   epilogue int32 accumulators (bounded: <= 32 units * 16 k-steps * 32 * 2^14 * 4 warps < 2^31)
//            -> int64 global atomics, one per tap pair per CTA per plane.
#include "g1s_kernels.h"

namespace g1s {

namespace {

constexpr int kSuThreads = 192;
constexpr int kSuRun = 30;        // super-units per CTA (bounds the int32 accumulators, see above)
constexpr int kPL = 20;           // luma tile pitch in 32-bit words (72 bytes used)
constexpr int kLumaRows = 35;     // 3 halo rows + 32
constexpr int kPC = 12;           // chroma tile pitch in words (40 bytes used)
constexpr int kChromaRows = 19;   // 3 halo rows + 16

struct __align__(16) SuSmem {
  uint32_t luma[kLumaRows * kPL];
  uint32_t chroma[2][kChromaRows * kPC];
  uint32_t hs[16 * 8];   // luma tap >> 3, one s8 per chroma pixel of the unit (32 x 16)
  uint32_t ls[16 * 8];   // luma tap & 7
  int dl[6 * 4 * 32];    // luma accumulators reduced over warps 0-3
  int st_rs[6];
  unsigned st_rq[6];
  unsigned st_ls[2];
  int ovf[3];
};

__device__ __forceinline__ void imma_16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                           uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Four consecutive samples reduced to 8 bit (util.rs::frame_into_u8), zero outside [0,lim_w)x[0,lim_h).
template <bool ALIGNED>
__device__ __forceinline__ void load4(const void *base, uint32_t stride, int y, int x, int bytes, int shift,
                                      int lim_w, int lim_h, int (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0;
  if (y < 0 || y >= lim_h || x < 0 || x >= lim_w) return;
  const uint8_t *row = reinterpret_cast<const uint8_t *>(base) + (size_t)y * stride;
  if (ALIGNED && x + 3 < lim_w) {
    if (bytes == 2) {
      const uint2 p = __ldg(reinterpret_cast<const uint2 *>(row + 2 * x));
      v[0] = ((p.x & 0xFFFFu) >> shift) & 0xFF;
      v[1] = ((p.x >> 16) >> shift) & 0xFF;
      v[2] = ((p.y & 0xFFFFu) >> shift) & 0xFF;
      v[3] = ((p.y >> 16) >> shift) & 0xFF;
    } else {
      const uint32_t p = __ldg(reinterpret_cast<const uint32_t *>(row + x));
      v[0] = p & 0xFF;
      v[1] = (p >> 8) & 0xFF;
      v[2] = (p >> 16) & 0xFF;
      v[3] = p >> 24;
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (x + i < lim_w) {
      v[i] = bytes == 2 ? ((reinterpret_cast<const uint16_t *>(row)[x + i] >> shift) & 0xFF) : row[x + i];
    }
  }
}

// Residual word for 4 samples: returns packed s8, accumulates statistics, reports overflow.
template <bool ALIGNED>
__device__ __forceinline__ uint32_t residual_word(const void *sp, uint32_t ss, const void *dp, uint32_t ds, int y,
                                                  int x, const Geometry &g, int lim_w, int lim_h, int &rs,
                                                  unsigned &rq, unsigned &ls, bool &ovf) {
  int s[4], d[4];
  load4<ALIGNED>(sp, ss, y, x, g.src_bytes, g.src_shift, lim_w, lim_h, s);
  load4<ALIGNED>(dp, ds, y, x, g.den_bytes, g.den_shift, lim_w, lim_h, d);
  uint32_t w = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = s[i] - d[i];
    ovf |= (r < -128) | (r > 127);
    rs += r;
    rq += (unsigned)(r * r);
    ls += (unsigned)s[i];
    w |= (uint32_t)(r & 0xFF) << (8 * i);
  }
  return w;
}

__device__ __forceinline__ uint32_t byte_mask(int first, int lo, int hi) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (first + i >= lo && first + i < hi) m |= 0xFFu << (8 * i);
  return m;
}

// The k-loop of one warp over rows [ys, ye) of a 32-pixel-wide unit.
//   tile/pitch : s8 residual tile, row r <-> plane row (unit origin - 3 + r), col 0 <-> origin - 4
//   colw       : word column of this lane's window for half 0 ( = unit word base + t + ((g+1)>>2) )
//   sh         : funnel shift in bits ( = 8 * ((g+1) & 3) )
//   mx[h]      : byte mask of the observed pixels of half h (x margins, frame clip, block not flat)
//   y0h[h]     : first observed row of half h (rows below it are masked at use; chroma pairs only)
template <bool CHROMA>
__device__ __forceinline__ void gram_rows(const uint32_t *__restrict__ tile, int pitch, int colw, int sh, int ys,
                                          int ye, const uint32_t (&mx)[2], const int (&y0h)[2],
                                          const uint32_t *__restrict__ special, bool is_special, int t,
                                          int (&acc)[6][4]) {
  uint32_t W[4][2];
  auto window = [&](int row, int h) -> uint32_t {
    const uint32_t *p = tile + row * pitch + colw + 4 * h;
    return __funnelshift_r(p[0], p[1], sh) & mx[h];
  };
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    W[q][0] = window(ys + q, 0);
    W[q][1] = window(ys + q, 1);
  }
#pragma unroll 4
  for (int y = ys; y < ye; ++y) {
    W[3][0] = window(y + 3, 0);
    W[3][1] = window(y + 3, 1);
    uint32_t u3[2] = {W[3][0], W[3][1]};
    if (CHROMA) {
      // lanes g = 4 / 5 carry the luma-tap hi / lo bytes in their (otherwise unused) cy = 0 slot
      const uint32_t x0 = special[y * 8 + t] & mx[0];
      const uint32_t x1 = special[y * 8 + 4 + t] & mx[1];
      u3[0] = is_special ? x0 : u3[0];
      u3[1] = is_special ? x1 : u3[1];
    }
    uint32_t A[4][2];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      A[q][0] = W[q][0];
      A[q][1] = W[q][1];
    }
    A[3][0] = u3[0];
    A[3][1] = u3[1];
    if (CHROMA) {
      // rows above a block's first observed row contribute nothing for that half
      const uint32_t r0 = y >= y0h[0] ? 0xFFFFFFFFu : 0u, r1 = y >= y0h[1] ? 0xFFFFFFFFu : 0u;
      uint32_t B[4][2];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        B[q][0] = A[q][0] & r0;
        B[q][1] = A[q][1] & r1;
      }
      imma_16832(acc[0], A[0][0], A[1][0], A[0][1], A[1][1], B[0][0], B[0][1]);
      imma_16832(acc[1], A[0][0], A[1][0], A[0][1], A[1][1], B[1][0], B[1][1]);
      imma_16832(acc[2], A[0][0], A[1][0], A[0][1], A[1][1], B[2][0], B[2][1]);
      imma_16832(acc[3], A[0][0], A[1][0], A[0][1], A[1][1], B[3][0], B[3][1]);
      imma_16832(acc[4], A[2][0], A[3][0], A[2][1], A[3][1], B[2][0], B[2][1]);
      imma_16832(acc[5], A[2][0], A[3][0], A[2][1], A[3][1], B[3][0], B[3][1]);
    } else {
      imma_16832(acc[0], A[0][0], A[1][0], A[0][1], A[1][1], A[0][0], A[0][1]);
      imma_16832(acc[1], A[0][0], A[1][0], A[0][1], A[1][1], A[1][0], A[1][1]);
      imma_16832(acc[2], A[0][0], A[1][0], A[0][1], A[1][1], A[2][0], A[2][1]);
      imma_16832(acc[3], A[0][0], A[1][0], A[0][1], A[1][1], A[3][0], A[3][1]);
      imma_16832(acc[4], A[2][0], A[3][0], A[2][1], A[3][1], A[2][0], A[2][1]);
      imma_16832(acc[5], A[2][0], A[3][0], A[2][1], A[3][1], A[3][0], A[3][1]);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      W[q][0] = W[q + 1][0];
      W[q][1] = W[q + 1][1];
    }
  }
}

// MMA tap index a = 8q+g  ->  record tap index (0..23 AR taps, 25 centre), -1 unused, -2/-3 luma-tap hi/lo.
__device__ __forceinline__ int record_tap(int a) {
  const int q = a >> 3, g = a & 7;
  if (g == 7) return -1;
  if (q < 3) return 7 * q + g;
  if (g < 3) return 21 + g;
  if (g == 3) return 25;
  if (g == 4) return -2;
  if (g == 5) return -3;
  return -1;
}

__device__ __forceinline__ int pair_index(int i, int j) { return i * kTaps - i * (i - 1) / 2 + (j - i); }

// Adds one accumulator element D[a][b] into the plane's 351-entry int64 Gram (global atomics).
__device__ __forceinline__ void emit(unsigned long long *gram, int a, int b, long long v, bool chroma) {
  if (a > b || v == 0) return;  // the mirrored element is always covered by another tile
  int ia = record_tap(a), ib = record_tap(b);
  if (ia == -1 || ib == -1) return;
  long long wgt = 1;
  if (ia < -1 || ib < -1) {
    if (!chroma) return;
    if (ia < -1) {
      wgt *= ia == -2 ? 8 : 1;
      ia = 24;
    }
    if (ib < -1) {
      wgt *= ib == -2 ? 8 : 1;
      ib = 24;
    }
    if (a != b && ia == 24 && ib == 24) wgt *= 2;  // (hi,lo) appears once (a<b) but counts twice in (8h+l)^2
  }
  const int i = min(ia, ib), j = max(ia, ib);
  atomicAdd(&gram[pair_index(i, j)], (unsigned long long)(v * wgt));
}

template <bool ALIGNED>
__global__ void __launch_bounds__(kSuThreads)
gram_imma_kernel(const FrameDesc *__restrict__ frames, Geometry g, uint8_t *__restrict__ records, RecordLayout rl,
                 int runs_per_row) {
  __shared__ SuSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;  // mma "groupID" (= cx + 3) and thread-in-group
  const int by = blockIdx.x / runs_per_row;
  const int run = blockIdx.x - by * runs_per_row;
  const int f = blockIdx.y;
  const FrameDesc fd = frames[f];
  uint8_t *rec = records + (size_t)f * rl.bytes;
  const uint8_t *flat = rec + rl.off_flat;
  uint8_t *ovf_out = rec + rl.off_ovf;
  const bool has_chroma = g.planes == 3;
  const int W = g.width, H = g.height, pw = W >> 1, ph = H >> 1;
  const int nsu = (g.nbw + 1) >> 1;
  const int u_beg = run * kSuRun, u_end = min(nsu, u_beg + kSuRun);

  int acc[6][4];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[i][r] = 0;
  long long nobs[3] = {0, 0, 0};  // thread 0 only

  for (int i = tid; i < (int)(sizeof(SuSmem) / 4); i += kSuThreads) reinterpret_cast<uint32_t *>(&sm)[i] = 0;
  __syncthreads();

  const int sh = 8 * ((gq + 1) & 3);
  const int dxw = (gq + 1) >> 2;

  for (int u = u_beg; u < u_end; ++u) {
    const int bx0 = 2 * u;
    const bool ex1 = bx0 + 1 < g.nbw;
    const int b0 = by * g.nbw + bx0;
    const bool fl0 = flat[b0] != 0, fl1 = ex1 && flat[b0 + 1] != 0;
    if (!fl0 && !fl1) continue;  // uniform across the CTA

    // ------------------------------------------------------------------ staging
    if (tid < 6) {
      sm.st_rs[tid] = 0;
      sm.st_rq[tid] = 0;
      if (tid < 2) sm.st_ls[tid] = 0;
      if (tid < 3) sm.ovf[tid] = 0;
    }
    __syncthreads();  // previous unit's Gram loops are done with the tiles; stats zeroed
    const int X0 = 64 * u, Y0 = 32 * by, CX0 = 32 * u, CY0 = 16 * by;
    {
      // luma main words: two tile rows per warp pass, lanes 0-15 / 16-31, word w = 1 + (lane & 15)
      for (int p = warp; p < 18; p += 6) {
        const int ty = 2 * p + (lane >> 4), w = 1 + (lane & 15);
        int rs = 0;
        unsigned rq = 0, ls = 0;
        bool ov = false;
        if (ty < kLumaRows) {
          const uint32_t word = residual_word<ALIGNED>(fd.src[0], fd.src_stride[0], fd.den[0], fd.den_stride[0],
                                                       Y0 - 3 + ty, X0 - 4 + 4 * w, g, W, H, rs, rq, ls, ov);
          sm.luma[ty * kPL + w] = word;
          if (ty < 3) rs = 0, rq = 0, ls = 0;  // halo rows belong to the block above
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          rs += __shfl_xor_sync(0xffffffffu, rs, o);
          rq += __shfl_xor_sync(0xffffffffu, rq, o);
          ls += __shfl_xor_sync(0xffffffffu, ls, o);
        }
        if ((lane & 7) == 0) {
          const int blk = (lane >> 3) & 1;
          atomicAdd(&sm.st_rs[blk], rs);
          atomicAdd(&sm.st_rq[blk], rq);
          atomicAdd(&sm.st_ls[blk], ls);
        }
        if (__any_sync(0xffffffffu, ov) && lane == 0) sm.ovf[0] = 1;
      }
      // luma halo words 0 and 17 (70 items), chroma halo words 0 and 9 (2 planes x 38 items)
      if (tid < 70) {
        const int ty = tid >> 1, w = (tid & 1) * 17;
        int rs = 0;
        unsigned rq = 0, ls = 0;
        bool ov = false;
        sm.luma[ty * kPL + w] = residual_word<ALIGNED>(fd.src[0], fd.src_stride[0], fd.den[0], fd.den_stride[0],
                                                       Y0 - 3 + ty, X0 - 4 + 4 * w, g, W, H, rs, rq, ls, ov);
        if (ov) sm.ovf[0] = 1;
      } else if (has_chroma && tid < 70 + 76) {
        const int idx = tid - 70, c = idx / 38, rem = idx - c * 38;
        const int ty = rem >> 1, w = (rem & 1) * 9;
        int rs = 0;
        unsigned rq = 0, ls = 0;
        bool ov = false;
        sm.chroma[c][ty * kPC + w] =
            residual_word<ALIGNED>(fd.src[1 + c], fd.src_stride[1 + c], fd.den[1 + c], fd.den_stride[1 + c],
                                   CY0 - 3 + ty, CX0 - 4 + 4 * w, g, pw, ph, rs, rq, ls, ov);
        if (ov) sm.ovf[1 + c] = 1;
      }
      // chroma main words: four tile rows per warp pass, word w = 1 + (lane & 7)
      if (has_chroma) {
        for (int pp = warp; pp < 10; pp += 6) {
          const int c = pp / 5, p = pp - 5 * c;
          const int ty = 4 * p + (lane >> 3), w = 1 + (lane & 7);
          int rs = 0;
          unsigned rq = 0, ls = 0;
          bool ov = false;
          if (ty < kChromaRows) {
            const uint32_t word =
                residual_word<ALIGNED>(fd.src[1 + c], fd.src_stride[1 + c], fd.den[1 + c], fd.den_stride[1 + c],
                                       CY0 - 3 + ty, CX0 - 4 + 4 * w, g, pw, ph, rs, rq, ls, ov);
            sm.chroma[c][ty * kPC + w] = word;
            if (ty < 3) rs = 0, rq = 0;
          }
#pragma unroll
          for (int o = 1; o < 4; o <<= 1) {
            rs += __shfl_xor_sync(0xffffffffu, rs, o);
            rq += __shfl_xor_sync(0xffffffffu, rq, o);
          }
          if ((lane & 3) == 0) {
            const int blk = (lane >> 2) & 1;
            atomicAdd(&sm.st_rs[2 + 2 * c + blk], rs);
            atomicAdd(&sm.st_rq[2 + 2 * c + blk], rq);
          }
          if (__any_sync(0xffffffffu, ov) && lane == 0) sm.ovf[1 + c] = 1;
        }
      }
    }
    __syncthreads();  // tiles + statistics + overflow flags complete

    // luma tap of the chroma planes: sum of the co-sited 2x2 luma residuals = 8*hi + lo
    if (has_chroma && tid < 128) {
      const int cy = tid >> 3, w = tid & 7;
      const uint32_t *r0 = &sm.luma[(3 + 2 * cy) * kPL + 1 + 2 * w];
      const uint32_t *r1 = r0 + kPL;
      uint32_t hw = 0, lw = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t a = (i < 2 ? r0[0] : r0[1]) >> (16 * (i & 1));
        const uint32_t b = (i < 2 ? r1[0] : r1[1]) >> (16 * (i & 1));
        const int l4 = (int)(int8_t)(a & 0xFF) + (int)(int8_t)((a >> 8) & 0xFF) + (int)(int8_t)(b & 0xFF) +
                       (int)(int8_t)((b >> 8) & 0xFF);
        hw |= (uint32_t)((l4 >> 3) & 0xFF) << (8 * i);
        lw |= (uint32_t)(l4 & 7) << (8 * i);
      }
      sm.hs[cy * 8 + w] = hw;
      sm.ls[cy * 8 + w] = lw;
    }
    // statistics and overflow flags out (each block belongs to exactly one CTA)
    const bool ovl = sm.ovf[0] != 0;
    const bool ovc[2] = {ovl || sm.ovf[1] != 0, ovl || sm.ovf[2] != 0};  // the luma tap needs an exact luma tile
    if (tid < 6) {
      const int c = tid >> 1, blk = tid & 1;
      const bool fl = blk ? fl1 : fl0;
      if (fl && (c == 0 || has_chroma)) {
        reinterpret_cast<int32_t *>(rec + rl.off_rsum)[c * g.nb + b0 + blk] = sm.st_rs[tid];
        reinterpret_cast<uint32_t *>(rec + rl.off_rsq)[c * g.nb + b0 + blk] = sm.st_rq[tid];
        if (c == 0) reinterpret_cast<uint32_t *>(rec + rl.off_luma_sum)[b0 + blk] = sm.st_ls[blk];
        const bool o = c == 0 ? ovl : ovc[c - 1];
        if (o) {
          ovf_out[(size_t)c * g.nb + b0 + blk] = 1;
          atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_ovf_count), 1ull);
        }
      }
    }
    __syncthreads();  // hs / ls visible

    // ------------------------------------------------------------------ observation rectangles
    // add_block_observations: margins of 3 unless the neighbour block is flat too
    const bool up0 = by > 0 && flat[b0 - g.nbw], up1 = by > 0 && ex1 && flat[b0 + 1 - g.nbw];
    const bool lf0 = bx0 > 0 && flat[b0 - 1], rt1 = bx0 + 2 < g.nbw && flat[b0 + 2];
    int xs[2], ys0[2];
    xs[0] = lf0 ? 0 : kLag;
    xs[1] = fl0 ? 0 : kLag;
    ys0[0] = up0 ? 0 : kLag;
    ys0[1] = up1 ? 0 : kLag;
    const bool rn[2] = {fl1, rt1};  // right neighbour flat

    if (warp < 4) {
      const int j = warp >> 1;
      const bool fl = j ? fl1 : fl0;
      if (fl && !ovl) {
        const int x_o = 32 * (bx0 + j);
        const int x1 = min(W - x_o - kLag, rn[j] ? 32 : 32 - kLag);
        const int y1 = min(H - Y0, 32);
        const int y0 = ys0[j];
        if (x1 > xs[j] && y1 > y0) {
          if (tid == 0 || tid == 64) nobs[0] += (long long)(x1 - xs[j]) * (y1 - y0);
          const int mid = y0 + ((y1 - y0 + 1) >> 1);
          const int ys = (warp & 1) ? mid : y0, ye = (warp & 1) ? y1 : mid;
          const uint32_t mx[2] = {byte_mask(4 * t, xs[j], x1), byte_mask(16 + 4 * t, xs[j], x1)};
          const int y0h[2] = {0, 0};
          gram_rows<false>(sm.luma, kPL, 8 * j + t + dxw, sh, ys, ye, mx, y0h, nullptr, false, t, acc);
        }
      }
    } else if (has_chroma) {
      const int c = warp - 4;
      const int y1 = min(ph - CY0, 16);
      uint32_t mx[2] = {0, 0};
      int y0h[2] = {99, 99};
      int ymin = 99;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const bool fl = j ? fl1 : fl0;
        if (!fl || ovc[c]) continue;
        const int x_o = 16 * (bx0 + j);
        const int x1 = min(pw - x_o - kLag, rn[j] ? 16 : 16 - kLag);
        if (x1 > xs[j] && y1 > ys0[j]) {
          mx[j] = byte_mask(4 * t, xs[j], x1);
          y0h[j] = ys0[j];
          ymin = min(ymin, ys0[j]);
          if (lane == 0) nobs[1 + c] += (long long)(x1 - xs[j]) * (y1 - ys0[j]);
        }
      }
      if (ymin < y1)
        gram_rows<true>(sm.chroma[c], kPC, t + dxw, sh, ymin, y1, mx, y0h, gq == 4 ? sm.hs : sm.ls,
                        gq == 4 || gq == 5, t, acc);
    }
  }

  // ---------------------------------------------------------------------- epilogue
  __syncthreads();
  if (warp < 4) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (acc[i][r]) atomicAdd(&sm.dl[(i * 4 + r) * 32 + lane], acc[i][r]);
  }
  __syncthreads();
  if (warp == 0 || warp >= 4) {
    const int plane = warp == 0 ? 0 : warp - 3;
    if (plane == 0 || has_chroma) {
      unsigned long long *gram = reinterpret_cast<unsigned long long *>(rec + rl.off_gram) + (size_t)plane * kPairs;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int mrow = i >= 4 ? 16 : 0;
        const int ncol = i >= 4 ? 8 * (i - 2) : 8 * i;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int v = warp == 0 ? sm.dl[(i * 4 + r) * 32 + lane] : acc[i][r];
          emit(gram, mrow + gq + 8 * (r >> 1), ncol + 2 * t + (r & 1), (long long)v, plane > 0);
        }
      }
    }
  }
  // observation counts: thread 0 (luma block 0), thread 64 (luma block 1), lanes 0 of warps 4 / 5
  if (nobs[0]) atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_nobs), (unsigned long long)nobs[0]);
  if (nobs[1]) atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_nobs) + 1, (unsigned long long)nobs[1]);
  if (nobs[2]) atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_nobs) + 2, (unsigned long long)nobs[2]);
}

}  // namespace

bool gram_imma_supported(const Geometry &g) {
  return (g.planes == 1) || (g.planes == 3 && g.ss_x == 1 && g.ss_y == 1);
}

void launch_gram_imma(const FrameDesc *frames, int nframes, const Geometry &g, uint8_t *records,
                      const RecordLayout &rl, bool aligned, cudaStream_t st) {
  const int nsu = (g.nbw + 1) / 2;
  const int runs = (nsu + kSuRun - 1) / kSuRun;
  dim3 grid(runs * g.nbh, nframes);
  if (aligned)
    gram_imma_kernel<true><<<grid, kSuThreads, 0, st>>>(frames, g, records, rl, runs);
  else
    gram_imma_kernel<false><<<grid, kSuThreads, 0, st>>>(frames, g, records, rl, runs);
}

}  // namespace g1s
