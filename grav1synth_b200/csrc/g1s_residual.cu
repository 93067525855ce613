// residual_kernel — the streaming half of the tensor-core path: one pass over every sample of both
// frames (the only kernel that touches the caller's planes after the flat-block finder).
//
// Replaces, for 4:2:0 and monochrome streams, the per-sample part of av1-grain's diff module reached
// from /root/reference/src/main.rs:442 (DiffGenerator::diff_frame):
//   util.rs::frame_into_u8            truncating >> (bd-8) of both frames, never materialised
//   source - denoised                 the residual every tap of extract_ar_row reads, written once as
//                                     an s8 plane per channel (engine-owned, pitch a multiple of 16 so
//                                     the Gram kernel's TMA boxes can fetch tiles with hardware zero fill)
//   chroma's luma tap                 sum of the co-sited 2x2 luma residuals as one s8 plane of chroma size (a
//                                     tensor-core operand); where it leaves int8 the chroma blocks are flagged
//   get_block_mean / get_noise_var    per 32x32 block: sum of source luma, sum r, sum r^2 per plane
//   int8 range check                  blocks whose Gram reach contains |r| > 127 are flagged for the exact
//                                     int32 kernel (gram_generic_kernel)
//
// Bound: HBM.  Per frame pair at 3840x2160 10-bit it reads 49.8 MB and writes 14.5 MB; a warp owns a
// 256-sample wide strip of one block row (8 luma blocks, or 16 chroma blocks of one plane) and walks
// it two rows at a time with 128-bit loads (four in flight per lane), so each warp instruction moves a
// contiguous 512-byte segment; block statistics stay in registers until the strip is done (two
// shuffles per block, no atomics).
#include "g1s_kernels.h"

namespace g1s {

namespace {

constexpr int kResThreads = 128;
constexpr int kStripW = 256;  // samples per warp row segment: 32 lanes x 8

// Eight consecutive samples reduced to 8 bit, two per register in 16-bit lanes (even sample low).
// Samples at or beyond `lim` read as zero.
template <int BYTES>
__device__ __forceinline__ void load8_lanes(const uint8_t *__restrict__ row, int x8, int lim, int shift, bool fast,
                                            uint32_t (&v)[4]) {
  if (fast && x8 + 8 <= lim) {
    if (BYTES == 2) {
      const uint4 q = __ldg(reinterpret_cast<const uint4 *>(row + 2 * (size_t)x8));
      v[0] = (q.x >> shift) & 0x00FF00FFu;
      v[1] = (q.y >> shift) & 0x00FF00FFu;
      v[2] = (q.z >> shift) & 0x00FF00FFu;
      v[3] = (q.w >> shift) & 0x00FF00FFu;
    } else {
      const uint2 q = __ldg(reinterpret_cast<const uint2 *>(row + x8));
      v[0] = __byte_perm(q.x, 0u, 0x4140);
      v[1] = __byte_perm(q.x, 0u, 0x4342);
      v[2] = __byte_perm(q.y, 0u, 0x4140);
      v[3] = __byte_perm(q.y, 0u, 0x4342);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t a = 0, b = 0;
      const int xa = x8 + 2 * i, xb = xa + 1;
      if (BYTES == 2) {
        if (xa < lim) a = (reinterpret_cast<const uint16_t *>(row)[xa] >> shift) & 0xFFu;
        if (xb < lim) b = (reinterpret_cast<const uint16_t *>(row)[xb] >> shift) & 0xFFu;
      } else {
        if (xa < lim) a = row[xa];
        if (xb < lim) b = row[xb];
      }
      v[i] = a | (b << 16);
    }
  }
}

// r + 256 per 16-bit lane (never borrows across lanes), the int8 range check folded into `ov`
// (a lane leaves [128, 383] <=> r leaves int8 <=> the lane's high byte of (lane - 128) is non-zero; a
// borrow out of the low lane only happens when that lane is already out of range), and the packed s8 words.
__device__ __forceinline__ void residual8(const uint32_t (&s)[4], const uint32_t (&d)[4], uint32_t (&b)[4],
                                          uint32_t &w0, uint32_t &w1, uint32_t &ov) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    b[i] = (s[i] | 0x01000100u) - d[i];
    ov |= b[i] - 0x00800080u;
  }
  w0 = __byte_perm(b[0], b[1], 0x6420);
  w1 = __byte_perm(b[2], b[3], 0x6420);
}

// Marks every block whose Gram reach (the block plus 3 samples left / right / above) can contain a sample
// of columns [x8, x8 + 8) of block row `by`; conservative (the exact kernel redoes a superset).
__device__ __noinline__ void flag_overflow(uint8_t *ovf, unsigned long long *count, int nbw, int nbh, int c_first,
                                           int c_last, int x8, int log2_bw, int by) {
  const int bx_lo = max((x8 - 3) >> log2_bw, 0), bx_hi = min((x8 + 10) >> log2_bw, nbw - 1);
  for (int yy = by; yy <= min(by + 1, nbh - 1); ++yy)
    for (int bx = bx_lo; bx <= bx_hi; ++bx)
      for (int c = c_first; c <= c_last; ++c) ovf[(size_t)c * nbw * nbh + yy * nbw + bx] = 1;
  atomicAdd(count, 1ull);
}

template <int SB, int DB>
__global__ void __launch_bounds__(kResThreads)
residual_kernel(const FrameDesc *__restrict__ frames, int nframes, Geometry g, ResidualStore rs,
                uint8_t *__restrict__ records, RecordLayout rl, int aligned, int nl, int nc) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (kResThreads / 32) + (threadIdx.x >> 5);
  const bool has_chroma = g.planes == 3;
  const int per_row = nl + (has_chroma ? 2 * nc : 0);
  const int per_frame = per_row * g.nbh;
  if (item >= per_frame * nframes) return;
  const int f = item / per_frame;
  int rem = item - f * per_frame;
  const int by = rem / per_row;
  rem -= by * per_row;
  const FrameDesc &fd = frames[f];  // read straight from global memory: a local copy indexed by plane would live on the stack
  uint8_t *rec = records + (size_t)f * rl.bytes;
  int8_t *store = rs.base + (size_t)f * rs.frame_bytes;
  const bool fast = aligned != 0;
  const int W = g.width, H = g.height, pw = W >> 1, ph = H >> 1;

  if (rem < nl) {
    // ------------------------------------------------------------------ luma strip (8 blocks)
    const int x8 = rem * kStripW + 8 * lane;
    const bool act = x8 < W;  // lanes beyond the frame stay for the shuffles
    const int Y0 = 32 * by, rows = act ? min(32, H - Y0) : 0;
    const uint32_t sstr = fd.src_stride[0], dstr = fd.den_stride[0];
    const uint8_t *sp = static_cast<const uint8_t *>(fd.src[0]) + (size_t)Y0 * sstr;
    const uint8_t *dp = static_cast<const uint8_t *>(fd.den[0]) + (size_t)Y0 * dstr;
    int8_t *out = store + rs.off_res[0] + (size_t)Y0 * rs.pitch_l + x8;
    int8_t *y8 = store + rs.off_y8 + (size_t)Y0 * rs.pitch_l + x8;
    const bool taps = has_chroma && (x8 >> 1) < pw;
    int8_t *tap = store + rs.off_tap + (size_t)(Y0 >> 1) * rs.pitch_c + (x8 >> 1);
    int sum_r = 0;
    unsigned sum_q = 0, sum_l = 0;
    uint32_t ov = 0, ovt = 0;
#pragma unroll 2
    for (int y = 0; y < rows; y += 2) {
      const bool two = y + 1 < rows;
      uint32_t s0[4], d0[4], s1[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0};
      load8_lanes<SB>(sp, x8, W, g.src_shift, fast, s0);
      load8_lanes<DB>(dp, x8, W, g.den_shift, fast, d0);
      if (two) {
        load8_lanes<SB>(sp + sstr, x8, W, g.src_shift, fast, s1);
        load8_lanes<DB>(dp + dstr, x8, W, g.den_shift, fast, d1);
      }
      uint32_t b0[4], b1[4], w00, w01, w10, w11;
      residual8(s0, d0, b0, w00, w01, ov);
      residual8(s1, d1, b1, w10, w11, ov);
      *reinterpret_cast<uint2 *>(out) = make_uint2(w00, w01);
      if (two) *reinterpret_cast<uint2 *>(out + rs.pitch_l) = make_uint2(w10, w11);
      // the 8-bit source luma itself, for the flat-block finder that runs next
      *reinterpret_cast<uint2 *>(y8) = make_uint2(__byte_perm(s0[0], s0[1], 0x6420), __byte_perm(s0[2], s0[3], 0x6420));
      if (two)
        *reinterpret_cast<uint2 *>(y8 + rs.pitch_l) =
            make_uint2(__byte_perm(s1[0], s1[1], 0x6420), __byte_perm(s1[2], s1[3], 0x6420));
      sum_r = __dp4a((int)w00, 0x01010101, sum_r);
      sum_r = __dp4a((int)w01, 0x01010101, sum_r);
      sum_r = __dp4a((int)w10, 0x01010101, sum_r);
      sum_r = __dp4a((int)w11, 0x01010101, sum_r);
      sum_q = (unsigned)__dp4a((int)w00, (int)w00, (int)sum_q);
      sum_q = (unsigned)__dp4a((int)w01, (int)w01, (int)sum_q);
      sum_q = (unsigned)__dp4a((int)w10, (int)w10, (int)sum_q);
      sum_q = (unsigned)__dp4a((int)w11, (int)w11, (int)sum_q);
#pragma unroll
      for (int i = 0; i < 4; ++i) sum_l = __dp2a_lo(s0[i] + s1[i], 0x0101u, sum_l);  // lanes <= 510, no carry
      if (taps && two) {
        // u = l + 1024 per chroma sample (l = sum of the 2x2 luma residuals, |l| <= 1020); 1024 = 0 mod 256, so the
        // s8 value is the low byte of u, and l fits int8 <=> 896 <= u < 1152 <=> (u - 896) >> 8 == 0
        uint32_t u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          u[i] = __dp2a_lo(b0[i] + b1[i], 0x0101u, 0u);
          ovt |= (u[i] - 896u) >> 8;
        }
        *reinterpret_cast<uint32_t *>(tap) =
            __byte_perm(__byte_perm(u[0], u[1], 0x5410), __byte_perm(u[2], u[3], 0x5410), 0x6420);
      }
      sp += 2 * (size_t)sstr;
      dp += 2 * (size_t)dstr;
      out += 2 * (size_t)rs.pitch_l;
      y8 += 2 * (size_t)rs.pitch_l;
      tap += rs.pitch_c;
    }
    // block statistics: four lanes per 32-sample block
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      sum_r += __shfl_xor_sync(0xffffffffu, sum_r, o);
      sum_q += __shfl_xor_sync(0xffffffffu, sum_q, o);
      sum_l += __shfl_xor_sync(0xffffffffu, sum_l, o);
    }
    if (act && (lane & 3) == 0) {
      const int b = by * g.nbw + (x8 >> 5);
      reinterpret_cast<int32_t *>(rec + rl.off_rsum)[b] = sum_r;
      reinterpret_cast<uint32_t *>(rec + rl.off_rsq)[b] = sum_q;
      reinterpret_cast<uint32_t *>(rec + rl.off_luma_sum)[b] = sum_l;
    }
    if (ov & 0xFF00FF00u)
      flag_overflow(rec + rl.off_ovf, reinterpret_cast<unsigned long long *>(rec + rl.off_ovf_count), g.nbw, g.nbh, 0,
                    has_chroma ? 2 : 0, x8, 5, by);  // the luma tap needs exact luma
    else if (ovt)  // only chroma's luma tap left int8: the chroma blocks over these four chroma samples
      flag_overflow(rec + rl.off_ovf, reinterpret_cast<unsigned long long *>(rec + rl.off_ovf_count), g.nbw, g.nbh, 1, 2,
                    x8 >> 1, 4, by);
  } else {
    // ------------------------------------------------------------------ chroma strip (16 blocks of one plane)
    rem -= nl;
    const int c = 1 + rem / nc;
    const int x8 = (rem - (c - 1) * nc) * kStripW + 8 * lane;
    const bool act = x8 < pw;
    const int Y0 = 16 * by, rows = act ? min(16, ph - Y0) : 0;
    const uint32_t sstr = fd.src_stride[c], dstr = fd.den_stride[c];
    const uint8_t *sp = static_cast<const uint8_t *>(fd.src[c]) + (size_t)Y0 * sstr;
    const uint8_t *dp = static_cast<const uint8_t *>(fd.den[c]) + (size_t)Y0 * dstr;
    int8_t *out = store + (c == 1 ? rs.off_res[1] : rs.off_res[2]) + (size_t)Y0 * rs.pitch_c + x8;
    int sum_r = 0;
    unsigned sum_q = 0;
    uint32_t ov = 0;
#pragma unroll 2
    for (int y = 0; y < rows; y += 2) {
      const bool two = y + 1 < rows;
      uint32_t s0[4], d0[4], s1[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0};
      load8_lanes<SB>(sp, x8, pw, g.src_shift, fast, s0);
      load8_lanes<DB>(dp, x8, pw, g.den_shift, fast, d0);
      if (two) {
        load8_lanes<SB>(sp + sstr, x8, pw, g.src_shift, fast, s1);
        load8_lanes<DB>(dp + dstr, x8, pw, g.den_shift, fast, d1);
      }
      uint32_t b0[4], b1[4], w00, w01, w10, w11;
      residual8(s0, d0, b0, w00, w01, ov);
      residual8(s1, d1, b1, w10, w11, ov);
      *reinterpret_cast<uint2 *>(out) = make_uint2(w00, w01);
      if (two) *reinterpret_cast<uint2 *>(out + rs.pitch_c) = make_uint2(w10, w11);
      sum_r = __dp4a((int)w00, 0x01010101, sum_r);
      sum_r = __dp4a((int)w01, 0x01010101, sum_r);
      sum_r = __dp4a((int)w10, 0x01010101, sum_r);
      sum_r = __dp4a((int)w11, 0x01010101, sum_r);
      sum_q = (unsigned)__dp4a((int)w00, (int)w00, (int)sum_q);
      sum_q = (unsigned)__dp4a((int)w01, (int)w01, (int)sum_q);
      sum_q = (unsigned)__dp4a((int)w10, (int)w10, (int)sum_q);
      sum_q = (unsigned)__dp4a((int)w11, (int)w11, (int)sum_q);
      sp += 2 * (size_t)sstr;
      dp += 2 * (size_t)dstr;
      out += 2 * (size_t)rs.pitch_c;
    }
    sum_r += __shfl_xor_sync(0xffffffffu, sum_r, 1);  // two lanes per 16-sample block
    sum_q += __shfl_xor_sync(0xffffffffu, sum_q, 1);
    if (act && rows > 0 && (lane & 1) == 0) {
      const int b = by * g.nbw + (x8 >> 4);
      reinterpret_cast<int32_t *>(rec + rl.off_rsum)[(size_t)c * g.nb + b] = sum_r;
      reinterpret_cast<uint32_t *>(rec + rl.off_rsq)[(size_t)c * g.nb + b] = sum_q;
    }
    if (ov & 0xFF00FF00u)
      flag_overflow(rec + rl.off_ovf, reinterpret_cast<unsigned long long *>(rec + rl.off_ovf_count), g.nbw, g.nbh, c, c,
                    x8, 4, by);
  }
}

}  // namespace

ResidualStore ResidualStore::make(const Geometry &g) {
  ResidualStore r{};
  auto up = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
  const size_t pw = (size_t)(g.width >> 1), ph = (size_t)(g.height >> 1);
  r.pitch_l = (uint32_t)up((size_t)g.width, 16);
  r.pitch_c = (uint32_t)up(pw ? pw : 1, 16);
  size_t o = 0;
  r.off_res[0] = o, o += up((size_t)r.pitch_l * g.height, 256);
  r.off_y8 = o, o += up((size_t)r.pitch_l * g.height, 256);
  r.off_res[1] = r.off_res[2] = r.off_tap = 0;
  if (g.planes == 3) {
    const size_t cb = up((size_t)r.pitch_c * (ph ? ph : 1), 256);
    r.off_res[1] = o, o += cb;
    r.off_res[2] = o, o += cb;
    r.off_tap = o, o += cb;
  }
  r.frame_bytes = o;
  return r;
}

void launch_residual(const FrameDesc *frames, int nframes, const Geometry &g, const ResidualStore &rs,
                     uint8_t *records, const RecordLayout &rl, bool aligned, cudaStream_t st) {
  const int nl = (g.width + kStripW - 1) / kStripW;
  const int nc = g.planes == 3 ? ((g.width >> 1) + kStripW - 1) / kStripW : 0;
  const long long items = (long long)(nl + 2 * nc) * g.nbh * nframes;
  const int warps = kResThreads / 32;
  const int grid = (int)((items + warps - 1) / warps);
  const int al = aligned ? 1 : 0;
#define G1S_LAUNCH(SB, DB) \
  residual_kernel<SB, DB><<<grid, kResThreads, 0, st>>>(frames, nframes, g, rs, records, rl, al, nl, nc)
  if (g.src_bytes == 2 && g.den_bytes == 2)
    G1S_LAUNCH(2, 2);
  else if (g.src_bytes == 2)
    G1S_LAUNCH(2, 1);
  else if (g.den_bytes == 2)
    G1S_LAUNCH(1, 2);
  else
    G1S_LAUNCH(1, 1);
#undef G1S_LAUNCH
}

}  // namespace g1s
