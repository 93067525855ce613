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
//   staging  all warps: 64-bit loads of source and denoised, >> (bd-8) and subtract two samples
//            per 32-bit op (16-bit SIMD lanes), pack to s8 words in shared memory (tile origin
//            4 samples left of the unit so loads are aligned); per-block sum r / sum r^2 /
//            sum luma by dp4a while the values are in registers;
//   Gram     warps 0-3: the two luma blocks (half the rows each); warp 4: Cb pair; warp 5: Cr pair.
//            One k-step = 32 pixels of one row.  X[k][a] = residual at pixel k shifted by tap a;
//            D += X^T X over the upper block-triangle (6 m16n8k32 MMAs).  Tap a = 8q+g with
//            g = cx+3 (lane group), q = cy+3: a thread's four taps are the SAME column offset on
//            four consecutive rows, so its operands slide down one row per k-step: one new
//            32-bit window (two LDS + funnel shift) per half is fetched, the B operand of a row
//            is a register pair and the A operand of two consecutive rows a register quad that
//            serves as the lower m-tile now and as the upper m-tile two steps later.
//            The observation mask (block margins, frame clipping) is a byte mask on k applied
//            to the operand words (mask^2 = mask, so masking both A and B is exact).
//            Chroma adds one n-tile whose columns 0/1 are the luma tap split as 8*hi + lo.
//   epilogue int32 accumulators (bounded: <= 12 units * 16 k-steps * 32 * 2^14 * 4 warps < 2^31)
//            -> int64 global atomics, one per tap pair per CTA per plane.
#include "g1s_kernels.h"

namespace g1s {

namespace {

constexpr int kSuThreads = 192;
constexpr int kSuRun = 12;        // super-units per CTA (bounds the int32 accumulators, see above)
constexpr int kPL = 20;           // luma tile pitch in 32-bit words (72 bytes used)
constexpr int kLumaRows = 35;     // 3 halo rows + 32
constexpr int kPC = 12;           // chroma tile pitch in words (40 bytes used)
constexpr int kChromaRows = 19;   // 3 halo rows + 16

constexpr int kFlagCols = 2 * kSuRun + 2;  // flat flags of the run's blocks plus one neighbour each side

// Everything one super-unit needs to know about its observation rectangles (add_block_observations:
// margins of 3 unless the neighbouring block is flat too), computed once per unit by one thread.
struct UnitInfo {
  int xs0, xs1;      // first observed column of block 0 / 1 (block-local)
  int y00, y01;      // first observed row of block 0 / 1
  int x1l0, x1l1;    // luma: one past the last observed column of block 0 / 1
  int x1c0, x1c1;    // chroma: the same in chroma samples
  int y1l, y1c;      // one past the last observed row (frame clip), luma / chroma
};

struct __align__(16) SuSmem {
  uint32_t luma[kLumaRows * kPL];
  uint32_t chroma[2][kChromaRows * kPC];
  uint32_t hs[kChromaRows * kPC];        // luma tap >> 3 (s8 per chroma pixel), chroma-tile pitch, rows >= 16 stay 0
  uint32_t ls[(kChromaRows + 1) * kPC];  // luma tap & 7, stored one row down (row -1 is readable and 0)
  int dl[6 * 4 * 32];                    // luma accumulators reduced over warps 0-3
  int st_rs[6];
  unsigned st_rq[6];
  unsigned st_ls[2];
  int ovf[3];
  int self[2][3];                        // per chroma plane: sum h*h, h*l, l*l over observed pixels
  UnitInfo info;
  uint8_t flat[2][kFlagCols + 2];        // [0] this block row, [1] the row above; column 0 <-> block 2*u_beg - 1
};

__device__ __forceinline__ void imma_16832(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ------------------------------------------------------------------------------ staging
//
// Fast path (interior unit, 8-byte aligned rows): four samples of each plane arrive in one
// vector load and are reduced to 8 bit two at a time in 16-bit lanes.

// Four consecutive samples as two words of two 16-bit lanes each, values 0..255
// (util.rs::frame_into_u8: truncating shift, then `as u8`).
template <int BYTES, bool SMEM>
__device__ __forceinline__ void load4_lanes(const uint8_t *p, int shift, uint32_t &lo2, uint32_t &hi2) {
  if (BYTES == 2) {
    const uint2 v = SMEM ? *reinterpret_cast<const uint2 *>(p) : __ldg(reinterpret_cast<const uint2 *>(p));
    lo2 = (v.x >> shift) & 0x00FF00FFu;
    hi2 = (v.y >> shift) & 0x00FF00FFu;
  } else {
    const uint32_t v = SMEM ? *reinterpret_cast<const uint32_t *>(p) : __ldg(reinterpret_cast<const uint32_t *>(p));
    lo2 = __byte_perm(v, 0u, 0x4140);
    hi2 = __byte_perm(v, 0u, 0x4342);
  }
}

// Packed s8 residual word of four samples + running statistics.  The statistics use the s8 word,
// so they are exact only when no sample overflowed; overflowed blocks are redone (statistics
// included) by the generic kernel.
template <int SB, int DB, bool STATS, bool LUMA_SUM, bool SMEM = false>
__device__ __forceinline__ uint32_t residual4(const uint8_t *sp, const uint8_t *dp, int sshift, int dshift, int &rs,
                                              int &rq, unsigned &ls, uint32_t &ovf) {
  uint32_t s0, s1, d0, d1;
  load4_lanes<SB, SMEM>(sp, sshift, s0, s1);
  load4_lanes<DB, SMEM>(dp, dshift, d0, d1);
  const uint32_t b0 = (s0 | 0x01000100u) - d0;  // per lane: r + 256, never borrows across lanes
  const uint32_t b1 = (s1 | 0x01000100u) - d1;
  ovf |= ((b0 - 0x00800080u) | (b1 - 0x00800080u)) & 0xFF00FF00u;  // lane outside [128, 383] <=> r outside int8
  const uint32_t w = __byte_perm(b0, b1, 0x6420);
  if (STATS) {
    rs = __dp4a((int)w, 0x01010101, rs);
    rq = __dp4a((int)w, (int)w, rq);
    if (LUMA_SUM) ls = __dp4a(__byte_perm(s0, s1, 0x6420), 0x01010101u, ls);
  }
  return w;
}

// Slow path (frame edges, unaligned rows): scalar, bounds-checked, zero outside [0,lim_w)x[0,lim_h).
__device__ __forceinline__ uint32_t residual4_slow(const void *sp, uint32_t ss, const void *dp, uint32_t ds, int y,
                                                   int x, const Geometry &g, int lim_w, int lim_h, bool stats,
                                                   int &rs, int &rq, unsigned &ls, uint32_t &ovf) {
  uint32_t w = 0;
  if (y < 0 || y >= lim_h) return 0;
  const uint8_t *srow = reinterpret_cast<const uint8_t *>(sp) + (size_t)y * ss;
  const uint8_t *drow = reinterpret_cast<const uint8_t *>(dp) + (size_t)y * ds;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int xx = x + i;
    if (xx < 0 || xx >= lim_w) continue;
    const int s = g.src_bytes == 2 ? ((reinterpret_cast<const uint16_t *>(srow)[xx] >> g.src_shift) & 0xFF) : srow[xx];
    const int d = g.den_bytes == 2 ? ((reinterpret_cast<const uint16_t *>(drow)[xx] >> g.den_shift) & 0xFF) : drow[xx];
    const int r = s - d;
    if (r < -128 || r > 127) ovf |= 1u;
    if (stats) {
      rs += r;
      rq += r * r;
      ls += (unsigned)s;
    }
    w |= (uint32_t)(r & 0xFF) << (8 * i);
  }
  return w;
}

// Byte i of the result is 0xFF iff lo <= first + i < hi (i = 0..3).
__device__ __forceinline__ uint32_t byte_mask(int first, int lo, int hi) {
  const int a = min(max(lo - first, 0), 4), b = min(max(hi - first, 0), 4);  // bytes [a, b) are on
  const uint32_t below_b = b >= 4 ? 0xFFFFFFFFu : ((1u << (8 * b)) - 1u);
  const uint32_t below_a = a >= 4 ? 0xFFFFFFFFu : ((1u << (8 * a)) - 1u);
  return b > a ? (below_b & ~below_a) : 0u;
}

// ------------------------------------------------------------------------------ k-loops
//
//   tile/pitch : s8 residual tile, row r <-> plane row (unit origin - 3 + r), col 0 <-> origin - 4
//   colw       : word column of this lane's window for half 0 ( = unit word base + t + ((g+1)>>2) )
//   sh         : funnel shift in bits ( = 8 * ((g+1) & 3) )
//   mx[h]      : byte mask of the observed pixels of half h (x margins, frame clip, block not flat)
// Row slots are indexed by (row - ys) & 3; P[s] is the operand pair of a row, Q[s] the operand
// quad of rows (s, s+1).

struct Window {
  uint32_t Q[4][4];
  uint32_t P[4][2];
};

// x & m, emitted as a distinct instruction per TAG.  A window word is consumed from three different
// operand tuples (the row's B pair, the upper half of one A quad, the lower half of the next) and
// mma.sync needs each tuple in consecutive aligned registers; without distinct values ptxas keeps one
// copy and rebuilds the quads with ~12 moves per k-step.  Three ANDs per word is the minimum.
template <int TAG>
__device__ __forceinline__ uint32_t and_tag(uint32_t x, uint32_t m) {
  uint32_t r;
  if (TAG == 0) asm("lop3.b32 %0, %1, %2, 0, 0xC0;" : "=r"(r) : "r"(x), "r"(m));
  if (TAG == 1) asm("lop3.b32 %0, %1, %2, 1, 0xC0;" : "=r"(r) : "r"(x), "r"(m));
  if (TAG == 2) asm("lop3.b32 %0, %1, %2, 2, 0xC0;" : "=r"(r) : "r"(x), "r"(m));
  return r;
}

// Fetches the lane's window of one tile row (p: word of half 0; half 1 is 4 words = 16 pixels further)
// and files it as: B pair of the row, upper row of quad `qprev`, lower row of quad `qthis`.
__device__ __forceinline__ void fetch_row(const uint32_t *__restrict__ p, int sh, const uint32_t (&mx)[2],
                                          uint32_t (&pair)[2], uint32_t (&qprev)[4], uint32_t (&qthis)[4]) {
  const uint32_t r0 = __funnelshift_r(p[0], p[1], sh);
  const uint32_t r1 = __funnelshift_r(p[4], p[5], sh);
  pair[0] = and_tag<0>(r0, mx[0]);
  pair[1] = and_tag<0>(r1, mx[1]);
  qprev[1] = and_tag<1>(r0, mx[0]);
  qprev[3] = and_tag<1>(r1, mx[1]);
  qthis[0] = and_tag<2>(r0, mx[0]);
  qthis[2] = and_tag<2>(r1, mx[1]);
}

// Rows ys, ys+1, ys+2 -> slots 0, 1, 2 (quad 3's upper row is scratch here).
template <int PITCH>
__device__ __forceinline__ void window_init(Window &w, const uint32_t *__restrict__ p, int sh,
                                            const uint32_t (&mx)[2]) {
  fetch_row(p, sh, mx, w.P[0], w.Q[3], w.Q[0]);
  fetch_row(p + PITCH, sh, mx, w.P[1], w.Q[0], w.Q[1]);
  fetch_row(p + 2 * PITCH, sh, mx, w.P[2], w.Q[1], w.Q[2]);
}

// One k-step of the six common tiles; K is the (compile-time) slot of the step's first row,
// pnew the lane's window word in the tile row three below it.
template <int K>
__device__ __forceinline__ void step6(Window &w, const uint32_t *__restrict__ pnew, int sh, const uint32_t (&mx)[2],
                                      int (&acc)[6][4]) {
  fetch_row(pnew, sh, mx, w.P[(K + 3) & 3], w.Q[(K + 2) & 3], w.Q[(K + 3) & 3]);
  imma_16832(acc[0], w.Q[K & 3], w.P[K & 3]);
  imma_16832(acc[1], w.Q[K & 3], w.P[(K + 1) & 3]);
  imma_16832(acc[2], w.Q[K & 3], w.P[(K + 2) & 3]);
  imma_16832(acc[3], w.Q[K & 3], w.P[(K + 3) & 3]);
  imma_16832(acc[4], w.Q[(K + 2) & 3], w.P[(K + 2) & 3]);
  imma_16832(acc[5], w.Q[(K + 2) & 3], w.P[(K + 3) & 3]);
}

// p0: the lane's window word in tile row ys (the first observed row's cy = -3 tap row).
__device__ __forceinline__ void luma_rows(const uint32_t *__restrict__ p0, int sh, int nrows,
                                          const uint32_t (&mx)[2], int (&acc)[6][4]) {
  Window w;
  window_init<kPL>(w, p0, sh, mx);
  const uint32_t *p = p0 + 3 * kPL;
  int n = nrows;
  for (; n >= 4; n -= 4, p += 4 * kPL) {
    step6<0>(w, p, sh, mx, acc);
    step6<1>(w, p + kPL, sh, mx, acc);
    step6<2>(w, p + 2 * kPL, sh, mx, acc);
    step6<3>(w, p + 3 * kPL, sh, mx, acc);
  }
  if (n > 0) step6<0>(w, p, sh, mx, acc);
  if (n > 1) step6<1>(w, p + kPL, sh, mx, acc);
  if (n > 2) step6<2>(w, p + 2 * kPL, sh, mx, acc);
}

// Chroma.  The luma tap (split as 8*hi + lo so both parts fit int8) rides in the otherwise unused lane
// group g = 7: those lanes step through hs instead of the residual tile (same pitch, funnel shift 0), so
// their A rows 7 / 15 of the lower m-tile are hi(y) / lo(y) and the four existing tiles (m0 x n0..n3)
// deliver every (tap, luma tap) product for free.  The quad's upper row is a separate register from the
// next quad's lower row, which is what lets it carry lo(y) instead of hi(y+1).
//   p  : window word of the fetched tile row (g = 7: hs row of the same index)
//   lp : ls word of the row above the fetched one (only g = 7 lanes use the value)
__device__ __forceinline__ void fetch_row_c(const uint32_t *__restrict__ p, const uint32_t *__restrict__ lp, int sh,
                                            bool is7, const uint32_t (&mx)[2], uint32_t (&pair)[2],
                                            uint32_t (&qprev)[4], uint32_t (&qthis)[4]) {
  const uint32_t r0 = __funnelshift_r(p[0], p[1], sh);
  const uint32_t r1 = __funnelshift_r(p[4], p[5], sh);
  const uint32_t x0 = is7 ? lp[0] : r0;
  const uint32_t x1 = is7 ? lp[4] : r1;
  pair[0] = and_tag<0>(r0, mx[0]);
  pair[1] = and_tag<0>(r1, mx[1]);
  qprev[1] = and_tag<1>(x0, mx[0]);
  qprev[3] = and_tag<1>(x1, mx[1]);
  qthis[0] = and_tag<2>(r0, mx[0]);
  qthis[2] = and_tag<2>(r1, mx[1]);
}

template <int K>
__device__ __forceinline__ void step6c(Window &w, const uint32_t *__restrict__ pnew, const uint32_t *__restrict__ lp,
                                       int sh, bool is7, const uint32_t (&mx)[2], int (&acc)[6][4]) {
  fetch_row_c(pnew, lp, sh, is7, mx, w.P[(K + 3) & 3], w.Q[(K + 2) & 3], w.Q[(K + 3) & 3]);
  imma_16832(acc[0], w.Q[K & 3], w.P[K & 3]);
  imma_16832(acc[1], w.Q[K & 3], w.P[(K + 1) & 3]);
  imma_16832(acc[2], w.Q[K & 3], w.P[(K + 2) & 3]);
  imma_16832(acc[3], w.Q[K & 3], w.P[(K + 3) & 3]);
  imma_16832(acc[4], w.Q[(K + 2) & 3], w.P[(K + 2) & 3]);
  imma_16832(acc[5], w.Q[(K + 2) & 3], w.P[(K + 3) & 3]);
}

// p0 / lp0: the lane's words for tile row ys (lp0 already points one ls row above it).
__device__ __forceinline__ void chroma_rows(const uint32_t *__restrict__ p0, const uint32_t *__restrict__ lp0, int sh,
                                            bool is7, int nrows, const uint32_t (&mx)[2], int (&acc)[6][4]) {
  Window w;
  fetch_row_c(p0, lp0, sh, is7, mx, w.P[0], w.Q[3], w.Q[0]);
  fetch_row_c(p0 + kPC, lp0 + kPC, sh, is7, mx, w.P[1], w.Q[0], w.Q[1]);
  fetch_row_c(p0 + 2 * kPC, lp0 + 2 * kPC, sh, is7, mx, w.P[2], w.Q[1], w.Q[2]);
  const uint32_t *p = p0 + 3 * kPC, *lp = lp0 + 3 * kPC;
  int n = nrows;
  for (; n >= 4; n -= 4, p += 4 * kPC, lp += 4 * kPC) {
    step6c<0>(w, p, lp, sh, is7, mx, acc);
    step6c<1>(w, p + kPC, lp + kPC, sh, is7, mx, acc);
    step6c<2>(w, p + 2 * kPC, lp + 2 * kPC, sh, is7, mx, acc);
    step6c<3>(w, p + 3 * kPC, lp + 3 * kPC, sh, is7, mx, acc);
  }
  if (n > 0) step6c<0>(w, p, lp, sh, is7, mx, acc);
  if (n > 1) step6c<1>(w, p + kPC, lp + kPC, sh, is7, mx, acc);
  if (n > 2) step6c<2>(w, p + 2 * kPC, lp + 2 * kPC, sh, is7, mx, acc);
}

// MMA tap index a = 8q+g  ->  record tap index (0..23 AR taps, 25 centre sample), -1 unused.
__device__ __forceinline__ int record_tap(int a) {
  const int q = a >> 3, g = a & 7;
  if (g == 7) return -1;
  if (q < 3) return 7 * q + g;
  if (g < 3) return 21 + g;
  if (g == 3) return 25;
  return -1;
}

__device__ __forceinline__ int pair_index(int i, int j) { return i * kTaps - i * (i - 1) / 2 + (j - i); }

// Adds accumulator element D[a][b] (a <= b: the mirrored element is covered by another tile).
__device__ __forceinline__ void emit(unsigned long long *gram, int a, int b, int v) {
  if (a > b || v == 0) return;
  const int ia = record_tap(a), ib = record_tap(b);
  if (ia < 0 || ib < 0) return;
  atomicAdd(&gram[pair_index(min(ia, ib), max(ia, ib))], (unsigned long long)(long long)v);
}
// D[luma tap part][b]: weight 8 for the high part, 1 for the low part.
__device__ __forceinline__ void emit_luma_tap(unsigned long long *gram, int b, int v, int weight) {
  const int ib = record_tap(b);
  if (ib < 0 || v == 0) return;
  atomicAdd(&gram[pair_index(min(ib, 24), max(ib, 24))], (unsigned long long)((long long)v * weight));
}

// ------------------------------------------------------------------------------ TMA staging
//
// With 16-byte aligned planes the raw source / denoised tiles of a super-unit are fetched by the TMA
// engine (cp.async.bulk.tensor.2d, SASS UTMALDG) straight into shared memory: six boxes per unit, one
// elected thread, no per-thread address arithmetic, frame edges zero-filled by the hardware, and the
// fetch of the NEXT unit overlaps the k-loops of the current one (single mbarrier, phase per unit).
// The innermost box coordinate must be a multiple of 16 BYTES (anything else raises an illegal-instruction
// fault, tools/tma_probe.cu), so the box starts 16 bytes left of the unit (8 or 16 samples, of which the
// tile uses the last 4), and its width keeps the inner extent a multiple of 16 bytes.
__host__ __device__ constexpr int tma_origin(int bytes) { return 16 / bytes; }
__host__ __device__ constexpr int luma_box_w(int bytes) { return bytes == 2 ? 80 : 96; }    // >= origin + 64 + 3
__host__ __device__ constexpr int chroma_box_w(int bytes) { return bytes == 2 ? 48 : 64; }  // >= origin + 32 + 3
__host__ __device__ constexpr int align128(int v) { return (v + 127) & ~127; }

template <int SB, int DB>
struct RawLayout {
  static constexpr int kLumaS = 0;
  static constexpr int kLumaSBytes = kLumaRows * luma_box_w(SB) * SB;
  static constexpr int kLumaD = kLumaS + align128(kLumaSBytes);
  static constexpr int kLumaDBytes = kLumaRows * luma_box_w(DB) * DB;
  static constexpr int kChromaSBytes = kChromaRows * chroma_box_w(SB) * SB;
  static constexpr int kChromaDBytes = kChromaRows * chroma_box_w(DB) * DB;
  static constexpr int kCbS = kLumaD + align128(kLumaDBytes);
  static constexpr int kCbD = kCbS + align128(kChromaSBytes);
  static constexpr int kCrS = kCbD + align128(kChromaDBytes);
  static constexpr int kCrD = kCrS + align128(kChromaSBytes);
  static constexpr int kBytes = kCrD + align128(kChromaDBytes);
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "G1S_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra G1S_DONE;\n"
      "bra G1S_WAIT;\n"
      "G1S_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int x, int y, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

template <int SB, int DB, bool TMA>
__global__ void __launch_bounds__(kSuThreads, 4)
gram_imma_kernel(const FrameDesc *__restrict__ frames, Geometry g, uint8_t *__restrict__ records, RecordLayout rl,
                 int runs_per_row, int aligned, const uint8_t *__restrict__ tmaps) {
  __shared__ SuSmem sm;
  using RL = RawLayout<SB, DB>;
  __shared__ __align__(128) uint8_t raw[TMA ? RL::kBytes : 16];
  __shared__ __align__(8) uint64_t raw_bar;
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
  const int Y0 = 32 * by, CY0 = 16 * by;
  const bool rows_inside = Y0 >= 3 && Y0 + 32 <= H && (!has_chroma || CY0 + 16 <= ph);

  int acc[6][4];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[i][r] = 0;
  long long nobs[3] = {0, 0, 0};  // kept by the bookkeeping thread only
  int selfp[2][3] = {{0, 0, 0}, {0, 0, 0}};  // luma-tap self products per chroma plane, threads 0..127

  for (int i = tid; i < (int)(sizeof(SuSmem) / 4); i += kSuThreads) reinterpret_cast<uint32_t *>(&sm)[i] = 0;
  __syncthreads();
  // flat flags of this run: this block row and the one above, one extra block on each side
  const int fbase = 2 * u_beg - 1;
  for (int i = tid; i < 2 * kFlagCols; i += kSuThreads) {
    const int r = i >= kFlagCols ? 1 : 0, k = i - r * kFlagCols;
    const int bx = fbase + k, yy = by - r;
    uint8_t v = 0;
    if (bx >= 0 && bx < g.nbw && yy >= 0) v = flat[yy * g.nbw + bx];
    sm.flat[r][k] = v;
  }
  __syncthreads();

  const int sh = 8 * ((gq + 1) & 3);
  const int dxw = (gq + 1) >> 2;
  const bool is7 = gq == 7;
  constexpr int kBook = kSuThreads - 1;  // bookkeeping thread (observation rectangles, counts, flags)

  // next super-unit of the run with at least one flat block (uniform across the CTA)
  auto next_flat = [&](int u) {
    while (u < u_end && !(sm.flat[0][2 * u - fbase] | sm.flat[0][2 * u - fbase + 1])) ++u;
    return u;
  };
  // one elected thread asks the TMA engine for the six raw tiles of a unit
  const uint8_t *fmaps = TMA ? tmaps + (size_t)f * 6 * 128 : nullptr;
  auto issue_tma = [&](int u) {
    const int X0 = 64 * u, CX0 = 32 * u;
    const uint32_t bytes = RL::kLumaSBytes + RL::kLumaDBytes + (has_chroma ? 2 * (RL::kChromaSBytes + RL::kChromaDBytes) : 0);
    mbar_expect_tx(&raw_bar, bytes);
    tma_load_2d(raw + RL::kLumaS, fmaps + 0 * 128, X0 - tma_origin(SB), Y0 - 3, &raw_bar);
    tma_load_2d(raw + RL::kLumaD, fmaps + 1 * 128, X0 - tma_origin(DB), Y0 - 3, &raw_bar);
    if (has_chroma) {
      tma_load_2d(raw + RL::kCbS, fmaps + 2 * 128, CX0 - tma_origin(SB), CY0 - 3, &raw_bar);
      tma_load_2d(raw + RL::kCbD, fmaps + 3 * 128, CX0 - tma_origin(DB), CY0 - 3, &raw_bar);
      tma_load_2d(raw + RL::kCrS, fmaps + 4 * 128, CX0 - tma_origin(SB), CY0 - 3, &raw_bar);
      tma_load_2d(raw + RL::kCrD, fmaps + 5 * 128, CX0 - tma_origin(DB), CY0 - 3, &raw_bar);
    }
  };
  int u = next_flat(u_beg);
  uint32_t raw_phase = 0;
  if (TMA) {
    if (tid == 0) {
      mbar_init(&raw_bar, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (u < u_end) issue_tma(u);
    }
    __syncthreads();
  }

  while (u < u_end) {
    const int bx0 = 2 * u, fk = bx0 - fbase;
    const bool fl0 = sm.flat[0][fk] != 0, fl1 = sm.flat[0][fk + 1] != 0;
    const int b0 = by * g.nbw + bx0;

    if (tid < 6) {
      sm.st_rs[tid] = 0;
      sm.st_rq[tid] = 0;
      if (tid < 2) sm.st_ls[tid] = 0;
      if (tid < 3) sm.ovf[tid] = 0;
    }
    if (tid == kBook) {
      const bool lf0 = sm.flat[0][fk - 1] != 0, rt1 = sm.flat[0][fk + 2] != 0;
      const bool up0 = sm.flat[1][fk] != 0, up1 = sm.flat[1][fk + 1] != 0;
      UnitInfo in;
      in.xs0 = lf0 ? 0 : kLag;
      in.xs1 = fl0 ? 0 : kLag;
      in.y00 = up0 ? 0 : kLag;
      in.y01 = up1 ? 0 : kLag;
      in.x1l0 = min(W - 32 * bx0 - kLag, fl1 ? 32 : 32 - kLag);
      in.x1l1 = min(W - 32 * (bx0 + 1) - kLag, rt1 ? 32 : 32 - kLag);
      in.x1c0 = min(pw - 16 * bx0 - kLag, fl1 ? 16 : 16 - kLag);
      in.x1c1 = min(pw - 16 * (bx0 + 1) - kLag, rt1 ? 16 : 16 - kLag);
      in.y1l = min(H - Y0, 32);
      in.y1c = min(ph - CY0, 16);
      sm.info = in;
    }
    __syncthreads();  // previous unit's k-loops are done with the tiles; statistics zeroed; info written

    // ------------------------------------------------------------------ staging
    const int X0 = 64 * u, CX0 = 32 * u;
    const bool fast = aligned && rows_inside && X0 >= 4 && X0 + 68 <= W && (!has_chroma || CX0 + 36 <= pw);
    if (TMA) {
      // raw tiles of this unit were requested one unit ago; each thread converts the words it owns
      mbar_wait(&raw_bar, raw_phase);
      raw_phase ^= 1;
      constexpr int kLpS = luma_box_w(SB) * SB, kLpD = luma_box_w(DB) * DB;        // raw row pitches in bytes
      constexpr int kCpS = chroma_box_w(SB) * SB, kCpD = chroma_box_w(DB) * DB;
      constexpr int kOs = (tma_origin(SB) - 4) * SB, kOd = (tma_origin(DB) - 4) * DB;  // tile column 0 inside a raw row
      {
        uint32_t ov = 0;
        int rs = 0, rq = 0;
        unsigned ls = 0;
        const int ty0 = 2 * warp + (lane >> 4), w = 1 + (lane & 15);
        const uint8_t *sp = raw + RL::kLumaS + ty0 * kLpS + kOs + 4 * w * SB;
        const uint8_t *dp = raw + RL::kLumaD + ty0 * kLpD + kOd + 4 * w * DB;
        uint32_t *dst = &sm.luma[ty0 * kPL + w];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const int ty = ty0 + 12 * p;
          if (ty < kLumaRows) {
            int trs = 0, trq = 0;
            unsigned tls = 0;
            dst[12 * p * kPL] = residual4<SB, DB, true, true, true>(sp + 12 * p * kLpS, dp + 12 * p * kLpD, g.src_shift,
                                                                    g.den_shift, trs, trq, tls, ov);
            if (ty >= 3) rs += trs, rq += trq, ls += tls;  // halo rows belong to the block above
          }
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
          atomicAdd(&sm.st_rq[blk], (unsigned)rq);
          atomicAdd(&sm.st_ls[blk], ls);
        }
        if (__any_sync(0xffffffffu, ov != 0) && lane == 0) sm.ovf[0] = 1;
      }
      if (has_chroma) {
        const int c = warp >= 3 ? 1 : 0, wk = warp - 3 * c;
        int rs = 0, rq = 0;
        unsigned ls = 0;
        uint32_t ovc = 0;
        const int ty0 = 4 * wk + (lane >> 3), w = 1 + (lane & 7);
        const uint8_t *sp = raw + (c ? RL::kCrS : RL::kCbS) + ty0 * kCpS + kOs + 4 * w * SB;
        const uint8_t *dp = raw + (c ? RL::kCrD : RL::kCbD) + ty0 * kCpD + kOd + 4 * w * DB;
        uint32_t *dst = &sm.chroma[c][ty0 * kPC + w];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int ty = ty0 + 12 * p;
          if (ty < kChromaRows) {
            int trs = 0, trq = 0;
            dst[12 * p * kPC] = residual4<SB, DB, true, false, true>(sp + 12 * p * kCpS, dp + 12 * p * kCpD, g.src_shift,
                                                                     g.den_shift, trs, trq, ls, ovc);
            if (ty >= 3) rs += trs, rq += trq;
          }
        }
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          rs += __shfl_xor_sync(0xffffffffu, rs, o);
          rq += __shfl_xor_sync(0xffffffffu, rq, o);
        }
        if ((lane & 3) == 0) {
          const int blk = (lane >> 2) & 1;
          atomicAdd(&sm.st_rs[2 + 2 * c + blk], rs);
          atomicAdd(&sm.st_rq[2 + 2 * c + blk], (unsigned)rq);
        }
        if (__any_sync(0xffffffffu, ovc != 0) && lane == 0) sm.ovf[1 + c] = 1;
      }
      {
        int rs = 0, rq = 0;
        unsigned ls = 0;
        uint32_t ovh = 0;
        if (tid < 70) {
          const int ty = tid >> 1, w = (tid & 1) * 17;
          sm.luma[ty * kPL + w] = residual4<SB, DB, false, false, true>(
              raw + RL::kLumaS + ty * kLpS + kOs + 4 * w * SB, raw + RL::kLumaD + ty * kLpD + kOd + 4 * w * DB, g.src_shift,
              g.den_shift, rs, rq, ls, ovh);
          if (ovh) sm.ovf[0] = 1;
        } else if (has_chroma && tid < 70 + 76) {
          const int idx = tid - 70, c = idx >= 38 ? 1 : 0, rem = idx - 38 * c;
          const int ty = rem >> 1, w = (rem & 1) * 9;
          sm.chroma[c][ty * kPC + w] = residual4<SB, DB, false, false, true>(
              raw + (c ? RL::kCrS : RL::kCbS) + ty * kCpS + kOs + 4 * w * SB,
              raw + (c ? RL::kCrD : RL::kCbD) + ty * kCpD + kOd + 4 * w * DB, g.src_shift, g.den_shift, rs, rq, ls, ovh);
          if (ovh) sm.ovf[1 + c] = 1;
        }
      }
    } else if (fast) {
      {
        // luma main words: two tile rows per pass (lanes 0-15 / 16-31), rows warp*2 + 12*pass
        uint32_t ov = 0;
        int rs = 0, rq = 0;
        unsigned ls = 0;
        const int ty0 = 2 * warp + (lane >> 4), w = 1 + (lane & 15);
        const uint32_t sstr = fd.src_stride[0], dstr = fd.den_stride[0];
        const uint8_t *sp = static_cast<const uint8_t *>(fd.src[0]) + (size_t)(Y0 - 3 + ty0) * sstr +
                            (size_t)(X0 - 4 + 4 * w) * SB;
        const uint8_t *dp = static_cast<const uint8_t *>(fd.den[0]) + (size_t)(Y0 - 3 + ty0) * dstr +
                            (size_t)(X0 - 4 + 4 * w) * DB;
        uint32_t *dst = &sm.luma[ty0 * kPL + w];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const int ty = ty0 + 12 * p;
          if (ty < kLumaRows) {
            int trs = 0, trq = 0;
            unsigned tls = 0;
            dst[12 * p * kPL] = residual4<SB, DB, true, true>(sp, dp, g.src_shift, g.den_shift, trs, trq, tls, ov);
            if (ty >= 3) rs += trs, rq += trq, ls += tls;  // halo rows belong to the block above
          }
          sp += (size_t)12 * sstr;
          dp += (size_t)12 * dstr;
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
          atomicAdd(&sm.st_rq[blk], (unsigned)rq);
          atomicAdd(&sm.st_ls[blk], ls);
        }
        if (__any_sync(0xffffffffu, ov != 0) && lane == 0) sm.ovf[0] = 1;
      }
      if (has_chroma) {
        // chroma main words: warps 0-2 Cb, 3-5 Cr; four tile rows per pass, rows 4*(warp%3) + 12*pass
        const int c = warp >= 3 ? 1 : 0, wk = warp - 3 * c;
        int rs = 0, rq = 0;
        unsigned ls = 0;
        uint32_t ovc = 0;
        const int ty0 = 4 * wk + (lane >> 3), w = 1 + (lane & 7);
        const uint32_t sstr = fd.src_stride[1 + c], dstr = fd.den_stride[1 + c];
        const uint8_t *sp = static_cast<const uint8_t *>(fd.src[1 + c]) + (size_t)(CY0 - 3 + ty0) * sstr +
                            (size_t)(CX0 - 4 + 4 * w) * SB;
        const uint8_t *dp = static_cast<const uint8_t *>(fd.den[1 + c]) + (size_t)(CY0 - 3 + ty0) * dstr +
                            (size_t)(CX0 - 4 + 4 * w) * DB;
        uint32_t *dst = &sm.chroma[c][ty0 * kPC + w];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int ty = ty0 + 12 * p;
          if (ty < kChromaRows) {
            int trs = 0, trq = 0;
            dst[12 * p * kPC] = residual4<SB, DB, true, false>(sp, dp, g.src_shift, g.den_shift, trs, trq, ls, ovc);
            if (ty >= 3) rs += trs, rq += trq;
          }
          sp += (size_t)12 * sstr;
          dp += (size_t)12 * dstr;
        }
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          rs += __shfl_xor_sync(0xffffffffu, rs, o);
          rq += __shfl_xor_sync(0xffffffffu, rq, o);
        }
        if ((lane & 3) == 0) {
          const int blk = (lane >> 2) & 1;
          atomicAdd(&sm.st_rs[2 + 2 * c + blk], rs);
          atomicAdd(&sm.st_rq[2 + 2 * c + blk], (unsigned)rq);
        }
        if (__any_sync(0xffffffffu, ovc != 0) && lane == 0) sm.ovf[1 + c] = 1;
      }
      // halo words: luma 0 and 17 (70 items), chroma 0 and 9 (2 planes x 38 items)
      {
        int rs = 0, rq = 0;
        unsigned ls = 0;
        uint32_t ovh = 0;
        if (tid < 70) {
          const int ty = tid >> 1, w = (tid & 1) * 17;
          const uint8_t *sp = static_cast<const uint8_t *>(fd.src[0]) + (size_t)(Y0 - 3 + ty) * fd.src_stride[0] +
                              (size_t)(X0 - 4 + 4 * w) * SB;
          const uint8_t *dp = static_cast<const uint8_t *>(fd.den[0]) + (size_t)(Y0 - 3 + ty) * fd.den_stride[0] +
                              (size_t)(X0 - 4 + 4 * w) * DB;
          sm.luma[ty * kPL + w] = residual4<SB, DB, false, false>(sp, dp, g.src_shift, g.den_shift, rs, rq, ls, ovh);
          if (ovh) sm.ovf[0] = 1;
        } else if (has_chroma && tid < 70 + 76) {
          const int idx = tid - 70, c = idx >= 38 ? 1 : 0, rem = idx - 38 * c;
          const int ty = rem >> 1, w = (rem & 1) * 9;
          const uint8_t *sp = static_cast<const uint8_t *>(fd.src[1 + c]) +
                              (size_t)(CY0 - 3 + ty) * fd.src_stride[1 + c] + (size_t)(CX0 - 4 + 4 * w) * SB;
          const uint8_t *dp = static_cast<const uint8_t *>(fd.den[1 + c]) +
                              (size_t)(CY0 - 3 + ty) * fd.den_stride[1 + c] + (size_t)(CX0 - 4 + 4 * w) * DB;
          sm.chroma[c][ty * kPC + w] =
              residual4<SB, DB, false, false>(sp, dp, g.src_shift, g.den_shift, rs, rq, ls, ovh);
          if (ovh) sm.ovf[1 + c] = 1;
        }
      }
    } else {
      // frame-edge or unaligned unit: scalar bounds-checked loads, same tile contents
      int rs[6] = {0, 0, 0, 0, 0, 0}, rq[6] = {0, 0, 0, 0, 0, 0};
      unsigned ls[2] = {0, 0};
      uint32_t ov[3] = {0, 0, 0};
      for (int e = tid; e < kLumaRows * 18; e += kSuThreads) {
        const int ty = e / 18, w = e - 18 * ty;
        const bool st = ty >= 3 && w >= 1 && w <= 16;
        const int blk = w > 8 ? 1 : 0;
        int a = 0, b = 0;
        unsigned l = 0;
        sm.luma[ty * kPL + w] = residual4_slow(fd.src[0], fd.src_stride[0], fd.den[0], fd.den_stride[0], Y0 - 3 + ty,
                                               X0 - 4 + 4 * w, g, W, H, st, a, b, l, ov[0]);
        if (blk) rs[1] += a, rq[1] += b, ls[1] += l;
        else rs[0] += a, rq[0] += b, ls[0] += l;
      }
      if (has_chroma) {
        for (int e = tid; e < 2 * kChromaRows * 10; e += kSuThreads) {
          const int c = e >= kChromaRows * 10 ? 1 : 0, r = e - c * kChromaRows * 10;
          const int ty = r / 10, w = r - 10 * ty;
          const bool st = ty >= 3 && w >= 1 && w <= 8;
          const int blk = w > 4 ? 1 : 0;
          int a = 0, b = 0;
          unsigned l = 0;
          const uint32_t word =
              residual4_slow(fd.src[1 + c], fd.src_stride[1 + c], fd.den[1 + c], fd.den_stride[1 + c], CY0 - 3 + ty,
                             CX0 - 4 + 4 * w, g, pw, ph, st, a, b, l, c ? ov[2] : ov[1]);
          sm.chroma[c][ty * kPC + w] = word;
          if (c) {
            if (blk) rs[5] += a, rq[5] += b;
            else rs[4] += a, rq[4] += b;
          } else {
            if (blk) rs[3] += a, rq[3] += b;
            else rs[2] += a, rq[2] += b;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        if (rs[k]) atomicAdd(&sm.st_rs[k], rs[k]);
        if (rq[k]) atomicAdd(&sm.st_rq[k], (unsigned)rq[k]);
      }
      if (ls[0]) atomicAdd(&sm.st_ls[0], ls[0]);
      if (ls[1]) atomicAdd(&sm.st_ls[1], ls[1]);
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (ov[k]) sm.ovf[k] = 1;
    }
    __syncthreads();  // tiles + statistics + overflow flags complete (and the raw tiles are free again)
    const int u_next = next_flat(u + 1);
    if (TMA && tid == 0 && u_next < u_end) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of raw before the async refill
      issue_tma(u_next);
    }

    const UnitInfo in = sm.info;
    const bool ovl = sm.ovf[0] != 0;
    const bool ovcb = ovl || sm.ovf[1] != 0, ovcr = ovl || sm.ovf[2] != 0;  // the luma tap needs an exact luma tile

    // luma tap of the chroma planes: sum of the co-sited 2x2 luma residuals = 8*hi + lo, and its
    // self products over the observed pixels (the only Gram entries the k-loops do not produce)
    if (has_chroma && tid < 128) {
      const int cy = tid >> 3, w = tid & 7;
      const uint32_t *r0 = &sm.luma[(3 + 2 * cy) * kPL + 1 + 2 * w];
      const uint32_t *r1 = r0 + kPL;
      const uint32_t a0 = r0[0], a1 = r0[1], c0 = r1[0], c1 = r1[1];
      uint32_t hw = 0, lw = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t a = (i < 2 ? a0 : a1) >> (16 * (i & 1));
        const uint32_t b = (i < 2 ? c0 : c1) >> (16 * (i & 1));
        const int l4 = __dp4a((int)__byte_perm(a, b, 0x5410), 0x01010101, 0);
        hw |= (uint32_t)((l4 >> 3) & 0xFF) << (8 * i);
        lw |= (uint32_t)(l4 & 7) << (8 * i);
      }
      sm.hs[cy * kPC + w] = hw;
      sm.ls[(cy + 1) * kPC + w] = lw;
      const int j = w >> 2;  // block of the pair
      const bool flj = j ? fl1 : fl0;
      const int xs = j ? in.xs1 : in.xs0, x1 = j ? in.x1c1 : in.x1c0, y0 = j ? in.y01 : in.y00;
      const uint32_t m = (flj && cy >= y0 && cy < in.y1c) ? byte_mask(4 * (w & 3), xs, x1) : 0u;
      const int hm = (int)(hw & m), lm = (int)(lw & m);
      // per thread and unit at most 4 * 128^2; summed over the run in registers, reduced once at the end
      const int hh = __dp4a(hm, hm, 0), hl = __dp4a(hm, lm, 0), ll = __dp4a(lm, lm, 0);
      if (!ovcb) selfp[0][0] += hh, selfp[0][1] += hl, selfp[0][2] += ll;
      if (!ovcr) selfp[1][0] += hh, selfp[1][1] += hl, selfp[1][2] += ll;
    }
    // statistics and overflow flags out (each block belongs to exactly one CTA)
    if (tid >= 128 && tid < 134) {
      const int k = tid - 128, c = k >> 1, blk = k & 1;
      const bool fl = blk ? fl1 : fl0;
      if (fl && (c == 0 || has_chroma)) {
        reinterpret_cast<int32_t *>(rec + rl.off_rsum)[c * g.nb + b0 + blk] = sm.st_rs[k];
        reinterpret_cast<uint32_t *>(rec + rl.off_rsq)[c * g.nb + b0 + blk] = sm.st_rq[k];
        if (c == 0) reinterpret_cast<uint32_t *>(rec + rl.off_luma_sum)[b0 + blk] = sm.st_ls[blk];
        const bool o = c == 0 ? ovl : (c == 1 ? ovcb : ovcr);
        if (o) {
          ovf_out[(size_t)c * g.nb + b0 + blk] = 1;
          atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_ovf_count), 1ull);
        }
      }
    }
    if (tid == kBook) {  // observation counts of the blocks this kernel accumulates
      const long long hl0 = (in.x1l0 > in.xs0 && in.y1l > in.y00) ? (long long)(in.x1l0 - in.xs0) * (in.y1l - in.y00) : 0;
      const long long hl1 = (in.x1l1 > in.xs1 && in.y1l > in.y01) ? (long long)(in.x1l1 - in.xs1) * (in.y1l - in.y01) : 0;
      const long long hc0 = (in.x1c0 > in.xs0 && in.y1c > in.y00) ? (long long)(in.x1c0 - in.xs0) * (in.y1c - in.y00) : 0;
      const long long hc1 = (in.x1c1 > in.xs1 && in.y1c > in.y01) ? (long long)(in.x1c1 - in.xs1) * (in.y1c - in.y01) : 0;
      if (!ovl) nobs[0] += (fl0 ? hl0 : 0) + (fl1 ? hl1 : 0);
      if (has_chroma && !ovcb) nobs[1] += (fl0 ? hc0 : 0) + (fl1 ? hc1 : 0);
      if (has_chroma && !ovcr) nobs[2] += (fl0 ? hc0 : 0) + (fl1 ? hc1 : 0);
    }
    __syncthreads();  // hs / ls visible

    // ------------------------------------------------------------------ k-loops
    if (warp < 4) {
      const int j = warp >> 1;
      const bool fl = j ? fl1 : fl0;
      if (fl && !ovl) {
        const int xs = j ? in.xs1 : in.xs0, y0 = j ? in.y01 : in.y00, x1 = j ? in.x1l1 : in.x1l0, y1 = in.y1l;
        if (x1 > xs && y1 > y0) {
          // first warp of the block takes a multiple of four rows so only one warp has a ragged tail
          const int n = y1 - y0, na = min(n, ((n >> 1) + 3) & ~3);
          const int ys = (warp & 1) ? y0 + na : y0, nr = (warp & 1) ? n - na : na;
          const uint32_t mx[2] = {byte_mask(4 * t, xs, x1), byte_mask(16 + 4 * t, xs, x1)};
          if (nr > 0) luma_rows(&sm.luma[ys * kPL + 8 * j + t + dxw], sh, nr, mx, acc);
        }
      }
    } else if (has_chroma) {
      const int c = warp - 4;
      const bool ovc = c ? ovcr : ovcb;
      const int y1 = in.y1c;
      // half 0 <-> chroma block bx0, half 1 <-> block bx0 + 1
      const bool on0 = fl0 && !ovc && in.x1c0 > in.xs0 && y1 > in.y00;
      const bool on1 = fl1 && !ovc && in.x1c1 > in.xs1 && y1 > in.y01;
      const uint32_t m0 = on0 ? byte_mask(4 * t, in.xs0, in.x1c0) : 0u;
      const uint32_t m1 = on1 ? byte_mask(4 * t, in.xs1, in.x1c1) : 0u;
      const int ya = on0 ? in.y00 : 99, yb = on1 ? in.y01 : 99;
      const int ylo = min(ya, yb), yhi = min(max(ya, yb), y1);
      // g = 7 lanes walk hs (and ls one row up) instead of the residual tile
      const uint32_t *base = is7 ? &sm.hs[t] : &sm.chroma[c][t + dxw];
      const uint32_t *lbase = &sm.ls[t];  // storage row r holds ls row r - 1
      if (ylo < y1) {
        // rows where only one block of the pair is observed (its top margin is 0, the other's is 3)
        if (yhi > ylo) {
          const uint32_t mx[2] = {ya <= ylo ? m0 : 0u, yb <= ylo ? m1 : 0u};
          chroma_rows(base + ylo * kPC, lbase + ylo * kPC, sh, is7, yhi - ylo, mx, acc);
        }
        if (y1 > yhi) {
          const uint32_t mx[2] = {m0, m1};
          chroma_rows(base + yhi * kPC, lbase + yhi * kPC, sh, is7, y1 - yhi, mx, acc);
        }
      }
    }
    u = u_next;
  }

  // ---------------------------------------------------------------------- epilogue
  if (has_chroma && warp < 4) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        int v = selfp[c][k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&sm.self[c][k], v);
      }
  }
  __syncthreads();
  if (warp < 4) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (acc[i][r]) atomicAdd(&sm.dl[(i * 4 + r) * 32 + lane], acc[i][r]);
  }
  __syncthreads();
  if (warp == 0 || (warp >= 4 && has_chroma)) {
    const int plane = warp == 0 ? 0 : warp - 3;
    unsigned long long *gram = reinterpret_cast<unsigned long long *>(rec + rl.off_gram) + (size_t)plane * kPairs;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int mrow = i >= 4 ? 16 : 0;
      const int ncol = i >= 4 ? 8 * (i - 2) : 8 * i;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int v = warp == 0 ? sm.dl[(i * 4 + r) * 32 + lane] : acc[i][r];
        const int b = ncol + 2 * t + (r & 1);
        if (plane > 0 && is7 && i < 4)
          emit_luma_tap(gram, b, v, (r >> 1) ? 1 : 8);  // rows 7 / 15 of the lower m-tile: luma tap hi / lo
        else
          emit(gram, mrow + gq + 8 * (r >> 1), b, v);
      }
    }
    if (plane > 0 && lane == 0) {  // (8h + l)^2 = 64 hh + 16 hl + ll
      const long long v = 64ll * sm.self[plane - 1][0] + 16ll * sm.self[plane - 1][1] + (long long)sm.self[plane - 1][2];
      if (v) atomicAdd(&gram[pair_index(24, 24)], (unsigned long long)v);
    }
  }
  if (tid == kBook) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (nobs[c])
        atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_nobs) + c, (unsigned long long)nobs[c]);
  }
}

}  // namespace

bool gram_imma_supported(const Geometry &g) {
  return (g.planes == 1) || (g.planes == 3 && g.ss_x == 1 && g.ss_y == 1);
}

void gram_imma_tma_boxes(int bytes, int *luma_w, int *luma_h, int *chroma_w, int *chroma_h) {
  *luma_w = luma_box_w(bytes);
  *luma_h = kLumaRows;
  *chroma_w = chroma_box_w(bytes);
  *chroma_h = kChromaRows;
}

void launch_gram_imma(const FrameDesc *frames, int nframes, const Geometry &g, uint8_t *records,
                      const RecordLayout &rl, bool aligned, const void *tmaps, cudaStream_t st) {
  const int nsu = (g.nbw + 1) / 2;
  const int runs = (nsu + kSuRun - 1) / kSuRun;
  dim3 grid(runs * g.nbh, nframes);
  const int al = aligned ? 1 : 0;
  const uint8_t *tm = static_cast<const uint8_t *>(tmaps);
#define G1S_LAUNCH(SB, DB)                                                                                   \
  do {                                                                                                       \
    if (tm)                                                                                                  \
      gram_imma_kernel<SB, DB, true><<<grid, kSuThreads, 0, st>>>(frames, g, records, rl, runs, al, tm);     \
    else                                                                                                     \
      gram_imma_kernel<SB, DB, false><<<grid, kSuThreads, 0, st>>>(frames, g, records, rl, runs, al, tm);    \
  } while (0)
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
