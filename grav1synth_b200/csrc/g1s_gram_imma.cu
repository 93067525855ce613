// gram_plan_kernel + gram_imma_kernel — the AR normal-equation (Gram) accumulation on the int8 tensor-core path
// (mma.sync.m16n8k32.s8, SASS IMMA.16832.S8.S8), 4:2:0 and monochrome.
//
// Replaces NoiseModel::add_block_observations (extract_ar_row + the n x n outer-product accumulation) of
// av1-grain's diff module (reached from /root/reference/src/main.rs:442) for every flat block whose
// residuals (and, for chroma, luma tap) fit in int8; the rare block that does not was flagged by
// residual_kernel and is redone exactly by gram_generic_kernel.  All sums are integers, so the result is
// bit-identical to the oracle whatever the summation order.
//
// Two kernels.
//  * gram_plan_kernel (a few CTAs per frame and plane, one thread per block position) turns the flat-block map into
//    the work list of the plane: it applies add_block_observations' rectangle rules (3-sample margins unless the
//    neighbour block is flat, frame clipping), joins vertically adjacent blocks whose column ranges agree into
//    STRIPS (one row window, so the window bookkeeping below is paid once per strip instead of once per block and
//    the 3 halo rows between joined blocks are not read twice), and writes one 16-byte descriptor per strip and
//    tile ("unit").  It also counts the observations and hands slivers of fewer than kMinRows rows to the generic
//    kernel.  All the irregular, scalar work of the path lives here, massively parallel and off the hot loop.
//  * gram_imma_kernel never touches the frames or the flags: residual_kernel left the s8 residual of Y / Cb / Cr
//    and chroma's luma tap in engine-owned planes, and every warp walks a contiguous share of its plane's units,
//    pulling each unit's tile into its own slice of shared memory with the TMA engine (cp.async.bulk.tensor.2d,
//    SASS UTMALDG; frame edges are zero-filled by the hardware, so there is no edge path), two units ahead.
//
// Why warps are autonomous.  Measured on the B200 (tools/imma_probe.cu, profiles/): the legacy IMMA path
// holds a sub-partition's issue port for its whole 8.4 cycles, so a sub-partition's time is
// 8.4 * IMMAs + (every other warp instruction it issues) -- nothing overlaps, polls and barrier spins are
// paid in full.  So: no CTA barriers, no shared rings, no inter-warp waits.  Every warp is the same: it takes chunks of
// 64 units of one plane from a global counter (round 2; a fixed 4 + 4 luma / chroma split measured 25.1 us per frame,
// 5 + 3 28.1, 6 + 2 35.1 -- the dynamic form needs no split), has its own mbarriers and TMA ring, and adds its int32
// accumulators to the frame's int64 record directly at the end of a chunk.
//   unit        one tile of a strip: luma one 32x32 block, chroma two adjacent 16x16 blocks (32 columns either way)
//   Tap a = 8q+g with g = cx+3 (the mma lane group), q = cy+3.  A lane's operand of a residual row is ONE
//   32-bit window per half (two LDS + funnel shift), and the same window is the row's B pair and its half of
//   an A quad.  The arithmetic is the row-pair deduplicated form described above the k-loops: 2.5 MMAs per
//   residual row instead of 6 per observed row (round 1), every row fetched once.
//   The observation mask (block margins, frame clipping) is a byte mask on k applied to the A operand.
//   Chroma's luma tap rides in lane group g = 7 (those lanes walk the luma-tap tile), so it costs no MMA.
#include "g1s_kernels.h"

#include <algorithm>
#include <cstdlib>

namespace g1s {

namespace {

#ifndef G1S_GRAM_WARPS
#define G1S_GRAM_WARPS 8
#endif
constexpr int kGramWarps = G1S_GRAM_WARPS;
constexpr int kGramThreads = 32 * kGramWarps;
constexpr int kMaxObs = 130000;   // observations between two flushes of a warp: bounds the int32 strip accumulators (x 127^2 < 2^31)
constexpr int kMinRows = 6;       // shortest row window the strip code handles (two opening + three closing pair steps)
#ifndef G1S_MAX_STRIP
#define G1S_MAX_STRIP 64
#endif
constexpr int kMaxStrip = G1S_MAX_STRIP;  // block rows per strip: 64 * 1024 observations * 127^2 < 2^31
constexpr int kLumaRows = 35;     // 3 halo rows + 32
constexpr int kChromaRows = 19;   // 3 halo rows + 16 (residual and luma-tap tiles alike)
constexpr int kBoxW = 64;         // 16 + 32 + 16 samples: one luma block, or two chroma blocks
// Boxes start 16 samples left of the unit (the innermost TMA coordinate must be a multiple of 16 bytes,
// tools/tma_probe.cu).  The k-loops address tiles whose column 0 is the unit's origin - 4 samples:
constexpr int kResCol0 = 3;       // word of that column inside a residual box row
constexpr int kTapCol0 = 4;       // word of the unit's first sample inside a luma-tap box row
constexpr int kLumaBytes = kLumaRows * kBoxW;      // 2240
constexpr int kChromaBytes = kChromaRows * kBoxW;  // 1216
// A pair step may read one row past the box (the unused upper half of a strip's last pair): slots hold one row more.
constexpr int kLumaSlot = 2304;                    // 36 rows
constexpr int kOffTap = 1280, kChromaSlot = 2560;  // residual tile (20 rows) | luma-tap tile (20 rows)
constexpr int kStages = 3;
constexpr int kWarpSmem = 7680;
static_assert(kLumaBytes + kBoxW <= kLumaSlot && kChromaBytes + kBoxW <= kOffTap && kOffTap + kChromaBytes + kBoxW <= kChromaSlot &&
                  kLumaSlot * kStages <= kWarpSmem && kChromaSlot * kStages <= kWarpSmem && kLumaSlot % 128 == 0 &&
                  kOffTap % 128 == 0 && kChromaSlot % 128 == 0,
              "per-warp tile layout");

struct __align__(16) GramSmem {
  uint64_t full[kGramWarps][kStages];  // one mbarrier per warp and stage: the warp's own TMA completions
  uint4 fifo[kGramWarps][4];           // per warp: descriptors of the units requested from the TMA engine, not yet consumed
};

// ------------------------------------------------------------------------------ unit descriptors
//
//   w0  frame (8 bits) | unit column (12: luma block column, chroma pair column) | block row (12)
//   w1  r0: first tile row of the unit's pair steps (6) | pair steps (6) | first unit of its strip (1) | last (1) |
//       the strip's row count n is odd (1) | k range of half 0 lo (6), hi (6)
//   w2  k range of half 1 lo (6), hi (6)            (k = 0..32: observed columns of the unit; luma: both halves the same)
//   w3  observations of the strip (first unit only)
constexpr uint32_t kFirst = 1u << 12, kLast = 1u << 13, kOdd = 1u << 14;

__device__ __forceinline__ void imma_16832(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
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
// Row-pair deduplication.  With tap a = 8q+g (q = cy+3 the tap ROW, g = cx+3 the tap column) the Gram entry of two
// taps only involves the two residual rows s = y+q-3 and s' = y+q'-3 of each observed row y:
//     G[(q,g)][(q',g')] = sum over observed rows y of  D_{s'}[dy][g'][g],     dy = q'-q = s'-s in 0..3,
//     D_{s'}[dy][g'][g] = sum_k mask(k) r(s', k+g'-3) r(s'-dy, k+g-3)          (k = the 32 columns of the item)
// and the x mask (block margins, frame clip, flatness) is the same for every row of a block.  D_{s'}[dy] is
// therefore shared by every q' >= dy: the ten (q, q') row-pair blocks of the Gram need only FOUR products per
// residual row, not ten per observed row.  One k-step takes two consecutive rows a, a+1 as the A operand
// (16 = 2 rows x 8 columns g') and the rows a-3 .. a+1 as five B operands: five m16n8k32 per two rows (2.5 per row
// against 6 per row for the direct form), and every row is fetched from shared memory exactly once.
// The five accumulators RUN over the whole share; what an item contributes to G[(q,.)][(q',.)] is the difference of
// the running sum between the end and the start of the row window [y0+q', y1+q') of q' ("snapshots", a few integer
// adds at the two ends of an item), separately for the rows that sit in the lower (even) and upper (odd) half of
// their pair.  Differences cancel whatever the accumulators held before, so nothing is ever reset, rows above the
// item may be stale registers and rows below it stale shared memory.
//
//   base       : the lane's window word in tile row 0 (g = 7 lanes, chroma: in the luma-tap tile; funnel shift 0)
//   sh         : funnel shift in bits ( = 8 * ((g+1) & 3) )
//   mx[h]      : byte mask of the observed pixels of half h (x margins, frame clip, block not flat), applied to A only
// Tile rows: row r <-> plane row (item origin - 3 + r); observed rows y in [y0, y1) are tile rows [y0+3, y1+3), the
// rows entering the products are tile rows [y0, y1+3).

struct Ring {
  uint32_t B[6][2];  // B operand pairs (unmasked windows) of the six most recent rows, slot = row position mod 6
};

constexpr int kRowWords = 16;  // 64-byte box rows

// Pair step S (mod 3): rows a, a+1 -> ring slots 2S, 2S+1; products with rows a-3 .. a+1.
template <int S>
__device__ __forceinline__ void pair_step(Ring &w, const uint32_t *__restrict__ p, int sh, const uint32_t (&mx)[2],
                                          int (&acc)[5][4]) {
  const uint32_t r00 = __funnelshift_r(p[0], p[1], sh), r01 = __funnelshift_r(p[4], p[5], sh);
  const uint32_t r10 = __funnelshift_r(p[kRowWords], p[kRowWords + 1], sh);
  const uint32_t r11 = __funnelshift_r(p[kRowWords + 4], p[kRowWords + 5], sh);
  w.B[(2 * S) % 6][0] = r00, w.B[(2 * S) % 6][1] = r01;
  w.B[(2 * S + 1) % 6][0] = r10, w.B[(2 * S + 1) % 6][1] = r11;
  const uint32_t q[4] = {r00 & mx[0], r10 & mx[0], r01 & mx[1], r11 & mx[1]};
  imma_16832(acc[0], q, w.B[(2 * S + 3) % 6]);  // row a-3: dy 3 (lower row) | 4 (upper row, unused)
  imma_16832(acc[1], q, w.B[(2 * S + 4) % 6]);  // row a-2: dy 2 | 3
  imma_16832(acc[2], q, w.B[(2 * S + 5) % 6]);  // row a-1: dy 1 | 2
  imma_16832(acc[3], q, w.B[(2 * S) % 6]);      // row a  : dy 0 | 1
  imma_16832(acc[4], q, w.B[(2 * S + 1) % 6]);  // row a+1: (unused) | dy 0
}

// Gacc index of (q', dy), dy <= q'.
__host__ __device__ constexpr int gidx(int q, int dy) { return q * (q + 1) / 2 + dy; }

// Adds (sign = +1) or subtracts the running sums to / from the item accumulators of the tap rows q' named by `ev`
// (bit q': rows in the lower half of their pair, bit 4+q': upper half).  Unsigned arithmetic: the running sums wrap.
__device__ __forceinline__ void snapshot(uint32_t (&G)[10][2], const int (&acc)[5][4], uint32_t ev, bool add) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (ev & (1u << q)) {
#pragma unroll
      for (int dy = 0; dy <= q; ++dy) {
        const uint32_t v0 = (uint32_t)acc[3 - dy][0], v1 = (uint32_t)acc[3 - dy][1];
        G[gidx(q, dy)][0] += add ? v0 : 0u - v0;
        G[gidx(q, dy)][1] += add ? v1 : 0u - v1;
      }
    }
    if (ev & (16u << q)) {
#pragma unroll
      for (int dy = 0; dy <= q; ++dy) {
        const uint32_t v0 = (uint32_t)acc[4 - dy][2], v1 = (uint32_t)acc[4 - dy][3];
        G[gidx(q, dy)][0] += add ? v0 : 0u - v0;
        G[gidx(q, dy)][1] += add ? v1 : 0u - v1;
      }
    }
  }
}

// Window bookkeeping of one strip with n observed rows [y0, y0 + n), pair steps i = 0 .. P-1 (pair i = strip rows
// y0+2i, +1 counted from the first tile's row 0), P = (n + 4) / 2 = m + 2 with m = n / 2: tap row q' sums the
// lower-half rows of pairs [ceil(q'/2), ceil((n+q')/2)) and the upper-half rows of pairs
// [ceil((q'-1)/2), ceil((n+q'-1)/2)).  Written out (E = lower half, O = upper half):
//   windows open   before pair 0: E0 O0 O1     after pair 0: E1 E2 O2 O3     after pair 1: E3
//   windows close  n even   after pair m-1: E0 O0 O1    m: E1 E2 O2 O3    m+1: E3
//                  n odd    after pair m-1: O0          m: E0 E1 O1 O2    m+1: E2 E3 O3
// so a strip is: two pair steps with fixed snapshots (in its first unit), hook-free steps, three steps with the closing
// snapshots (in its last unit).  That needs m >= 3: gram_plan_kernel hands shorter windows to the generic kernel.
constexpr uint32_t kEvA = 0x31u, kEvB = 0xC6u, kEvC = 0x08u;          // {E0 O0 O1}, {E1 E2 O2 O3}, {E3}
constexpr uint32_t kOddA = 0x10u, kOddB = 0x63u, kOddC = 0x8Cu;       // {O0}, {E0 E1 O1 O2}, {E2 E3 O3}

// A pair step whose ring phase is only known at run time (the steps outside the unrolled core).
__device__ __forceinline__ void pair_step_rt(int &ph, Ring &w, const uint32_t *__restrict__ p, int sh,
                                             const uint32_t (&mx)[2], int (&acc)[5][4]) {
  if (ph == 0) pair_step<0>(w, p, sh, mx, acc);
  else if (ph == 1) pair_step<1>(w, p, sh, mx, acc);
  else pair_step<2>(w, p, sh, mx, acc);
  ph = ph == 2 ? 0 : ph + 1;
}

// The pair steps of one unit.  p: the lane's window word in the unit's first row; np pair steps; ph: ring phase, carried
// from unit to unit inside a strip.
__device__ __forceinline__ void unit_rows(const uint32_t *__restrict__ p, int sh, int np, bool first, bool last, bool odd,
                                          const uint32_t (&mx)[2], int &ph, Ring &w, int (&acc)[5][4],
                                          uint32_t (&G)[10][2]) {
  if (first) {
    snapshot(G, acc, kEvA, false);
    pair_step<0>(w, p, sh, mx, acc);
    snapshot(G, acc, kEvB, false);
    pair_step<1>(w, p + 2 * kRowWords, sh, mx, acc);
    snapshot(G, acc, kEvC, false);
    p += 4 * kRowWords;
    np -= 2;
    ph = 2;
  }
  int core = np - (last ? 3 : 0);
#pragma unroll 1
  for (; core > 0 && ph != 2; --core, p += 2 * kRowWords) pair_step_rt(ph, w, p, sh, mx, acc);
#pragma unroll 1
  for (; core >= 3; core -= 3, p += 6 * kRowWords) {
    pair_step<2>(w, p, sh, mx, acc);
    pair_step<0>(w, p + 2 * kRowWords, sh, mx, acc);
    pair_step<1>(w, p + 4 * kRowWords, sh, mx, acc);
  }
#pragma unroll 1
  for (; core > 0; --core, p += 2 * kRowWords) pair_step_rt(ph, w, p, sh, mx, acc);
  if (last) {
    pair_step_rt(ph, w, p, sh, mx, acc);  // pair m-1
    if (odd) snapshot(G, acc, kOddA, true);
    else snapshot(G, acc, kEvA, true);
    pair_step_rt(ph, w, p + 2 * kRowWords, sh, mx, acc);  // pair m
    if (odd) snapshot(G, acc, kOddB, true);
    else snapshot(G, acc, kEvB, true);
    pair_step_rt(ph, w, p + 4 * kRowWords, sh, mx, acc);  // pair m+1
    if (odd) snapshot(G, acc, kOddC, true);
    else snapshot(G, acc, kEvC, true);
  }
}

// MMA tap index a = 8q+g  ->  record tap index (0..23 AR taps, 24 chroma's luma tap, 25 centre sample), -1 unused.
// Column g = 7 is the luma-tap tile (chroma; a junk window for luma): only its row q = 3 (cy = 0) is a tap.
__device__ __forceinline__ int record_tap(int a, bool chroma) {
  const int q = a >> 3, g = a & 7;
  if (g == 7) return (chroma && q == 3) ? 24 : -1;
  if (q < 3) return 7 * q + g;
  if (g < 3) return 21 + g;
  if (g == 3) return 25;
  return -1;
}

__device__ __forceinline__ int pair_index(int i, int j) { return i * kTaps - i * (i - 1) / 2 + (j - i); }

// Adds element (a, b) of the tap-by-tap Gram in MMA indices, a <= b (the mirrored element of a dy = 0 block is
// covered by the transposed position of the same block).
__device__ __forceinline__ void emit(unsigned long long *gram, int a, int b, int v, bool chroma) {
  if (a > b || v == 0) return;
  const int ia = record_tap(a, chroma), ib = record_tap(b, chroma);
  if (ia < 0 || ib < 0) return;
  atomicAdd(&gram[pair_index(min(ia, ib), max(ia, ib))], (unsigned long long)(long long)v);
}

// ------------------------------------------------------------------------------ TMA + mbarrier

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// A failed poll backs off with nanosleep so that waiting warps do not take issue slots from the k-loops of the
// others.  Bounded: a transaction-count mismatch would otherwise hang the GPU; ~0.1 s of failed polls traps.
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try(addr, parity)) return;
#pragma unroll 1
  for (uint32_t n = 0; !mbar_try(addr, parity); ++n) {
    __nanosleep(600);
    if (n > (1u << 17)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int x, int y, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}


// ------------------------------------------------------------------------------ gram_plan_kernel
//
// kPlanSlices CTAs per (plane, frame), one thread per block position (unit column, block row).  Whether a block row
// STARTS a strip is a local question (this row's and the previous row's state), so every start is found in parallel;
// the thread that owns a start walks down its own strip (1.6 block rows on the benchmark input, 64 at most), reserves
// the strip's units with one atomicAdd on the plane's unit count and writes them.  Units of a strip are contiguous,
// strips land in whatever order the atomics resolve: the sums are integers, so the order does not matter.
// Strip rules, per half h of the unit (luma has one half that spans all 32 columns):
//   observed columns [xs, x1) and rows [y0, y1) of a block exactly as add_block_observations computes them;
//   a block is ON when it is flat, not flagged for the generic kernel, and that rectangle is not empty;
//   a strip continues into the next block row while both halves keep their state (off, or on with the same columns):
//   the block above an ON block of a continuing strip is flat, so its y0 is 0 and the rows are contiguous;
//   a block row whose two ON halves start on different rows (y0 differs; then it cannot continue a strip) is two
//   single-half strips of its own; strips are cut every kMaxStrip block rows (int32 bound);
//   windows of fewer than kMinRows rows (bottom slivers of the frame) go to the generic kernel.
struct BlockObs {
  bool on;
  int xs, x1, y0;
};

// st[b]: bit 0 = block b is flat, bit 1 = it is flagged for the generic kernel (the CTA's shared-memory copy of the
// frame's flags)
__device__ __forceinline__ BlockObs block_obs(const uint8_t *st, int nbw, int bx, int by, int wb, int pw, int y1) {
  BlockObs o{false, 0, 0, 0};
  if (bx >= nbw) return o;
  const int b = by * nbw + bx;
  if (st[b] != 1) return o;  // not flat, or flagged
  o.xs = (bx > 0 && (st[b - 1] & 1)) ? 0 : kLag;
  o.x1 = min(pw - bx * wb - kLag, (bx + 1 < nbw && (st[b + 1] & 1)) ? wb : wb - kLag);
  o.y0 = (by > 0 && (st[b - nbw] & 1)) ? 0 : kLag;
  o.on = o.x1 > o.xs && y1 > o.y0;
  return o;
}

struct RowState {
  BlockObs a, b;
  int y1, sa, sb;  // sig = xs | x1 << 6 of an ON half, -1 when off
  __device__ bool any() const { return a.on || b.on; }
};

__device__ __forceinline__ RowState row_state(const uint8_t *st, const Geometry &g, bool luma, int col, int by) {
  const int wb = luma ? 32 : 16, hb = wb;
  const int pw = luma ? g.width : g.width >> 1, ph = luma ? g.height : g.height >> 1;
  RowState r;
  r.y1 = min(ph - by * hb, hb);
  r.a = block_obs(st, g.nbw, luma ? col : 2 * col, by, wb, pw, r.y1);
  r.b = luma ? BlockObs{false, 0, 0, 0} : block_obs(st, g.nbw, 2 * col + 1, by, wb, pw, r.y1);
  r.sa = r.a.on ? (r.a.xs | (r.a.x1 << 6)) : -1;
  r.sb = r.b.on ? (r.b.xs | (r.b.x1 << 6)) : -1;
  return r;
}

constexpr int kPlanSlices = 4;

__global__ void __launch_bounds__(256)
gram_plan_kernel(Geometry g, uint8_t *__restrict__ records, RecordLayout rl, uint4 *__restrict__ plan,
                 int *__restrict__ counts, uint8_t *__restrict__ scratch) {
  extern __shared__ uint8_t s_state[];  // nb bytes when the frame's flags fit (else a global scratch copy is used)
  __shared__ unsigned long long s_obs;
  const int c = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
  const bool luma = c == 0;
  uint8_t *rec = records + (size_t)f * rl.bytes;
  const uint8_t *flat = rec + rl.off_flat;
  uint8_t *ovf = rec + rl.off_ovf + (size_t)c * g.nb;
  uint8_t *st = scratch ? scratch + (((size_t)f * 3 + c) * kPlanSlices + blockIdx.z) * g.nb : s_state;
  for (int b = tid; b < g.nb; b += 256) st[b] = (flat[b] ? 1 : 0) | (ovf[b] ? 2 : 0);
  if (tid == 0) s_obs = 0ull;
  __syncthreads();
  const int ncols = luma ? g.nbw : (g.nbw + 1) >> 1;
  const int hb = luma ? 32 : 16, ph = luma ? g.height : g.height >> 1;
  uint4 *out = plan + ((size_t)f * 3 + c) * g.nb;
  int *count = counts + f * 3 + c;
  long long obs = 0;

  // one strip: block rows [b0, b1), first observed row y0 (in the first tile), signatures of the two halves; its units
  // go to dst[0 .. b1 - b0)
  auto close = [&](uint4 *dst, int col, int b0, int b1, int y0, int sa, int sb) {
    int lo0, hi0, lo1, hi1, width;
    if (luma) {
      lo0 = lo1 = sa & 63, hi0 = hi1 = sa >> 6;
      width = hi0 - lo0;
    } else {
      lo0 = sa < 0 ? 0 : (sa & 63), hi0 = sa < 0 ? 0 : (sa >> 6);
      lo1 = sb < 0 ? 0 : 16 + (sb & 63), hi1 = sb < 0 ? 0 : 16 + (sb >> 6);
      width = (hi0 - lo0) + (hi1 - lo1);
    }
    const int T = b1 - b0;
    const int y1_last = min(ph - (b1 - 1) * hb, hb);
    const int n = hb * (T - 1) + y1_last - y0;
    const int P = (n + 4) >> 1;
    obs += (long long)n * width;
    int cursor = y0, done = 0;  // next strip row, in rows from the first tile's row 0; pair steps emitted
    for (int k = 0; k < T; ++k) {
      const int r0 = cursor - hb * k;
      const int np = k + 1 < T ? (hb + 3 - r0) >> 1 : P - done;
      const uint32_t w0 = (uint32_t)f | ((uint32_t)col << 8) | ((uint32_t)(b0 + k) << 20);
      const uint32_t w1 = (uint32_t)r0 | ((uint32_t)np << 6) | (k == 0 ? kFirst : 0u) | (k + 1 == T ? kLast : 0u) |
                          ((n & 1) ? kOdd : 0u) | ((uint32_t)lo0 << 16) | ((uint32_t)hi0 << 22);
      dst[k] = make_uint4(w0, w1, (uint32_t)lo1 | ((uint32_t)hi1 << 6), k == 0 ? (uint32_t)(n * width) : 0u);
      cursor += 2 * np;
      done += np;
    }
  };
  auto sliver = [&](int bx, int by) {
    ovf[by * g.nbw + bx] = 1;  // read again only by the generic kernel, launched after the Gram kernel
    atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_ovf_count), 1ull);
  };

  const int items = ncols * g.nbh, lane = tid & 31;
  for (int it0 = blockIdx.z * 256 + (tid & ~31); it0 < items; it0 += 256 * kPlanSlices) {  // warp-uniform trip count
    const int it = it0 + lane;
    // what this lane's block position starts: nothing, one strip of T block rows, or two single-row strips
    int T = 0, col = 0, by = 0, e = 0;
    bool split = false;
    RowState r{};
    if (it < items) {
      by = it / ncols, col = it - by * ncols;
      r = row_state(st, g, luma, col, by);
      bool start = r.any() && (by == 0 || r.y1 < kMinRows || by % kMaxStrip == 0);
      if (r.any() && !start) {
        const RowState p = row_state(st, g, luma, col, by - 1);
        start = !p.any() || p.sa != r.sa || p.sb != r.sb || (p.a.on && p.b.on && p.a.y0 != p.b.y0);
      }
      if (start) {
        // halves whose window would be too short go to the generic kernel
        if (r.a.on && r.y1 - r.a.y0 < kMinRows) sliver(luma ? col : 2 * col, by), r.a.on = false, r.sa = -1;
        if (r.b.on && r.y1 - r.b.y0 < kMinRows) sliver(2 * col + 1, by), r.b.on = false, r.sb = -1;
        if (r.any()) {
          split = r.a.on && r.b.on && r.a.y0 != r.b.y0;  // two single-half strips of one block row
          if (split) {
            T = 2;
          } else {
            for (e = by + 1; e < g.nbh && e % kMaxStrip != 0; ++e) {
              const RowState n = row_state(st, g, luma, col, e);
              if (!n.any() || n.sa != r.sa || n.sb != r.sb || n.y1 < kMinRows) break;
            }
            T = e - by;
          }
        }
      }
    }
    // one reservation per warp: inclusive scan of the unit counts, one atomicAdd on the plane's counter
    int incl = T;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && warp_total) base = atomicAdd(count, warp_total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (T) {
      uint4 *dst = out + base + incl - T;
      if (split) {
        close(dst, col, by, by + 1, r.a.y0, r.sa, -1);
        close(dst + 1, col, by, by + 1, r.b.y0, -1, r.sb);
      } else {
        close(dst, col, by, e, r.a.on ? r.a.y0 : r.b.y0, r.sa, r.sb);
      }
    }
  }
  // observations: warp sums, then one shared-memory add per warp
  {
    unsigned long long v = (unsigned long long)obs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v) atomicAdd(&s_obs, v);
  }
  __syncthreads();
  if (tid == 0 && s_obs) atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_nobs) + c, s_obs);
}

// ------------------------------------------------------------------------------ gram_imma_kernel
//
// Work distribution: the units of the batch (all planes) are cut into chunks of kChunk consecutive units of one plane;
// warps take chunks from a global counter (luma chunks first: the long ones).  Any warp can run any plane -- the k-loop
// is the same, only the tile geometry differs -- so there is no luma / chroma split to balance, and a CTA that starts
// late (another stream's kernel was on its SM) simply takes fewer chunks.  A chunk is moved forward to strip boundaries
// at both ends (a strip's running sums live in one warp's registers).  The PRODUCER side of a warp (descriptor fetch,
// TMA requests) rolls from one chunk into the next on its own, one chunk claimed ahead, so the consumer side sees one
// uninterrupted stream of units: chunk boundaries cost no pipeline drain, and chunks can be small.
#ifndef G1S_CHUNK
#define G1S_CHUNK 16
#endif
constexpr int kChunk = G1S_CHUNK;
constexpr int kSlot = kChromaSlot;  // stage stride of a warp's tile ring, either plane
constexpr uint32_t kPlaneShift = 12;  // fifo copy of a descriptor: plane in w2 bits 12..13
static_assert(kLumaSlot <= kSlot && kSlot * kStages <= kWarpSmem, "stage stride");

__global__ void __launch_bounds__(kGramThreads, 2)
gram_imma_kernel(Geometry g, uint8_t *__restrict__ records, RecordLayout rl, int nframes, const uint8_t *__restrict__ tmaps,
                 const uint4 *__restrict__ plan, int *__restrict__ counts) {
  extern __shared__ __align__(128) uint8_t tiles[];
  __shared__ GramSmem sm;
  __shared__ int s_incl[3][65];  // [p][f + 1]: units of plane p in frames 0 .. f (a batch has at most 64 frames)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;  // mma "groupID" (= cx + 3) and thread-in-group

  // ---- unit counts of the batch as prefix sums over the frames (the only CTA-wide step of the kernel)
  if (warp < 3) {
    const int p = warp;
    int a = (p < g.planes && lane < nframes) ? counts[lane * 3 + p] : 0;
    int b = (p < g.planes && lane + 32 < nframes) ? counts[(lane + 32) * 3 + p] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int va = __shfl_up_sync(0xffffffffu, a, o), vb = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= o) a += va, b += vb;
    }
    const int ta = __shfl_sync(0xffffffffu, a, 31);
    s_incl[p][lane + 1] = a, s_incl[p][lane + 33] = ta + b;
    if (lane == 0) s_incl[p][0] = 0;
  }
  __syncthreads();
  const int nch0 = (s_incl[0][64] + kChunk - 1) / kChunk, nch1 = (s_incl[1][64] + kChunk - 1) / kChunk;
  const int nchunks = nch0 + nch1 + (s_incl[2][64] + kChunk - 1) / kChunk;
  int *const work = counts + 3 * nframes;  // the chunk counter (zeroed with the counts)

  uint8_t *const my_tiles = tiles + warp * kWarpSmem;
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&sm.full[warp][s], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  int acc[5][4];        // running sums of the five row products (never reset, see the k-loop notes)
  uint32_t G[10][2];    // Gram blocks (q', dy) of the current frame and plane since the last flush
  Ring ring;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[i][r] = 0;
#pragma unroll
  for (int i = 0; i < 10; ++i) G[i][0] = G[i][1] = 0u;
#pragma unroll
  for (int i = 0; i < 6; ++i) ring.B[i][0] = ring.B[i][1] = 0u;

  int sh = 8 * ((gq + 1) & 3);
  asm volatile("" : "+r"(sh));  // one register instead of four instructions per funnel shift group
  const int dxw = (gq + 1) >> 2;
  // the lane's window word in row 0 of a stage's tile: g = 7 lanes of a chroma unit walk the luma-tap tile
  const int off_luma = 4 * (kResCol0 + t + dxw);
  const int off_chroma = gq == 7 ? kOffTap + 4 * (kTapCol0 + t) : off_luma;

  // Gram blocks -> the frame's int64 record, straight from registers: element r of block (q', dy) in lane (gq, t)
  // is the product of the later tap (q', g' = gq) with the earlier tap (q' - dy, g = 2t + r).
  auto flush = [&](int f, int plane) {
    uint8_t *rec = records + (size_t)f * rl.bytes;
    unsigned long long *gram = reinterpret_cast<unsigned long long *>(rec + rl.off_gram) + (size_t)plane * kPairs;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int dy = 0; dy <= q; ++dy)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int v = (int)G[gidx(q, dy)][r];
          G[gidx(q, dy)][r] = 0u;
          emit(gram, 8 * (q - dy) + 2 * t + r, 8 * q + gq, v, plane != 0);
        }
  };

  // ---- producer side.  Position: unit u of the plane's list (all frames), which is unit u - pbeg of frame pf;
  // `more`: dn is the descriptor of a unit this warp still has to request.
  int nextc = 0;  // lane 0: the chunk claimed ahead
  if (lane == 0) nextc = atomicAdd(work, 1);
  int pplane = 0, pf = 0, pbeg = 0, pend = 0, u = 0, u_hi = 0, tot = 0;
  bool more = false, done = false;
  uint4 dn = make_uint4(0u, 0u, 0u, 0u);
  auto unit_at = [&](int f, int i) { return __ldg(plan + ((size_t)f * 3 + pplane) * g.nb + i); };
  auto next_chunk = [&]() {  // to the first strip start of the next chunk that has one; false when the work is out
    for (;;) {
      const int c = __shfl_sync(0xffffffffu, nextc, 0);
      if (c >= nchunks) {
        done = true;
        return false;
      }
      if (lane == 0) nextc = atomicAdd(work, 1);
      pplane = c < nch0 ? 0 : (c < nch0 + nch1 ? 1 : 2);
      const int *incl = s_incl[pplane];
      tot = incl[64];
      u = (c - (pplane == 0 ? 0 : (pplane == 1 ? nch0 : nch0 + nch1))) * kChunk;
      u_hi = min(tot, u + kChunk);
      // frame that holds unit u: the first whose inclusive prefix exceeds it
      const uint32_t in_lo = __ballot_sync(0xffffffffu, incl[lane + 1] > u);
      const uint32_t in_hi = __ballot_sync(0xffffffffu, incl[lane + 33] > u);
      pf = in_lo ? __ffs(in_lo) - 1 : 32 + __ffs(in_hi) - 1;
      pbeg = incl[pf], pend = incl[pf + 1];
      // a strip that started in the previous chunk belongs to that chunk's warp: 32 descriptors per look
      while (u < u_hi) {
        const int lim = min(pend, u_hi), i = u + lane;
        uint32_t w1 = 0u;
        if (i < lim) w1 = __ldg(reinterpret_cast<const uint32_t *>(plan + ((size_t)pf * 3 + pplane) * g.nb + (i - pbeg)) + 1);
        const uint32_t m = __ballot_sync(0xffffffffu, (w1 & kFirst) != 0u);
        if (m) {
          u += __ffs(m) - 1;
          dn = unit_at(pf, u - pbeg);
          more = true;
          return true;
        }
        u = min(u + 32, lim);
        while (u >= pend && u < u_hi) pbeg = pend, ++pf, pend = incl[pf + 1];
      }
    }
  };

  int head = 0, tail = 0;       // fifo positions (units requested / consumed)
  int hstage = 0, tstage = 0;   // their stages ( = position % kStages, kept incrementally)
  uint32_t phases = 0;          // bit s: parity to wait for on stage s
  auto produce = [&]() {        // false when the work is out
    if (!more && (done || !next_chunk())) return false;
    const uint4 d = dn;
    const int stage = hstage;
    if (++hstage == kStages) hstage = 0;
    if (lane == 0) {
      sm.fifo[warp][head & 3] = make_uint4(d.x, d.y, d.z | ((uint32_t)pplane << kPlaneShift), d.w);
      const int f = d.x & 255, col = (d.x >> 8) & 4095, by = d.x >> 20;
      const uint8_t *fmaps = tmaps + (size_t)f * kResidualMaps * 128;
      uint8_t *st = my_tiles + stage * kSlot;
      uint64_t *bar = &sm.full[warp][stage];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this warp's reads of the stage precede the refill
      if (pplane == 0) {
        mbar_expect_tx(bar, kLumaBytes);
        tma_load_2d(st, fmaps, 32 * col - 16, 32 * by - 3, bar);
      } else {
        const int cx = 32 * col - 16, cy = 16 * by - 3;
        mbar_expect_tx(bar, 2 * kChromaBytes);
        tma_load_2d(st, fmaps + pplane * 128, cx, cy, bar);
        tma_load_2d(st + kOffTap, fmaps + 3 * 128, cx, cy, bar);
      }
    }
    ++head;
    // the next unit of the chunk: the units up to u_hi, then the rest of the strip that crosses u_hi
    more = false;
    if (++u < tot) {
      while (u >= pend) pbeg = pend, ++pf, pend = s_incl[pplane][pf + 1];
      dn = unit_at(pf, u - pbeg);
      more = u < u_hi || !(dn.y & kFirst);
    }
    return true;
  };

  int cf = -1, cplane = 0;  // frame and plane the strip accumulators belong to
  int since = 0;            // observations accumulated since the last flush
  int ph = 0;               // ring phase, carried from unit to unit inside a strip
  for (;;) {
    __syncwarp();  // every lane is done with the stage about to be refilled
    while (head - tail < kStages && produce()) {
    }
    if (head == tail) break;
    __syncwarp();
    const uint4 d = sm.fifo[warp][tail & 3];
    const int stage = tstage;
    if (++tstage == kStages) tstage = 0;
    ++tail;
    const bool first = d.y & kFirst;
    const int plane = (d.z >> kPlaneShift) & 3;
    if (first) {  // G only changes at strip ends, so strip starts are the flush points
      const int f = d.x & 255;
      if (f != cf || plane != cplane || since + (int)d.w > kMaxObs) {
        if (cf >= 0) flush(cf, cplane);
        cf = f, cplane = plane;
        since = 0;
      }
      since += (int)d.w;
    }
    mbar_wait(&sm.full[warp][stage], (phases >> stage) & 1u);
    phases ^= 1u << stage;
    const int r0 = d.y & 63, np = (d.y >> 6) & 63;
    const int lo0 = (d.y >> 16) & 63, hi0 = (d.y >> 22) & 63, lo1 = d.z & 63, hi1 = (d.z >> 6) & 63;
    const uint32_t mx[2] = {byte_mask(4 * t, lo0, hi0), byte_mask(16 + 4 * t, lo1, hi1)};
    const uint32_t *p =
        reinterpret_cast<const uint32_t *>(my_tiles + stage * kSlot + (plane == 0 ? off_luma : off_chroma)) + r0 * kRowWords;
    unit_rows(p, sh, np, first, d.y & kLast, d.y & kOdd, mx, ph, ring, acc, G);
  }
  if (cf >= 0) flush(cf, cplane);
}

}  // namespace

bool gram_imma_supported(const Geometry &g) {
  const bool shape = (g.planes == 1) || (g.planes == 3 && g.ss_x == 1 && g.ss_y == 1);
  return shape && g.width >= 8 && g.height >= 8 && g.nbw <= 4095 && g.nbh <= 4095;
}

void gram_imma_tma_boxes(int box[kResidualMaps][2]) {
  box[0][0] = kBoxW, box[0][1] = kLumaRows;
  for (int k = 1; k < kResidualMaps; ++k) box[k][0] = kBoxW, box[k][1] = kChromaRows;
}

// descriptors, then (only for frames of more than kPlanSmemBlocks blocks) one state byte per block, plane and slice
constexpr int kPlanSmemBlocks = 40960;
size_t gram_plan_bytes(int nframes, const Geometry &g) {
  return (size_t)nframes * 3 * g.nb * sizeof(uint4) +
         (g.nb > kPlanSmemBlocks ? (size_t)nframes * 3 * kPlanSlices * g.nb : 0);
}

void launch_gram_plan(int nframes, const Geometry &g, uint8_t *records, const RecordLayout &rl, void *plan, int *counts,
                      cudaStream_t st) {
  const bool in_smem = g.nb <= kPlanSmemBlocks;
  uint8_t *scratch = in_smem ? nullptr : static_cast<uint8_t *>(plan) + (size_t)nframes * 3 * g.nb * sizeof(uint4);
  // the plan kernel reserves units with atomicAdd; counts[3 * nframes] is the Gram kernel's chunk counter
  cudaMemsetAsync(counts, 0, sizeof(int) * (3 * nframes + 1), st);
  gram_plan_kernel<<<dim3(g.planes, nframes, kPlanSlices), 256, in_smem ? (size_t)g.nb : 0, st>>>(
      g, records, rl, static_cast<uint4 *>(plan), counts, scratch);
}

void launch_gram_imma(int nframes, const Geometry &g, uint8_t *records, const RecordLayout &rl, const void *tmaps,
                      const void *plan, int *counts, cudaStream_t st) {
  const int smem = kGramWarps * kWarpSmem;
  // resident CTAs of the current device (the persistent grid is one wave); cached per device, and the
  // dynamic-shared-memory attribute is a per-device property of the function as well
  static int slots_of[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int &slots = slots_of[dev & 63];
  if (slots == 0) {
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(gram_imma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gram_imma_kernel, kGramThreads, smem);
    if (const char *e = std::getenv("G1S_GRAM_OCC")) per_sm = std::min(per_sm, std::max(1, std::atoi(e)));  // tuning aid
    slots = sms * std::max(per_sm, 1);
  }
  // one wave of persistent CTAs whatever the batch size (a warp flushes its int32 strip accumulators as it goes)
  const long long blocks = (long long)nframes * g.nb;
  const int grid = (int)std::min<long long>(slots, std::max<long long>(1, blocks / 8));
  gram_imma_kernel<<<grid, kGramThreads, smem, st>>>(g, records, rl, nframes, static_cast<const uint8_t *>(tmaps),
                                                     static_cast<const uint4 *>(plan), counts);
}

}  // namespace g1s
