// gram_imma_kernel — the AR normal-equation (Gram) accumulation on the int8 tensor-core path
// (mma.sync.m16n8k32.s8, SASS IMMA.16832.S8.S8), 4:2:0 and monochrome.
//
// Replaces NoiseModel::add_block_observations (extract_ar_row + the n x n outer-product accumulation) of
// av1-grain's diff module (reached from /root/reference/src/main.rs:442) for every flat block whose
// residuals fit in int8; the rare block that does not was flagged by residual_kernel and is redone
// exactly by gram_generic_kernel.  All sums are integers, so the result is bit-identical to the oracle
// whatever the summation order.
//
// The kernel never touches the frames: residual_kernel left the s8 residual of Y / Cb / Cr and the two
// halves of chroma's luma tap in engine-owned planes, and every warp pulls the tiles of its own work items
// into its own slice of shared memory with the TMA engine (cp.async.bulk.tensor.2d, SASS UTMALDG; frame
// edges are zero-filled by the hardware, so there is no edge path), one or two items ahead of its k-loop.
//
// Why warps are autonomous.  Measured on the B200 (tools/imma_probe.cu, profiles/): the legacy IMMA path
// holds a sub-partition's issue port for its whole 8.4 cycles, so a sub-partition's time is
// 8.4 * IMMAs + (every other warp instruction it issues) -- nothing overlaps, polls and barrier spins are
// paid in full.  So: no CTA barriers, no shared rings, no inter-warp waits.  A warp has a plane for life
// (7 luma + 5 chroma warps per CTA, the measured 3:1:1 cost ratio of Y:Cb:Cr), an equal contiguous share
// of that plane's blocks over the whole batch, its own mbarriers, and it adds its int32 accumulators to
// the frame's int64 record directly when its share leaves a frame (about 330 atomics, once or twice per
// warp per launch).
//   work item   luma: one 32x32 block (32 k-steps); chroma: two adjacent 16x16 blocks (16 k-steps)
//   One k-step = 32 pixels of one row.  X[k][a] = residual at pixel k shifted by tap a;
//   D += X^T X over the upper block-triangle (6 m16n8k32 MMAs).  Tap a = 8q+g with g = cx+3 (lane
//   group), q = cy+3: a thread's four taps are the SAME column offset on four consecutive rows, so its
//   operands slide down one row per k-step: one new 32-bit window (two LDS + funnel shift) per half is
//   fetched, the B operand of a row is a register pair and the A operand of two consecutive rows a
//   register quad that serves as the lower m-tile now and as the upper m-tile two steps later.
//   The observation mask (block margins, frame clipping) is a byte mask on k applied to the operand
//   words (mask^2 = mask, so masking both A and B is exact).  Chroma's luma tap rides in lane group 7.
//   int32 accumulators are bounded by the share: <= 96 blocks * 32 k-steps * 32 * 2^14 < 2^31.
#include "g1s_kernels.h"

#include <algorithm>
#include <cstdlib>

namespace g1s {

namespace {

#ifndef G1S_GRAM_WARPS
#define G1S_GRAM_WARPS 12
#define G1S_LUMA_WARPS 7
#endif
constexpr int kGramWarps = G1S_GRAM_WARPS;
constexpr int kGramThreads = 32 * kGramWarps;
constexpr int kLumaWarps = G1S_LUMA_WARPS;  // per CTA; the others are chroma warps (measured cost ratio Y : Cb : Cr about 3 : 1 : 1)
constexpr int kMaxShare = 96;     // luma blocks per warp and launch: bounds the int32 accumulators
constexpr int kWin = 30;          // blocks per flag window: one ballot holds blocks bx0-1 .. bx0+30
constexpr int kPL = 16;           // tile pitch in 32-bit words (64-byte box rows), luma
constexpr int kPC = 16;           // the same for chroma and the luma-tap tiles
constexpr int kLumaRows = 35;     // 3 halo rows + 32
constexpr int kChromaRows = 19;   // 3 halo rows + 16
constexpr int kLoRows = 20;       // the lo tile starts one row higher (row -1 of the first step is read, never used)
constexpr int kBoxW = 64;         // 16 + 32 + 16 samples: one luma block, or two chroma blocks
// Boxes start 16 samples left of the item (the innermost TMA coordinate must be a multiple of 16 bytes,
// tools/tma_probe.cu).  The k-loops address tiles whose column 0 is the item's origin - 4 samples:
constexpr int kResCol0 = 3;       // word of that column inside a residual box row
constexpr int kTapCol0 = 4;       // word of the item's first sample inside a hi / lo box row
constexpr int kLumaBytes = kLumaRows * kBoxW;      // 2240
constexpr int kChromaBytes = kChromaRows * kBoxW;  // 1216
constexpr int kLoBytes = kLoRows * kBoxW;          // 1280
constexpr int kLumaSlot = 2304, kLumaStages = 3;   // per luma warp: 3 x 2304 = 6912 bytes
constexpr int kOffHi = 1280, kOffLo = 2560, kChromaSlot = 3840, kChromaStages = 2;  // per chroma warp: 2 x 3840
constexpr int kWarpSmem = 7680;
constexpr int kMaxStages = 3;
static_assert(kLumaBytes <= kLumaSlot && kChromaBytes <= kOffHi && kLumaSlot * kLumaStages <= kWarpSmem &&
                  kChromaSlot * kChromaStages <= kWarpSmem && kLumaSlot % 128 == 0 && kOffHi % 128 == 0,
              "per-warp tile layout");

struct __align__(16) GramSmem {
  uint64_t full[kGramWarps][kMaxStages];  // one mbarrier per warp and stage: the warp's own TMA completions
  int4 fifo[kGramWarps][4];               // per warp: items requested from the TMA engine, not yet consumed
};

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
#pragma unroll 1
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
// group g = 7: those lanes step through the hi tile instead of the residual tile (same pitch, funnel shift 0), so
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
                                       int sh, bool is7, const uint32_t (&mx)[2], int (&acc)[6][4], int &ll) {
  fetch_row_c(pnew, lp, sh, is7, mx, w.P[(K + 3) & 3], w.Q[(K + 2) & 3], w.Q[(K + 3) & 3]);
  // g = 7 lanes: the upper row of this step's quad is the masked lo part of the luma tap of the observed
  // row; lo * lo is the one self product the tiles do not deliver (other lanes: ignored)
  ll = __dp4a((int)w.Q[K & 3][1], (int)w.Q[K & 3][1], ll);
  ll = __dp4a((int)w.Q[K & 3][3], (int)w.Q[K & 3][3], ll);
  imma_16832(acc[0], w.Q[K & 3], w.P[K & 3]);
  imma_16832(acc[1], w.Q[K & 3], w.P[(K + 1) & 3]);
  imma_16832(acc[2], w.Q[K & 3], w.P[(K + 2) & 3]);
  imma_16832(acc[3], w.Q[K & 3], w.P[(K + 3) & 3]);
  imma_16832(acc[4], w.Q[(K + 2) & 3], w.P[(K + 2) & 3]);
  imma_16832(acc[5], w.Q[(K + 2) & 3], w.P[(K + 3) & 3]);
}

// p0 / lp0: the lane's words for tile row ys (lp0 already points one ls row above it).
__device__ __forceinline__ void chroma_rows(const uint32_t *__restrict__ p0, const uint32_t *__restrict__ lp0, int sh,
                                            bool is7, int nrows, const uint32_t (&mx)[2], int (&acc)[6][4], int &ll) {
  Window w;
  fetch_row_c(p0, lp0, sh, is7, mx, w.P[0], w.Q[3], w.Q[0]);
  fetch_row_c(p0 + kPC, lp0 + kPC, sh, is7, mx, w.P[1], w.Q[0], w.Q[1]);
  fetch_row_c(p0 + 2 * kPC, lp0 + 2 * kPC, sh, is7, mx, w.P[2], w.Q[1], w.Q[2]);
  const uint32_t *p = p0 + 3 * kPC, *lp = lp0 + 3 * kPC;
  int n = nrows;
#pragma unroll 1
  for (; n >= 4; n -= 4, p += 4 * kPC, lp += 4 * kPC) {
    step6c<0>(w, p, lp, sh, is7, mx, acc, ll);
    step6c<1>(w, p + kPC, lp + kPC, sh, is7, mx, acc, ll);
    step6c<2>(w, p + 2 * kPC, lp + 2 * kPC, sh, is7, mx, acc, ll);
    step6c<3>(w, p + 3 * kPC, lp + 3 * kPC, sh, is7, mx, acc, ll);
  }
  if (n > 0) step6c<0>(w, p, lp, sh, is7, mx, acc, ll);
  if (n > 1) step6c<1>(w, p + kPC, lp + kPC, sh, is7, mx, acc, ll);
  if (n > 2) step6c<2>(w, p + 2 * kPC, lp + 2 * kPC, sh, is7, mx, acc, ll);
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


__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MODE 0: the product.  MODE 1 / 2 are measurement aids (G1S_GRAM_MODE, wrong results by design):
// 1 = TMA traffic and waits only, no k-loops; 2 = k-loops on whatever is in shared memory, no TMA traffic.
template <int MODE>
__global__ void __launch_bounds__(kGramThreads, 2)
gram_imma_kernel(Geometry g, uint8_t *__restrict__ records, RecordLayout rl, int nframes,
                 const uint8_t *__restrict__ tmaps) {
  extern __shared__ __align__(128) uint8_t tiles[];
  __shared__ GramSmem sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;  // mma "groupID" (= cx + 3) and thread-in-group
  const bool has_chroma = g.planes == 3;
  const int W = g.width, H = g.height, pw = W >> 1, ph = H >> 1;

  // ---- this warp's plane and its share of the plane's items
  const int cta = blockIdx.x, ncta = gridDim.x;
  int plane, rank, nranks;
  if (!has_chroma) {
    plane = 0, rank = cta * kGramWarps + warp, nranks = ncta * kGramWarps;
  } else if (warp < kLumaWarps) {
    plane = 0, rank = cta * kLumaWarps + warp, nranks = ncta * kLumaWarps;
  } else {
    // chroma warps split evenly between Cb and Cr; with an odd count the extra warp alternates with the CTA's parity
    constexpr int kC = kGramWarps - kLumaWarps, kHi = (kC + 1) / 2, kLo = kC / 2;
    const int k = warp - kLumaWarps, ncb = (cta & 1) ? kLo : kHi;
    const int even = (ncta + 1) >> 1, odd = ncta >> 1, e_before = (cta + 1) >> 1, o_before = cta >> 1;
    if (k < ncb) plane = 1, rank = kHi * e_before + kLo * o_before + k, nranks = kHi * even + kLo * odd;
    else plane = 2, rank = kLo * e_before + kHi * o_before + (k - ncb), nranks = kLo * even + kHi * odd;
  }
  const bool luma = plane == 0;
  const int per_row = luma ? g.nbw : (g.nbw + 1) >> 1;  // items per block row: blocks, or chroma block pairs
  const int per_win = luma ? kWin : kWin / 2;
  const long long total = (long long)nframes * g.nbh * per_row;
  const int i_lo = (int)(total * rank / nranks), i_hi = (int)(total * (rank + 1) / nranks);
  if (i_lo >= i_hi) return;  // whole warp; nothing below synchronises across warps

  uint8_t *const my_tiles = tiles + warp * kWarpSmem;
  const int nstages = luma ? kLumaStages : kChromaStages;
  const int slot = luma ? kLumaSlot : kChromaSlot;
  if (lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) mbar_init(&sm.full[warp][s], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  int acc[6][4];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[i][r] = 0;
  int ll = 0;    // chroma, g = 7 lanes: sum lo*lo of the luma tap over observed pixels
  int nobs = 0;  // observations of the items accumulated since the last flush

  const int sh = 8 * ((gq + 1) & 3);
  const int dxw = (gq + 1) >> 2;
  const bool is7 = gq == 7;

  // Accumulators -> the frame's int64 record, straight from registers (every (lane, tile, element) owns one tap pair).
  auto flush = [&](int f) {
    uint8_t *rec = records + (size_t)f * rl.bytes;
    unsigned long long *gram = reinterpret_cast<unsigned long long *>(rec + rl.off_gram) + (size_t)plane * kPairs;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int mrow = i >= 4 ? 16 : 0;
      const int ncol = i >= 4 ? 8 * (i - 2) : 8 * i;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int v = acc[i][r];
        acc[i][r] = 0;
        if (v == 0) continue;
        const int b = ncol + 2 * t + (r & 1);
        if (plane > 0 && is7 && i < 4) {
          // rows 7 / 15 of the lower m-tile: luma tap hi / lo
          if (b == 7)  // column 7 of the first n-tile is hi again: (8h + l)^2 = 64 hh + 16 hl + ll
            atomicAdd(&gram[pair_index(24, 24)], (unsigned long long)((long long)v * ((r >> 1) ? 16 : 64)));
          else
            emit_luma_tap(gram, b, v, (r >> 1) ? 1 : 8);
        } else {
          emit(gram, mrow + gq + 8 * (r >> 1), b, v);
        }
      }
    }
    if (plane > 0) {
      int v = is7 ? ll : 0;
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (lane == 28 && v) atomicAdd(&gram[pair_index(24, 24)], (unsigned long long)(long long)v);
      ll = 0;
    }
    if (lane == 0 && nobs)
      atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_nobs) + plane, (unsigned long long)(long long)nobs);
    nobs = 0;
  };

  // ---- producer side: scan the share for items with work, request their tiles, queue them
  // position of the scan: frame, block row, first item of the current flag window, and the window's ballots
  int pf, pby, pwx;
  {
    const int per_frame = g.nbh * per_row;
    pf = i_lo / per_frame;
    const int r = i_lo - pf * per_frame;
    pby = r / per_row;
    pwx = ((r - pby * per_row) / per_win) * per_win;
  }
  int pidx = ((pf * g.nbh + pby) * per_row) + pwx;  // global index of the window's first item
  uint32_t m_flat = 0, m_up = 0, m_ovf = 0, m_todo = 0;  // bit l <-> block (window's first block - 1 + l); todo: items left in the window
  auto load_window = [&]() {
    const uint8_t *rec = records + (size_t)pf * rl.bytes;
    const int bx = (luma ? pwx : 2 * pwx) - 1 + lane;
    const bool in = bx >= 0 && bx < g.nbw;
    const int b = pby * g.nbw + bx;
    const uint8_t fl = in ? (rec + rl.off_flat)[b] : 0;
    const uint8_t up = (in && pby > 0) ? (rec + rl.off_flat)[b - g.nbw] : 0;
    const uint8_t ov = in ? (rec + rl.off_ovf)[(size_t)plane * g.nb + b] : 0;
    m_flat = __ballot_sync(0xffffffffu, fl != 0);
    m_up = __ballot_sync(0xffffffffu, up != 0);
    m_ovf = __ballot_sync(0xffffffffu, ov != 0);
    // items of the window that lie in the share, in the row, and have at least one flat, non-overflowed block
    const uint32_t ok = m_flat & ~m_ovf;
    uint32_t items = luma ? (ok >> 1) & ((1u << kWin) - 1u) : 0u;
    if (!luma)
      for (int k = 0; k < kWin / 2; ++k)
        if ((ok >> (2 * k + 1)) & 3u) items |= 1u << k;
    const int first = max(i_lo - pidx, 0), last = min(min(i_hi - pidx, per_row - pwx), per_win);  // [first, last)
    const uint32_t keep = last > first ? (((last >= 32 ? 0u : (1u << last)) - 1u) & ~((1u << first) - 1u)) : 0u;
    m_todo = items & keep;
  };
  bool scan_done = false;
  auto next_window = [&]() {  // advance to the next window of the share; false when the share is exhausted
    pwx += per_win;
    pidx += per_win;
    if (pwx >= per_row) {
      pidx += per_row - pwx;  // the last window of a row is short
      pwx = 0;
      if (++pby == g.nbh) pby = 0, ++pf;
    }
    return pidx < i_hi;
  };
  load_window();

  int head = 0, tail = 0;       // fifo positions (items requested / consumed)
  int hstage = 0, tstage = 0;   // their stages ( = position % nstages, kept incrementally)
  uint32_t phases = 0;          // bit s: parity to wait for on stage s
  auto produce = [&]() {        // request the next item with work; false if there is none left
    while (m_todo == 0) {
      if (scan_done || !next_window()) {
        scan_done = true;
        return false;
      }
      load_window();
    }
    const int k = __ffs(m_todo) - 1;
    m_todo &= m_todo - 1;
    // flags of the item's blocks, from the window ballots (bit l <-> block first - 1 + l)
    uint32_t bits;
    if (luma) {
      const int l = k + 1;
      bits = ((m_flat >> (l - 1)) & 1u) | (((m_flat >> (l + 1)) & 1u) << 1) | (((m_up >> l) & 1u) << 2);
    } else {
      const int l = 2 * k + 1;  // blocks l (half 0) and l + 1 (half 1)
      bits = ((m_flat >> (l - 1)) & 15u)            // flat: left neighbour, A, B, right neighbour
             | (((m_up >> l) & 3u) << 4)             // flat above A, B
             | (((m_ovf >> l) & 3u) << 6);           // overflow A, B
    }
    const int stage = hstage;
    if (++hstage == nstages) hstage = 0;
    if (lane == 0) {
      sm.fifo[warp][head & 3] = make_int4(pf, pby, pwx + k, (int)bits);
      if (MODE != 2) {
        const uint8_t *fmaps = tmaps + (size_t)pf * kResidualMaps * 128;
        uint8_t *st = my_tiles + stage * slot;
        uint64_t *bar = &sm.full[warp][stage];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this warp's reads of the stage precede the refill
        if (luma) {
          mbar_expect_tx(bar, kLumaBytes);
          tma_load_2d(st, fmaps, 32 * (pwx + k) - 16, 32 * pby - 3, bar);
        } else {
          const int cx = 32 * (pwx + k) - 16, cy = 16 * pby;
          mbar_expect_tx(bar, 2 * kChromaBytes + kLoBytes);
          tma_load_2d(st, fmaps + plane * 128, cx, cy - 3, bar);
          tma_load_2d(st + kOffHi, fmaps + 3 * 128, cx, cy, bar);
          tma_load_2d(st + kOffLo, fmaps + 4 * 128, cx, cy - 1, bar);
        }
      }
    }
    ++head;
    return true;
  };

  int cf = -1;  // frame the accumulators belong to
  for (;;) {
    __syncwarp();  // every lane is done with the stage about to be refilled
    while (head - tail < nstages && produce()) {
    }
    if (head == tail) break;
    __syncwarp();
    const int4 item = sm.fifo[warp][tail & 3];
    const int stage = tstage;
    if (++tstage == nstages) tstage = 0;
    ++tail;
    const int f = item.x, by = item.y, ix = item.z;
    const uint32_t bits = (uint32_t)item.w;
    if (f != cf) {
      if (cf >= 0) flush(cf);
      cf = f;
    }
    if (MODE != 2) {
      mbar_wait(&sm.full[warp][stage], (phases >> stage) & 1u);
      phases ^= 1u << stage;
    }
    if (MODE == 1) continue;
    const uint8_t *st = my_tiles + stage * slot;
    if (luma) {
      const int xs = (bits & 1u) ? 0 : kLag, y0 = (bits & 4u) ? 0 : kLag;
      const int x1 = min(W - 32 * ix - kLag, (bits & 2u) ? 32 : 32 - kLag);
      const int y1 = min(H - 32 * by, 32);
      if (x1 > xs && y1 > y0) {
        uint32_t mx[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
        if (xs != 0 || x1 != 32) mx[0] = byte_mask(4 * t, xs, x1), mx[1] = byte_mask(16 + 4 * t, xs, x1);
        luma_rows(reinterpret_cast<const uint32_t *>(st) + y0 * kPL + kResCol0 + t + dxw, sh, y1 - y0, mx, acc);
        nobs += (x1 - xs) * (y1 - y0);
      }
    } else {
      // blocks 2 ix (half 0) and 2 ix + 1 (half 1); bits: flat left, A, B, right | flat above A, B | overflow A, B
      const int bxa = 2 * ix;
      const int y1 = min(ph - 16 * by, 16);
      const bool fla = (bits & 2u) && !(bits & 64u), flb = (bits & 4u) && !(bits & 128u);
      const int xsa = (bits & 1u) ? 0 : kLag, xsb = (bits & 2u) ? 0 : kLag;
      const int y0a = (bits & 16u) ? 0 : kLag, y0b = (bits & 32u) ? 0 : kLag;
      const int x1a = min(pw - 16 * bxa - kLag, (bits & 4u) ? 16 : 16 - kLag);
      const int x1b = min(pw - 16 * (bxa + 1) - kLag, (bits & 8u) ? 16 : 16 - kLag);
      const bool on0 = fla && x1a > xsa && y1 > y0a;
      const bool on1 = flb && x1b > xsb && y1 > y0b;
      const uint32_t m0 = !on0 ? 0u : (xsa == 0 && x1a == 16) ? 0xFFFFFFFFu : byte_mask(4 * t, xsa, x1a);
      const uint32_t m1 = !on1 ? 0u : (xsb == 0 && x1b == 16) ? 0xFFFFFFFFu : byte_mask(4 * t, xsb, x1b);
      const int ya = on0 ? y0a : 99, yb = on1 ? y0b : 99;
      const int ylo = min(ya, yb), yhi = min(max(ya, yb), y1);
      // g = 7 lanes walk the hi tile (and the lo tile one row up) instead of the residual tile
      const uint32_t *base = is7 ? reinterpret_cast<const uint32_t *>(st + kOffHi) + kTapCol0 + t
                                 : reinterpret_cast<const uint32_t *>(st) + kResCol0 + t + dxw;
      const uint32_t *lbase = reinterpret_cast<const uint32_t *>(st + kOffLo) + kTapCol0 + t;  // storage row r = lo row r - 1
      if (ylo < y1) {
        // rows where only one block of the pair is observed (its top margin is 0, the other's is 3)
        if (yhi > ylo) {
          const uint32_t mx[2] = {ya <= ylo ? m0 : 0u, yb <= ylo ? m1 : 0u};
          chroma_rows(base + ylo * kPC, lbase + ylo * kPC, sh, is7, yhi - ylo, mx, acc, ll);
        }
        if (y1 > yhi) {
          const uint32_t mx[2] = {m0, m1};
          chroma_rows(base + yhi * kPC, lbase + yhi * kPC, sh, is7, y1 - yhi, mx, acc, ll);
        }
      }
      nobs += (on0 ? (x1a - xsa) * (y1 - y0a) : 0) + (on1 ? (x1b - xsb) * (y1 - y0b) : 0);
    }
  }
  if (cf >= 0) flush(cf);
}

}  // namespace

bool gram_imma_supported(const Geometry &g) {
  const bool shape = (g.planes == 1) || (g.planes == 3 && g.ss_x == 1 && g.ss_y == 1);
  return shape && g.width >= 8 && g.height >= 8;
}

void gram_imma_tma_boxes(int box[kResidualMaps][2]) {
  box[0][0] = kBoxW, box[0][1] = kLumaRows;
  for (int k = 1; k < kResidualMaps; ++k) box[k][0] = kBoxW, box[k][1] = kChromaRows;
  box[4][1] = kLoRows;
}

void launch_gram_imma(int nframes, const Geometry &g, uint8_t *records, const RecordLayout &rl, const void *tmaps,
                      cudaStream_t st) {
  const int smem = kGramWarps * kWarpSmem;
  static int slots = 0, mode = 0;  // resident CTAs on the device: the persistent grid is one wave
  if (slots == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(gram_imma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gram_imma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gram_imma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gram_imma_kernel<0>, kGramThreads, smem);
    if (const char *e = std::getenv("G1S_GRAM_MODE")) mode = std::atoi(e);
    slots = sms * std::max(per_sm, 1);
  }
  // A warp's share must stay below kMaxShare items (int32 accumulators): more CTAs than one wave if the batch is huge.
  const long long blocks = (long long)nframes * g.nb;
  const long long need = (blocks + (long long)kMaxShare * kLumaWarps - 1) / ((long long)kMaxShare * kLumaWarps);
  const int grid = (int)std::max<long long>(std::min<long long>(slots, std::max<long long>(1, blocks / 8)), need);
  const uint8_t *tm = static_cast<const uint8_t *>(tmaps);
  if (mode == 1)
    gram_imma_kernel<1><<<grid, kGramThreads, smem, st>>>(g, records, rl, nframes, tm);
  else if (mode == 2)
    gram_imma_kernel<2><<<grid, kGramThreads, smem, st>>>(g, records, rl, nframes, tm);
  else
    gram_imma_kernel<0><<<grid, kGramThreads, smem, st>>>(g, records, rl, nframes, tm);
}

}  // namespace g1s
