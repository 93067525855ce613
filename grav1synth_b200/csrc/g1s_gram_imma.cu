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
// halves of chroma's luma tap in engine-owned planes, and this kernel pulls tiles of them into shared
// memory with the TMA engine (cp.async.bulk.tensor.2d, SASS UTMALDG; frame edges are zero-filled by the
// hardware, so there is no edge path), through a kStages-deep ring of full / empty mbarriers.  Nothing but
// k-loops runs on the SM: no conversion, no statistics, no CTA-wide barrier inside a run.
//
// Work unit: eight horizontally adjacent 32x32 luma blocks plus their co-sited 16x16 Cb and Cr blocks.
// Persistent CTAs of 12 warps (three per SM sub-partition, all with the same 32 k-steps per unit), two per
// SM, each walking an equal contiguous share of the batch's units:
//   warps 0-7   luma: block = warp
//   warps 8-11  chroma: plane = (warp - 8) / 2, blocks 4 * ((warp - 8) % 2) .. + 3 as two block pairs
//   One k-step = 32 pixels of one row.  X[k][a] = residual at pixel k shifted by tap a;
//   D += X^T X over the upper block-triangle (6 m16n8k32 MMAs).  Tap a = 8q+g with g = cx+3 (lane
//   group), q = cy+3: a thread's four taps are the SAME column offset on four consecutive rows, so its
//   operands slide down one row per k-step: one new 32-bit window (two LDS + funnel shift) per half is
//   fetched, the B operand of a row is a register pair and the A operand of two consecutive rows a
//   register quad that serves as the lower m-tile now and as the upper m-tile two steps later.
//   The observation mask (block margins, frame clipping) is a byte mask on k applied to the operand
//   words (mask^2 = mask, so masking both A and B is exact).  Chroma's luma tap rides in lane group 7.
//   flush: int32 accumulators (bounded: <= 15 units * 32 k-steps * 32 * 2^14 * 8 warps < 2^31)
//   -> shared-memory reduction -> int64 global atomics, one per tap pair per CTA per plane.
#include "g1s_kernels.h"

#include <algorithm>
#include <cstdlib>

namespace g1s {

namespace {

constexpr int kUnitBlocks = 8;
constexpr int kGramWarps = 12;
constexpr int kGramThreads = 32 * kGramWarps;
constexpr int kFlushUnits = 15;   // units between accumulator flushes (bounds the int32 accumulators, see above)
constexpr int kStages = 3;        // TMA ring depth (2 x (3 x 23.1 KB + 18.1 KB static) fits one SM with room to spare)
constexpr int kRefillLag = 1;     // a stage is refilled this many iterations after its unit was consumed
constexpr int kPL = 40;           // luma tile pitch in 32-bit words (160-byte box rows)
constexpr int kLumaRows = 35;     // 3 halo rows + 32
constexpr int kPC = 40;           // chroma / tap tile pitch in words (160-byte box rows)
constexpr int kChromaRows = 19;   // 3 halo rows + 16
constexpr int kLoRows = 20;       // the lo tile starts one row higher (row -1 of the first step is read, never used)
constexpr int kBoxW = 160;        // 16 + 128 + 16 samples: half a unit of luma, a whole unit of chroma
// Boxes start 16 samples left of their first block (the innermost TMA coordinate must be a multiple of 16
// bytes, tools/tma_probe.cu).  The k-loops address tiles whose column 0 is that block's origin - 4 samples:
constexpr int kResCol0 = 3;       // word of that column inside a residual box row
constexpr int kTapCol0 = 4;       // word of the unit's first sample inside a hi / lo box row
constexpr int kLumaBytes = kLumaRows * kBoxW;      // 5600 (two of these per unit)
constexpr int kChromaBytes = kChromaRows * kBoxW;  // 3040
constexpr int kLoBytes = kLoRows * kBoxW;          // 3200
constexpr int kLumaSlot = 5632, kChromaSlot = 3072;
constexpr int kOffLuma = 0, kOffCb = 2 * kLumaSlot, kOffCr = kOffCb + kChromaSlot, kOffHi = kOffCr + kChromaSlot,
              kOffLo = kOffHi + kChromaSlot;
constexpr int kStageBytes = kOffLo + kLoBytes;     // 23680, every tile 128-byte aligned
static_assert(kLumaBytes <= kLumaSlot && kChromaBytes <= kChromaSlot && kStageBytes % 128 == 0, "stage layout");

struct __align__(16) GramSmem {
  int dl[2][3][6 * 4 * 32];        // [flush parity][plane]: accumulators reduced over the warps of a plane
  int ll[2][2];                    // [flush parity][chroma plane]: sum lo*lo of the luma tap over observed pixels
  int arrived[2];                  // [flush parity]: warps that have added their accumulators
  uint64_t full[kStages], empty[kStages];
  int prod[4];                     // producer cursor (thread 0 only): frame, block row, unit, units requested
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
// 1 = TMA ring and waits only, no k-loops; 2 = k-loops on whatever is in shared memory, no TMA traffic.
template <int MODE>
__global__ void __launch_bounds__(kGramThreads, 2)
gram_imma_kernel(Geometry g, uint8_t *__restrict__ records, RecordLayout rl, int nframes,
                 const uint8_t *__restrict__ tmaps) {
  extern __shared__ __align__(128) uint8_t stages[];
  __shared__ GramSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;  // mma "groupID" (= cx + 3) and thread-in-group
  const bool has_chroma = g.planes == 3;
  const int W = g.width, H = g.height, pw = W >> 1, ph = H >> 1;
  const int nsu = (g.nbw + kUnitBlocks - 1) / kUnitBlocks;
  // Persistent CTAs: the units of the whole batch in (frame, block row, unit) order, an equal contiguous
  // share per CTA, so the TMA ring never drains and all CTAs finish together.
  const int per_frame = g.nbh * nsu;
  const long long total = (long long)nframes * per_frame;
  const int L0 = (int)(total * blockIdx.x / gridDim.x), L1 = (int)(total * (blockIdx.x + 1) / gridDim.x);
  if (L0 >= L1) return;

  int acc[6][4];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[i][r] = 0;
  int ll = 0;          // chroma warps, g = 7 lanes
  int nobs = 0;        // observation count of this warp's blocks (<= kFlushUnits * 2 * 1024 between flushes)

  for (int i = tid; i < 2 * 3 * 6 * 4 * 32 + 4 + 2; i += kGramThreads) (&sm.dl[0][0][0])[i] = 0;  // dl, ll, arrived are contiguous
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], kGramWarps);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int sh = 8 * ((gq + 1) & 3);
  const int dxw = (gq + 1) >> 2;
  const bool is7 = gq == 7;
  const bool luma_warp = warp < 8;
  const int plane_of_warp = luma_warp ? 0 : 1 + ((warp - 8) >> 1);
  const int jblk = luma_warp ? warp : 4 * ((warp - 8) & 1);  // first block of the unit this warp works on

  // one elected thread asks the TMA engine for the tiles of a unit
  auto issue_tma = [&](int f, int by, int u, int s) {
    const uint8_t *fmaps = tmaps + (size_t)f * kResidualMaps * 128;
    uint8_t *st = stages + s * kStageBytes;
    uint64_t *bar = &sm.full[s];
    mbar_expect_tx(bar, 2 * kLumaBytes + (has_chroma ? 3 * kChromaBytes + kLoBytes : 0));
    tma_load_2d(st + kOffLuma, fmaps + 0 * 128, 256 * u - 16, 32 * by - 3, bar);
    tma_load_2d(st + kOffLuma + kLumaSlot, fmaps + 0 * 128, 256 * u + 112, 32 * by - 3, bar);
    if (has_chroma) {
      tma_load_2d(st + kOffCb, fmaps + 1 * 128, 128 * u - 16, 16 * by - 3, bar);
      tma_load_2d(st + kOffCr, fmaps + 2 * 128, 128 * u - 16, 16 * by - 3, bar);
      tma_load_2d(st + kOffHi, fmaps + 3 * 128, 128 * u - 16, 16 * by, bar);
      tma_load_2d(st + kOffLo, fmaps + 4 * 128, 128 * u - 16, 16 * by - 1, bar);
    }
  };
  auto advance = [&](int &f, int &by, int &u) {
    if (++u == nsu) {
      u = 0;
      if (++by == g.nbh) by = 0, ++f;
    }
  };
  // Flat / overflow flags a warp needs for a unit, one byte per lane, combined with a ballot:
  //   luma   lanes 0-2: flat(by, bx-1 .. bx+1); 3: flat(by-1, bx); 4: ovf[0](by, bx)
  //   chroma lanes 0-5: flat(by, bx0-1 .. bx0+4); 6-9: flat(by-1, bx0 .. bx0+3); 10-13: ovf[c](by, bx0 .. bx0+3)
  auto load_flag = [&](int f, int by, int u) -> uint32_t {
    const uint8_t *rec = records + (size_t)f * rl.bytes;
    const int bx0 = kUnitBlocks * u + jblk;
    int yy = by, xx = bx0;
    const uint8_t *base = rec + rl.off_flat;
    if (luma_warp) {
      if (lane < 3) xx = bx0 - 1 + lane;
      else if (lane == 3) yy = by - 1;
      else if (lane == 4) base = rec + rl.off_ovf;
      else return 0u;
    } else {
      if (lane < 6) xx = bx0 - 1 + lane;
      else if (lane < 10) yy = by - 1, xx = bx0 + lane - 6;
      else if (lane < 14) base = rec + rl.off_ovf + (size_t)plane_of_warp * g.nb, xx = bx0 + lane - 10;
      else return 0u;
    }
    if (xx < 0 || xx >= g.nbw || yy < 0) return 0u;
    return base[yy * g.nbw + xx];
  };

  // Accumulators out: int32 registers -> shared-memory sums per plane -> int64 global atomics into the
  // frame's record.  Needed at frame changes, at the end, and every kFlushUnits units (int32 bound).
  // No CTA barrier: every warp adds its registers to the buffer of the flush's parity and counts itself in;
  // the warp that arrives last emits the buffer and re-zeroes it.  The buffer of one parity is reused two
  // flushes later, by which time its emission is long over (warps are never more than kStages units apart).
  int flush_gen = 0;
  auto flush = [&](int f) {
    uint8_t *rec = records + (size_t)f * rl.bytes;
    const int pb = flush_gen & 1;
    ++flush_gen;
    if (luma_warp || has_chroma) {
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (acc[i][r]) atomicAdd(&sm.dl[pb][plane_of_warp][(i * 4 + r) * 32 + lane], acc[i][r]);
          acc[i][r] = 0;
        }
      if (!luma_warp && is7 && ll) atomicAdd(&sm.ll[pb][plane_of_warp - 1], ll);
      ll = 0;
      if (lane == 0 && nobs)
        atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_nobs) + plane_of_warp, (unsigned long long)(long long)nobs);
      nobs = 0;
    }
    __syncwarp();
    int order = 0;
    if (lane == 0) {
      __threadfence_block();
      order = atomicAdd(&sm.arrived[pb], 1);
    }
    order = __shfl_sync(0xffffffffu, order, 0);
    if (order != kGramWarps - 1) return;
    __threadfence_block();
#pragma unroll 1
    for (int plane = 0; plane < g.planes; ++plane) {
      unsigned long long *gram = reinterpret_cast<unsigned long long *>(rec + rl.off_gram) + (size_t)plane * kPairs;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int mrow = i >= 4 ? 16 : 0;
        const int ncol = i >= 4 ? 8 * (i - 2) : 8 * i;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          int *slot = &sm.dl[pb][plane][(i * 4 + r) * 32 + lane];
          const int v = *slot;
          if (v == 0) continue;
          *slot = 0;
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
      if (plane > 0 && lane == 0 && sm.ll[pb][plane - 1]) {
        atomicAdd(&gram[pair_index(24, 24)], (unsigned long long)(long long)sm.ll[pb][plane - 1]);
        sm.ll[pb][plane - 1] = 0;
      }
    }
    __syncwarp();
    if (lane == 0) sm.arrived[pb] = 0;
  };

  // unit cursors: consumers (all threads) and the producer (thread 0)
  int f = L0 / per_frame, by, u;
  {
    const int r = L0 - f * per_frame;
    by = r / nsu;
    u = r - by * nsu;
  }
  const int count = L1 - L0;
  if (tid == 0 && MODE != 2) {
    int pf = f, pby = by, pu = u, issued = 0;
    for (; issued < kStages - kRefillLag && issued < count; ++issued) {
      issue_tma(pf, pby, pu, issued);
      advance(pf, pby, pu);
    }
    sm.prod[0] = pf, sm.prod[1] = pby, sm.prod[2] = pu, sm.prod[3] = issued;
  }
  uint32_t flagv = load_flag(f, by, u);

  for (int it = 0; it < count; ++it) {
    const uint32_t bits = __ballot_sync(0xffffffffu, flagv != 0);
    const bool last = it + 1 == count;
    {
      int nf = f, nby = by, nu = u;
      advance(nf, nby, nu);
      flagv = last ? 0u : load_flag(nf, nby, nu);  // next unit's flags travel behind this unit's k-loops
    }

    const int s = it % kStages;
    if (MODE != 2) mbar_wait(&sm.full[s], (uint32_t)(it / kStages) & 1u);
    const uint8_t *st = stages + s * kStageBytes;
    const int Y0 = 32 * by, CY0 = 16 * by;

    if (MODE == 1) {
    } else if (luma_warp) {
      const int j = jblk;
      if ((bits & 2u) && !(bits & 16u)) {
        const int xs = (bits & 1u) ? 0 : kLag, y0 = (bits & 8u) ? 0 : kLag;
        const int x1 = min(W - 32 * (kUnitBlocks * u + j) - kLag, (bits & 4u) ? 32 : 32 - kLag);
        const int y1 = min(H - Y0, 32);
        if (x1 > xs && y1 > y0) {
          const uint32_t mx[2] = {byte_mask(4 * t, xs, x1), byte_mask(16 + 4 * t, xs, x1)};
          const uint32_t *tile = reinterpret_cast<const uint32_t *>(st + kOffLuma + (j >> 2) * kLumaSlot);
          luma_rows(tile + y0 * kPL + kResCol0 + 8 * (j & 3) + t + dxw, sh, y1 - y0, mx, acc);
          nobs += (x1 - xs) * (y1 - y0);
        }
      }
    } else if (has_chroma) {
      const int c = plane_of_warp - 1;
      const int y1 = min(ph - CY0, 16);
#pragma unroll 1
      for (int pr = 0; pr < 2; ++pr) {  // the two block pairs of this warp: blocks jblk + 2 pr (half 0) and + 1 (half 1)
        const uint32_t fb = bits >> (2 * pr), ub = bits >> (6 + 2 * pr), ob = bits >> (10 + 2 * pr);
        const int bxa = kUnitBlocks * u + jblk + 2 * pr;
        const bool fla = (fb & 2u) && !(ob & 1u), flb = (fb & 4u) && !(ob & 2u);
        const int xsa = (fb & 1u) ? 0 : kLag, xsb = (fb & 2u) ? 0 : kLag;
        const int y0a = (ub & 1u) ? 0 : kLag, y0b = (ub & 2u) ? 0 : kLag;
        const int x1a = min(pw - 16 * bxa - kLag, (fb & 4u) ? 16 : 16 - kLag);
        const int x1b = min(pw - 16 * (bxa + 1) - kLag, (fb & 8u) ? 16 : 16 - kLag);
        const bool on0 = fla && x1a > xsa && y1 > y0a;
        const bool on1 = flb && x1b > xsb && y1 > y0b;
        const uint32_t m0 = on0 ? byte_mask(4 * t, xsa, x1a) : 0u;
        const uint32_t m1 = on1 ? byte_mask(4 * t, xsb, x1b) : 0u;
        const int ya = on0 ? y0a : 99, yb = on1 ? y0b : 99;
        const int ylo = min(ya, yb), yhi = min(max(ya, yb), y1);
        const int wofs = 4 * jblk + 8 * pr + t;  // words from the unit's first sample to this lane's half 0
        // g = 7 lanes walk the hi tile (and the lo tile one row up) instead of the residual tile
        const uint32_t *base = is7 ? reinterpret_cast<const uint32_t *>(st + kOffHi) + kTapCol0 + wofs
                                   : reinterpret_cast<const uint32_t *>(st + (c ? kOffCr : kOffCb)) + kResCol0 + wofs + dxw;
        const uint32_t *lbase = reinterpret_cast<const uint32_t *>(st + kOffLo) + kTapCol0 + wofs;  // storage row r = lo row r - 1
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

    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);  // this warp is done with the stage
    if (MODE != 2 && tid == 0 && sm.prod[3] < count) {
      // refill the stage consumed kRefillLag iterations ago (every warp has almost surely left it)
      const int prev = sm.prod[3] - kStages;  // iteration whose stage the next requested unit reuses ( = it - kRefillLag )
      if (prev >= 0) mbar_wait(&sm.empty[prev % kStages], (uint32_t)(prev / kStages) & 1u);
      int pf = sm.prod[0], pby = sm.prod[1], pu = sm.prod[2];
      issue_tma(pf, pby, pu, sm.prod[3] % kStages);
      advance(pf, pby, pu);
      sm.prod[0] = pf, sm.prod[1] = pby, sm.prod[2] = pu, sm.prod[3] += 1;
    }

    // accumulators out at frame changes, at the end, and every kFlushUnits units (int32 bound)
    const bool frame_end = u == nsu - 1 && by == g.nbh - 1;
    if (last || frame_end || it % kFlushUnits == kFlushUnits - 1) flush(f);
    advance(f, by, u);
  }
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
  const int nsu = (g.nbw + kUnitBlocks - 1) / kUnitBlocks;
  const long long total = (long long)nframes * g.nbh * nsu;
  const int smem = kStages * kStageBytes;
  static int slots = 0, mode = 0;  // resident CTAs on the device: the persistent grid is exactly one wave
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
  const int grid = (int)std::min<long long>(total, slots);
  const uint8_t *tm = static_cast<const uint8_t *>(tmaps);
  if (mode == 1)
    gram_imma_kernel<1><<<grid, kGramThreads, smem, st>>>(g, records, rl, nframes, tm);
  else if (mode == 2)
    gram_imma_kernel<2><<<grid, kGramThreads, smem, st>>>(g, records, rl, nframes, tm);
  else
    gram_imma_kernel<0><<<grid, kGramThreads, smem, st>>>(g, records, rl, nframes, tm);
}

}  // namespace g1s
