// gram_imma_kernel — the AR normal-equation (Gram) accumulation on the int8 tensor-core path
// (mma.sync.m16n8k32.s8, SASS IMMA.16832.S8.S8), 4:2:0 and monochrome.
//
// Replaces NoiseModel::add_block_observations (extract_ar_row + the n x n outer-product accumulation) of
// av1-grain's diff module (reached from /root/reference/src/main.rs:442) for every flat block whose
// residuals (and, for chroma, luma tap) fit in int8; the rare block that does not was flagged by
// residual_kernel and is redone exactly by gram_generic_kernel.  All sums are integers, so the result is
// bit-identical to the oracle whatever the summation order.
//
// The kernel never touches the frames: residual_kernel left the s8 residual of Y / Cb / Cr and chroma's
// luma tap (sum of the co-sited 2x2 luma residuals) in engine-owned planes, and every warp pulls the tiles
// of its own work items into its own slice of shared memory with the TMA engine (cp.async.bulk.tensor.2d,
// SASS UTMALDG; frame edges are zero-filled by the hardware, so there is no edge path), one or two items
// ahead of its k-loop.
//
// Why warps are autonomous.  Measured on the B200 (tools/imma_probe.cu, profiles/): the legacy IMMA path
// holds a sub-partition's issue port for its whole 8.4 cycles, so a sub-partition's time is
// 8.4 * IMMAs + (every other warp instruction it issues) -- nothing overlaps, polls and barrier spins are
// paid in full.  So: no CTA barriers, no shared rings, no inter-warp waits.  A warp has a plane for life
// (5 luma + 3 chroma warps per CTA), an equal contiguous share of that plane's blocks over the whole batch,
// its own mbarriers, and it adds its int32 accumulators to the frame's int64 record directly when its share
// leaves a frame (about 330 atomics, once or twice per warp per launch).
//   work item   luma: one 32x32 block; chroma: two adjacent 16x16 blocks (32 columns either way)
//   Tap a = 8q+g with g = cx+3 (the mma lane group), q = cy+3.  A lane's operand of a residual row is ONE
//   32-bit window per half (two LDS + funnel shift), and the same window is the row's B pair and its half of
//   an A quad.  The arithmetic is the row-pair deduplicated form described above the k-loops: 2.5 MMAs per
//   residual row instead of 6 per observed row (round 1), every row fetched once.
//   The observation mask (block margins, frame clipping) is a byte mask on k applied to the A operand.
//   Chroma's luma tap rides in lane group g = 7 (those lanes walk the luma-tap tile), so it costs no MMA.
//   int32 item accumulators are bounded by the share: <= 96 blocks * 1024 * 127^2 < 2^31.
#include "g1s_kernels.h"

#include <algorithm>
#include <cstdlib>

namespace g1s {

namespace {

#ifndef G1S_GRAM_WARPS
#define G1S_GRAM_WARPS 8
#define G1S_LUMA_WARPS 5
#endif
constexpr int kGramWarps = G1S_GRAM_WARPS;
constexpr int kGramThreads = 32 * kGramWarps;
constexpr int kLumaWarps = G1S_LUMA_WARPS;  // per CTA; the others are chroma warps (pair steps per frame: Y 147 k, Cb + Cr 82 k)
constexpr int kMaxShare = 96;     // items between two flushes of a warp: bounds the int32 item accumulators (96 * 1024 * 127^2 < 2^31)
constexpr int kWin = 30;          // blocks per flag window: one ballot holds blocks bx0-1 .. bx0+30
constexpr int kLumaRows = 35;     // 3 halo rows + 32
constexpr int kChromaRows = 19;   // 3 halo rows + 16 (residual and luma-tap tiles alike)
constexpr int kBoxW = 64;         // 16 + 32 + 16 samples: one luma block, or two chroma blocks
// Boxes start 16 samples left of the item (the innermost TMA coordinate must be a multiple of 16 bytes,
// tools/tma_probe.cu).  The k-loops address tiles whose column 0 is the item's origin - 4 samples:
constexpr int kResCol0 = 3;       // word of that column inside a residual box row
constexpr int kTapCol0 = 4;       // word of the item's first sample inside a luma-tap box row
constexpr int kLumaBytes = kLumaRows * kBoxW;      // 2240
constexpr int kChromaBytes = kChromaRows * kBoxW;  // 1216
// A pair step may read one row past the box (the unused upper half of an item's last pair): slots hold one row more.
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

// Window bookkeeping of one item with n = y1 - y0 observed rows, pairs i = 0 .. P-1 (pair i = tile rows y0+2i, +1),
// P = (n + 4) / 2 = m + 2 with m = n / 2: tap row q' sums the lower-half rows of pairs [ceil(q'/2), ceil((n+q')/2)) and
// the upper-half rows of pairs [ceil((q'-1)/2), ceil((n+q'-1)/2)).  Written out (E = lower half, O = upper half):
//   windows open   before pair 0: E0 O0 O1     after pair 0: E1 E2 O2 O3     after pair 1: E3
//   windows close  n even   after pair m-1: E0 O0 O1    m: E1 E2 O2 O3    m+1: E3
//                  n odd    after pair m-1: O0          m: E0 E1 O1 O2    m+1: E2 E3 O3
// so an item is: two pair steps with fixed snapshots, a hook-free core of m - 3 steps, three steps with the closing
// snapshots.  That needs m >= 3; the rare shorter items (a sliver of rows at the bottom of the frame) are handed to
// the exact generic kernel by the caller.
constexpr int kMinRows = 6;
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

// All rows of one item (or of one block of a chroma pair): base points at tile row 0; n = y1 - y0 >= kMinRows.
__device__ __forceinline__ void item_rows(const uint32_t *__restrict__ base, int sh, int y0, int y1, const uint32_t (&mx)[2],
                                          Ring &w, int (&acc)[5][4], uint32_t (&G)[10][2]) {
  const int n = y1 - y0, m = n >> 1;
  const bool odd = n & 1;
  const uint32_t *p = base + y0 * kRowWords;
  int ph = 0;
  snapshot(G, acc, kEvA, false);
  pair_step<0>(w, p, sh, mx, acc);
  snapshot(G, acc, kEvB, false);
  pair_step<1>(w, p + 2 * kRowWords, sh, mx, acc);
  snapshot(G, acc, kEvC, false);
  p += 4 * kRowWords;
  int core = m - 3;  // pairs 2 .. m-2
#pragma unroll 1
  for (; core >= 3; core -= 3, p += 6 * kRowWords) {
    pair_step<2>(w, p, sh, mx, acc);
    pair_step<0>(w, p + 2 * kRowWords, sh, mx, acc);
    pair_step<1>(w, p + 4 * kRowWords, sh, mx, acc);
  }
  ph = 2;
#pragma unroll 1
  for (; core > 0; --core, p += 2 * kRowWords) pair_step_rt(ph, w, p, sh, mx, acc);
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
  const int slot = luma ? kLumaSlot : kChromaSlot;
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&sm.full[warp][s], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  int acc[5][4];        // running sums of the five row products (never reset, see the k-loop notes)
  uint32_t G[10][2];    // this share's Gram blocks (q', dy) since the last flush
  Ring ring;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[i][r] = 0;
#pragma unroll
  for (int i = 0; i < 10; ++i) G[i][0] = G[i][1] = 0u;
#pragma unroll
  for (int i = 0; i < 6; ++i) ring.B[i][0] = ring.B[i][1] = 0u;
  int nobs = 0;  // observations of the items accumulated since the last flush

  int sh = 8 * ((gq + 1) & 3);
  asm volatile("" : "+r"(sh));  // same: one register instead of four instructions per funnel shift group
  const int dxw = (gq + 1) >> 2;
  const bool is7 = gq == 7;

  // Gram blocks -> the frame's int64 record, straight from registers: element r of block (q', dy) in lane (gq, t)
  // is the product of the later tap (q', g' = gq) with the earlier tap (q' - dy, g = 2t + r).
  auto flush = [&](int f) {
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
          emit(gram, 8 * (q - dy) + 2 * t + r, 8 * q + gq, v, !luma);
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
  int hstage = 0, tstage = 0;   // their stages ( = position % kStages, kept incrementally)
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
    if (++hstage == kStages) hstage = 0;
    if (lane == 0) {
      sm.fifo[warp][head & 3] = make_int4(pf, pby, pwx + k, (int)bits);
      const uint8_t *fmaps = tmaps + (size_t)pf * kResidualMaps * 128;
      uint8_t *st = my_tiles + stage * slot;
      uint64_t *bar = &sm.full[warp][stage];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this warp's reads of the stage precede the refill
      if (luma) {
        mbar_expect_tx(bar, kLumaBytes);
        tma_load_2d(st, fmaps, 32 * (pwx + k) - 16, 32 * pby - 3, bar);
      } else {
        const int cx = 32 * (pwx + k) - 16, cy = 16 * pby - 3;
        mbar_expect_tx(bar, 2 * kChromaBytes);
        tma_load_2d(st, fmaps + plane * 128, cx, cy, bar);
        tma_load_2d(st + kOffTap, fmaps + 3 * 128, cx, cy, bar);
      }
    }
    ++head;
    return true;
  };

  int cf = -1;     // frame the accumulators belong to
  int since = 0;   // items accumulated since the last flush: the int32 item accumulators hold kMaxShare of them
  for (;;) {
    __syncwarp();  // every lane is done with the stage about to be refilled
    while (head - tail < kStages && produce()) {
    }
    if (head == tail) break;
    __syncwarp();
    const int4 item = sm.fifo[warp][tail & 3];
    const int stage = tstage;
    if (++tstage == kStages) tstage = 0;
    ++tail;
    const int f = item.x, by = item.y, ix = item.z;
    const uint32_t bits = (uint32_t)item.w;
    if (f != cf || since == kMaxShare) {
      if (cf >= 0) flush(cf);
      cf = f;
      since = 0;
    }
    ++since;
    mbar_wait(&sm.full[warp][stage], (phases >> stage) & 1u);
    phases ^= 1u << stage;
    const uint8_t *st = my_tiles + stage * slot;
    // up to two passes over the item's rows: (first observed row, masks of the two halves)
    int npass = 0, y1, ya = 0, yb = 0;
    bool short_a = false, short_b = false;  // blocks of the item left to the generic kernel (too few rows)
    uint32_t ma[2] = {0u, 0u}, mb[2] = {0u, 0u};
    const uint32_t *base;
    if (luma) {
      const int xs = (bits & 1u) ? 0 : kLag, y0 = (bits & 4u) ? 0 : kLag;
      const int x1 = min(W - 32 * ix - kLag, (bits & 2u) ? 32 : 32 - kLag);
      y1 = min(H - 32 * by, 32);
      if (x1 > xs && y1 > y0 && y1 - y0 < kMinRows) {
        short_a = true;
      } else if (x1 > xs && y1 > y0) {
        const bool full = xs == 0 && x1 == 32;
        ma[0] = full ? 0xFFFFFFFFu : byte_mask(4 * t, xs, x1);
        ma[1] = full ? 0xFFFFFFFFu : byte_mask(16 + 4 * t, xs, x1);
        ya = y0;
        npass = 1;
        nobs += (x1 - xs) * (y1 - y0);
      }
      base = reinterpret_cast<const uint32_t *>(st) + kResCol0 + t + dxw;
    } else {
      // blocks 2 ix (half 0) and 2 ix + 1 (half 1); bits: flat left, A, B, right | flat above A, B | overflow A, B
      const int bxa = 2 * ix;
      y1 = min(ph - 16 * by, 16);
      const bool fla = (bits & 2u) && !(bits & 64u), flb = (bits & 4u) && !(bits & 128u);
      const int xsa = (bits & 1u) ? 0 : kLag, xsb = (bits & 2u) ? 0 : kLag;
      const int y0a = (bits & 16u) ? 0 : kLag, y0b = (bits & 32u) ? 0 : kLag;
      const int x1a = min(pw - 16 * bxa - kLag, (bits & 4u) ? 16 : 16 - kLag);
      const int x1b = min(pw - 16 * (bxa + 1) - kLag, (bits & 8u) ? 16 : 16 - kLag);
      bool on0 = fla && x1a > xsa && y1 > y0a;
      bool on1 = flb && x1b > xsb && y1 > y0b;
      if (on0 && y1 - y0a < kMinRows) short_a = true, on0 = false;
      if (on1 && y1 - y0b < kMinRows) short_b = true, on1 = false;
      const uint32_t m0 = !on0 ? 0u : (xsa == 0 && x1a == 16) ? 0xFFFFFFFFu : byte_mask(4 * t, xsa, x1a);
      const uint32_t m1 = !on1 ? 0u : (xsb == 0 && x1b == 16) ? 0xFFFFFFFFu : byte_mask(4 * t, xsb, x1b);
      if (on0 && on1 && y0a == y0b) {
        ma[0] = m0, ma[1] = m1, ya = y0a, npass = 1;
      } else {
        // the two blocks start on different rows (one has a flat block above it, the other not) or only one is
        // observed: their row windows differ, so they are accumulated one after the other
        if (on0) ma[0] = m0, ya = y0a, npass = 1;
        if (on1) {
          if (npass == 0) ma[1] = m1, ya = y0b;
          else mb[1] = m1, yb = y0b;
          ++npass;
        }
      }
      // g = 7 lanes walk the luma-tap tile (same rows, no shift) instead of the residual tile
      base = is7 ? reinterpret_cast<const uint32_t *>(st + kOffTap) + kTapCol0 + t
                 : reinterpret_cast<const uint32_t *>(st) + kResCol0 + t + dxw;
      nobs += (on0 ? (x1a - xsa) * (y1 - y0a) : 0) + (on1 ? (x1b - xsb) * (y1 - y0b) : 0);
    }
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
      if (pass == 1) ma[0] = mb[0], ma[1] = mb[1], ya = yb;
      item_rows(base, sh, ya, y1, ma, ring, acc, G);
    }
    if (short_a | short_b) {
      // a sliver of fewer than kMinRows observed rows (bottom of the frame): the exact generic kernel, launched right
      // after this one, accumulates the block; only this warp ever reads or writes the flags of its own items
      uint8_t *rec = records + (size_t)f * rl.bytes;
      if (lane == 0) {
        const int b0 = by * g.nbw + (luma ? ix : 2 * ix);
        if (short_a) (rec + rl.off_ovf)[(size_t)plane * g.nb + b0] = 1;
        if (short_b) (rec + rl.off_ovf)[(size_t)plane * g.nb + b0 + 1] = 1;
        atomicAdd(reinterpret_cast<unsigned long long *>(rec + rl.off_ovf_count), 1ull);
      }
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
}

void launch_gram_imma(int nframes, const Geometry &g, uint8_t *records, const RecordLayout &rl, const void *tmaps,
                      cudaStream_t st) {
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
    slots = sms * std::max(per_sm, 1);
  }
  // one wave of persistent CTAs whatever the batch size: a warp flushes its int32 item accumulators every kMaxShare items
  const long long blocks = (long long)nframes * g.nb;
  const int grid = (int)std::min<long long>(slots, std::max<long long>(1, blocks / 8));
  gram_imma_kernel<<<grid, kGramThreads, smem, st>>>(g, records, rl, nframes, static_cast<const uint8_t *>(tmaps));
}

}  // namespace g1s
