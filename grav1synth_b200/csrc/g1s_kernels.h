// Device-side interface of the grain-estimation engine: data layout in HBM and the
// kernel launchers.  Host code (g1s_engine.cpp) includes this; no CUDA types leak
// beyond cudaStream_t.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace g1s {

constexpr int kBlock = 32;         // BLOCK_SIZE of av1-grain's diff module (32x32 luma blocks)
constexpr int kLag = 3;            // NOISE_MODEL_LAG
constexpr int kTaps = 26;          // 24 AR taps + [24] luma tap (chroma only) + [25] centre sample
constexpr int kPairs = kTaps * (kTaps + 1) / 2;  // 351 upper-triangular products

// One frame pair as the kernels see it.  Pointers are device pointers; strides in bytes.
struct FrameDesc {
  const void *src[3];
  const void *den[3];
  uint32_t src_stride[3];
  uint32_t den_stride[3];
};

// Stream geometry, constant for a handle.
struct Geometry {
  int width, height;      // luma samples
  int ss_x, ss_y;         // chroma subsampling log2
  int planes;             // 1 (monochrome) or 3
  int src_shift, den_shift;  // bit_depth - 8: samples are reduced with a truncating >> (frame_into_u8)
  int src_bytes, den_bytes;  // bytes per sample (1 or 2)
  int nbw, nbh, nb;       // 32x32 luma block grid
};

// Constants of FlatBlockFinder::new computed once on the host in f64.
struct FlatConsts {
  double ata_inv[9];
};

// Per-frame result record, written by the kernels and copied back to the host.
// Layout of one record (all offsets are 8-byte aligned), for nb blocks:
//   int64  gram[3][kPairs]     upper-triangular integer Gram sums per plane
//   int64  nobs[3]             observations per plane
//   int64  num_flat            blocks with a non-zero flat flag
//   uint32 luma_sum[nb]        sum of 8-bit source luma over each (frame-clipped) block
//   int32  rsum[3][nb]         sum of the residual over each block, per plane
//   uint32 rsq[3][nb]          sum of the squared residual over each block, per plane
//   float  score[nb]           sigmoid flatness score (0 when var <= threshold)
//   uint8  flat[nb]            0 / 1 / 255 flat flags (padded to 8 bytes)
//   int64  ovf_count           blocks whose residual tile left the int8 range (device bookkeeping)
//   uint8  ovf[3][nb]          per plane: block must be accumulated by the generic (int32) kernel
//   double gramf[3][kPairs]    strict mode only (gram_order = reference order): the same tap pairs accumulated
//                              term by term as RN(acc + RN(product / 255^2)) in the reference's pixel order
struct RecordLayout {
  size_t off_gram, off_nobs, off_num_flat, off_luma_sum, off_rsum, off_rsq, off_score, off_flat, off_ovf_count,
      off_ovf, off_gramf, bytes;
  static RecordLayout make(int nb);
};

// aligned: every source-luma row start is 16-byte aligned (vector loads allowed).
// y8 (optional): the batch's 8-bit source luma planes as residual_kernel wrote them (ResidualStore::off_y8, frame f at
// y8 + f * y8_frame_bytes, rows y8_pitch apart); when given, the frames themselves are not read.
void launch_flat_features(const FrameDesc *frames, int nframes, const Geometry &g, const FlatConsts &fc,
                          uint8_t *records, const RecordLayout &rl, bool aligned, cudaStream_t st,
                          const uint8_t *y8 = nullptr, size_t y8_frame_bytes = 0, uint32_t y8_pitch = 0);
void launch_flat_select(int nframes, const Geometry &g, uint8_t *records, const RecordLayout &rl, cudaStream_t st);
// only_overflow = false: every flat block, statistics included (any subsampling, any alignment).
// only_overflow = true: just the blocks the tensor-core kernel flagged (Gram sums and statistics).
void launch_gram_generic(const FrameDesc *frames, int nframes, const Geometry &g, uint8_t *records,
                         const RecordLayout &rl, bool only_overflow, cudaStream_t st);
// Strict mode (g1s_diff_config.gram_order = G1S_GRAM_REF_ORDER): fills RecordLayout::off_gramf with the f64 sums in
// the reference's accumulation order; needs the final flat flags.  Any subsampling, any residual magnitude.
void launch_gram_strict(const FrameDesc *frames, int nframes, const Geometry &g, uint8_t *records,
                        const RecordLayout &rl, cudaStream_t st);
// The per-frame half of the host model on the device (g1s_latest.cu): one LatestFrame digest (digest_doubles f64,
// LatestFrame::to_digest layout) per frame of the batch, bit-identical to NoiseModel::compute_latest on the same record.
// latest_supported: the frame's per-block arrays fit in shared memory (up to ~9 000 blocks: 4K yes, 8K no) and luma and
// chroma agree on which blocks measure (no sliver at the frame edge).
bool latest_supported(const Geometry &g);
void launch_latest(int nframes, const Geometry &g, const uint8_t *records, const RecordLayout &rl, bool strict,
                   double *digests, int digest_doubles, cudaStream_t st);
// Engine-owned s8 planes of a batch, written by residual_kernel and read (through TMA boxes) by
// gram_imma_kernel: per frame the residual of Y, Cb, Cr and chroma's luma tap, the sum of the co-sited 2x2 luma
// residuals (blocks where it leaves int8 are flagged for the exact kernel, like residuals that do).  Pitches are
// multiples of 16 bytes, plane offsets of 256.
struct ResidualStore {
  int8_t *base;            // frame f of the batch starts at base + f * frame_bytes
  size_t frame_bytes;
  size_t off_res[3];
  size_t off_tap;
  size_t off_y8;           // the source luma reduced to 8 bits (what FlatBlockFinder reads), same pitch as the luma residual
  uint32_t pitch_l, pitch_c;
  static ResidualStore make(const Geometry &g);  // offsets / pitches only; base stays null
};
constexpr int kResidualMaps = 4;  // TMA descriptors per frame: residual Y, Cb, Cr, chroma's luma tap

// int8 tensor-core (mma.sync m16n8k32) path for 4:2:0 / monochrome streams: residual_kernel then
// gram_imma_kernel.
bool gram_imma_supported(const Geometry &g);
// aligned: every plane base and stride of every frame is 16-byte aligned (vector loads allowed).
void launch_residual(const FrameDesc *frames, int nframes, const Geometry &g, const ResidualStore &rs,
                     uint8_t *records, const RecordLayout &rl, bool aligned, cudaStream_t st);
// gram_plan_kernel: flat / overflow flags -> the plane's work list (strips of vertically adjacent blocks with one row
// window, one 16-byte descriptor per strip and tile) + observation counts; may flag more blocks for the generic kernel.
// plan: gram_plan_bytes(nframes, g) bytes; counts: int[nframes][3] units per frame and plane + 1 (the Gram kernel's
// chunk counter).  nframes <= 64.
size_t gram_plan_bytes(int nframes, const Geometry &g);
void launch_gram_plan(int nframes, const Geometry &g, uint8_t *records, const RecordLayout &rl, void *plan, int *counts,
                      cudaStream_t st);
// tmaps: device array of CUtensorMap[nframes][kResidualMaps] over the ResidualStore planes (extents =
// the loop extents W x H and (W>>1) x (H>>1), boxes from gram_imma_tma_boxes, zero fill).
void launch_gram_imma(int nframes, const Geometry &g, uint8_t *records, const RecordLayout &rl, const void *tmaps,
                      const void *plan, int *counts, cudaStream_t st);
// box[k] = {width, height} in bytes / rows for descriptor k of a frame
void gram_imma_tma_boxes(int box[kResidualMaps][2]);

}  // namespace g1s
