/*
 * g1s.h — C ABI of the B200 film-grain estimation engine (grav1synth `diff` hot path).
 *
 * This is the drop-in boundary: every entry point below replaces one call the
 * reference makes into `av1_grain::DiffGenerator` (crate av1-grain 0.4.2, pinned in
 * /root/reference/Cargo.toml:15, Cargo.lock:92-104) from its `Commands::Diff` arm.
 * Plain pointers and sizes only; no torch / CUDA types cross this boundary.
 *
 *   reference call site (src/main.rs)                      entry point here
 *   -----------------------------------------------------  -------------------------
 *   DiffGenerator::new(fps, src_bd, den_bd)      :420-427  g1s_diff_create
 *   differ.diff_frame(&src, &den)?  :442 / 462 / 482 / 502  g1s_diff_push_frame
 *   differ.finish() -> Vec<GrainTableSegment>        :524  g1s_diff_finish
 *   (drop of `differ`)                                      g1s_diff_destroy
 *   anyhow::Error text (propagated by `?`)           :442  g1s_diff_last_error
 *   writeln!("filmgrn1") + write_film_grain_segment
 *                                        :525-530, 631-696  g1s_write_grain_table
 *   GrainTableSegment / FilmGrainParams
 *                 src/main.rs:698-713, parser/grain.rs:21-81  g1s_segment
 *
 * Threading: one handle is one stream of frame pairs; calls on a handle must be
 * serialised by the caller (the reference loop src/main.rs:432-521 is single
 * threaded too).  The engine itself is asynchronous: `push_frame` copies the
 * borrowed planes into a pinned staging ring and returns; an error raised while
 * processing frame k may therefore surface on a later push or on finish.
 *
 * There is NO CPU fallback: if no CUDA device is usable every call fails with
 * G1S_E_CUDA.
 */
#ifndef G1S_H_
#define G1S_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G1S_ABI_VERSION 2

/* Capacities fixed by the AV1 film-grain syntax; av1_grain::NUM_Y_POINTS etc. as
 * imported at /root/reference/src/parser/grain.rs:2 and used at :26-49. */
#define G1S_NUM_Y_POINTS 14
#define G1S_NUM_UV_POINTS 10
#define G1S_NUM_Y_COEFFS 24
#define G1S_NUM_UV_COEFFS 25

enum g1s_status {
  G1S_OK = 0,
  G1S_E_ARG = -1,         /* bad argument / unsupported configuration            */
  G1S_E_DIMS = -2,        /* source and denoised frame dimensions differ          */
  G1S_E_CUDA = -3,        /* CUDA runtime failure or no device (no CPU fallback)  */
  G1S_E_NCCL = -4,        /* multi-device handle: a peer device failed / is unusable */
  G1S_E_NOMEM = -5,
  G1S_E_STATE = -6,       /* call after finish, or capacity too small             */
  G1S_E_IO = -7,
  G1S_E_STREAM = -8       /* malformed AV1 / IVF data (inspect path)               */
};

/* One grain-table segment: field-for-field the data carried by
 * av1_grain::GrainTableSegment as consumed at src/parser/grain.rs:108-133 and
 * src/main.rs:705-713.  POD, caller-allocated. */
typedef struct g1s_segment {
  uint64_t start_time; /* 1e-7 s units */
  uint64_t end_time;
  uint8_t num_y_points, num_cb_points, num_cr_points;
  uint8_t scaling_shift;      /* 8..=11 */
  uint8_t ar_coeff_lag;       /* always 3 on this path */
  uint8_t ar_coeff_shift;     /* 6..=9 */
  uint8_t grain_scale_shift;
  uint8_t overlap_flag;
  uint8_t chroma_scaling_from_luma;
  uint8_t cb_mult, cb_luma_mult, cr_mult, cr_luma_mult;
  /* Explicit AR coefficient counts + 1 for Y, Cb, Cr; 0 = "follow the lag" (24/25/25 at lag 3), which is what the
   * diff path emits.  The inspect path needs them: a parsed header carries no luma coefficients when it has no luma
   * points and a single 0 for a chroma plane without coefficients (src/parser/grain.rs:221-243). */
  uint8_t num_ar_coeffs_plus1[3];
  uint16_t cb_offset, cr_offset;
  uint16_t random_seed;
  uint16_t clip_to_restricted_range; /* inspect path only (src/parser/grain.rs:65); the table text does not carry it */
  uint8_t scaling_points_y[G1S_NUM_Y_POINTS][2];
  uint8_t scaling_points_cb[G1S_NUM_UV_POINTS][2];
  uint8_t scaling_points_cr[G1S_NUM_UV_POINTS][2];
  int8_t ar_coeffs_y[G1S_NUM_Y_COEFFS];
  int8_t ar_coeffs_cb[G1S_NUM_UV_COEFFS];
  int8_t ar_coeffs_cr[G1S_NUM_UV_COEFFS];
} g1s_segment;

/* A borrowed planar frame (what `&Frame<T>` is at src/main.rs:442): Y, Cb, Cr.
 * Samples are uint8_t when the stream's bit depth is 8, little-endian uint16_t
 * when it is 9..16 (src/reader.rs:51-67).  Strides are in BYTES.  For monochrome
 * streams plane[1] and plane[2] are NULL.  width/height are the luma size of THIS
 * frame: the reference checks source against denoised per call
 * (verify_dimensions_match) and fails the call when they differ. */
typedef struct g1s_frame {
  const void *plane[3];
  size_t stride_bytes[3];
  int32_t width, height;
} g1s_frame;

/* How a handle is used.  FULL is the drop-in for DiffGenerator.  The other two split
 * it for frame-sharded multi-GPU runs: every rank runs a PRODUCER (kernels only; the
 * per-frame integer records leave through the record tap), the records are exchanged
 * (NCCL) and ONE rank feeds them, in frame order, to a CONSUMER (host model only, no
 * device work) with g1s_diff_consume_record. */
enum g1s_mode { G1S_MODE_FULL = 0, G1S_MODE_PRODUCER = 1, G1S_MODE_CONSUMER = 2 };

/* How the AR normal equations (A[i][j] += buf[i]*buf[j] / 255^2, the inner loop behind diff_frame,
 * src/main.rs:442) are accumulated.
 *   EXACT_INT  exact integer sums on the int8 tensor cores, one division per entry at the end.  More accurate
 *              than the reference and ~1e-13 away from it; the grain table can differ from the reference's in
 *              the choice of scaling points where fit_piecewise meets a tie (about 5 % of streams).
 *   REF_ORDER  strict mode: every entry is accumulated term by term in f64, in the reference's pixel order and
 *              with its roundings (gram_reforder_kernel), so every integer of every table equals the
 *              reference's.  FP64-bound, roughly 20x slower than EXACT_INT, still >1000x the CPU. */
enum g1s_gram_order { G1S_GRAM_EXACT_INT = 0, G1S_GRAM_REF_ORDER = 1 };
enum g1s_model_placement { G1S_MODEL_AUTO = 0, G1S_MODEL_HOST = 1, G1S_MODEL_DEVICE = 2 };

typedef struct g1s_diff_config {
  int64_t fps_num, fps_den;     /* Rational64 passed to DiffGenerator::new           */
  int32_t src_bit_depth;        /* 8..16, bit depth of the source stream             */
  int32_t den_bit_depth;        /* 8..16, bit depth of the denoised stream           */
  int32_t width, height;        /* luma size in samples                              */
  int32_t ss_x, ss_y;           /* chroma subsampling log2 (4:2:0 = 1,1)             */
  int32_t monochrome;           /* non-zero: only plane[0] is read                   */
  int32_t device;               /* CUDA device ordinal                               */
  int32_t batch_frames;         /* frames per device launch; 0 = engine default      */
  int32_t mode;                 /* enum g1s_mode                                     */
  int32_t gram_kernel;          /* 0 = auto (int8 tensor-core kernel when the stream
                                   allows it), 1 = force the generic int32 kernel    */
  int32_t host_threads;         /* threads evaluating the per-frame half of the host model;
                                   0 = auto (half the cores, at most 8)              */
  int32_t host_narrow;          /* non-zero: host frames wider than 8 bits are reduced to 8 bits (the truncating
                                   `>> (bit_depth - 8)` of frame_into_u8, which is all the path ever reads) by the
                                   staging threads, halving the bytes that cross PCIe; results are identical.
                                   Such a handle does not take device frames.  Default 0.   */
  int32_t gram_order;           /* enum g1s_gram_order; default 0 = EXACT_INT (fast)                  */
  int32_t n_devices;            /* 0 or 1: one GPU (`device`).  2..G1S_MAX_DEVICES: ONE handle drives that many GPUs
                                   of this process: batches of frames are dealt round-robin to
                                   device_ids[0..n_devices), every device runs the kernels and the per-frame half of
                                   the model, the per-frame digests are folded in frame order on the caller's
                                   process.  Same table as one GPU.  mode must be G1S_MODE_FULL.          */
  int32_t device_ids[8];        /* CUDA ordinals, used when n_devices >= 2                              */
  int32_t model_placement;      /* where the per-frame half of the noise model (AR solves, strength measurements and
                                   solve: 113 us of one host core per 4K frame) runs.  G1S_MODEL_AUTO: on the host
                                   threads, unless this handle drives more GPUs than the host has cores for (3 cores
                                   per GPU), then on the device; G1S_MODEL_HOST; G1S_MODEL_DEVICE (latest_kernel: only
                                   11 KB digests cross PCIe, no per-frame host work -- what a box with 8 GPUs and 32
                                   cores needs; costs ~8 % of the GPU).  Results are bit-identical.          */
  int32_t reserved_[3];
} g1s_diff_config;
#define G1S_MAX_DEVICES 8

typedef struct g1s_diff g1s_diff;

/* DiffGenerator::new — src/main.rs:420-427. */
int g1s_diff_create(const g1s_diff_config *cfg, g1s_diff **out);

/* DiffGenerator::diff_frame — src/main.rs:442.  Planes are borrowed only until
 * the call returns. */
int g1s_diff_push_frame(g1s_diff *d, const g1s_frame *source, const g1s_frame *denoised);

/* Same as push_frame but the planes are DEVICE pointers (already resident in HBM,
 * same layout); used by the benchmark's device-resident arm and by callers that
 * decode on the GPU.  The memory must stay valid until the next g1s_diff_flush /
 * g1s_diff_finish returns. */
int g1s_diff_push_frame_device(g1s_diff *d, const g1s_frame *source, const g1s_frame *denoised);

/* ---- source filters (`diff --filters`) ----------------------------------------------------------
 * The reference applies its FilterChain to the SOURCE frame only, before diff_frame (src/main.rs:621-624 ->
 * src/filters.rs:112-182 -> video_resize::{crop, resize}).  Here the chain runs on the device between the upload of the
 * source planes and the kernels of the path: crop is a pointer / pitch offset, resize two separable passes
 * (csrc/g1s_filters.cu; the resampling arithmetic of the un-vendored crate video-resize 0.2.0 cannot be pinned here,
 * the published kernels and filter construction it ports are restated -- parity UNPINNED for resize, exact for crop).
 * The caller parses the filter string (the reference's grammar, src/filters.rs:16-110) and passes the operations. */
enum g1s_filter_kind { G1S_FILTER_CROP = 0, G1S_FILTER_RESIZE = 1 };
enum g1s_resize_alg { G1S_RESIZE_HERMITE = 0, G1S_RESIZE_CATMULLROM = 1, G1S_RESIZE_MITCHELL = 2, G1S_RESIZE_LANCZOS = 3,
                      G1S_RESIZE_SPLINE36 = 4 };
typedef struct g1s_filter_op {
  int32_t kind;        /* enum g1s_filter_kind                                                      */
  int32_t a, b, c, d;  /* crop: top, bottom, left, right (luma samples); resize: width, height, enum g1s_resize_alg, 0 */
} g1s_filter_op;
/* Call once, before the first frame.  From then on `source` frames passed to g1s_diff_push_frame[_device] must be
 * src_width x src_height (the size BEFORE the filters); the chain's output size must equal the handle's width x height
 * (what the denoised frames have), else G1S_E_DIMS -- the reference fails the same way, in verify_dimensions_match. */
int g1s_diff_set_source_filters(g1s_diff *d, const g1s_filter_op *ops, size_t n, int32_t src_width, int32_t src_height);
/* Test hook: the weights of one resize axis as the device uses them.  left[dst] first source sample per output sample,
 * coef[dst * taps] f32 weights; returns taps (<= max_taps) or a negative status. */
int g1s_resize_table(int32_t alg, int32_t src, int32_t dst, int32_t *left, float *coef, int32_t max_taps);

/* Drains everything queued so far (blocks until the device and the host model have
 * consumed every pushed frame).  Optional; finish implies it. */
int g1s_diff_flush(g1s_diff *d);

/* DiffGenerator::finish — src/main.rs:524.  Writes up to `cap` segments, stores the
 * number of segments in *n.  Returns G1S_E_STATE if cap is too small (then *n is the
 * required capacity and the call may be repeated). */
int g1s_diff_finish(g1s_diff *d, g1s_segment *out, size_t cap, size_t *n);

void g1s_diff_destroy(g1s_diff *d);

/* Text of the last error on this handle (or of the last failed create when d is
 * NULL).  Never NULL. */
const char *g1s_diff_last_error(const g1s_diff *d);

/* Number of frames accepted so far. */
int64_t g1s_diff_frames_pushed(const g1s_diff *d);
/* Frames per device launch of this handle (the dealing unit of a multi-device handle). */
int g1s_diff_batch_frames(const g1s_diff *d);
/* Where the per-frame half of the noise model runs for this handle (g1s_model_placement resolved): 1 on the device
 * (latest_kernel, digests cross PCIe), 0 on the host threads (records cross PCIe). */
int g1s_diff_model_on_device(const g1s_diff *d);
/* CUDA ordinal that processes frame `frame_index` (counted from 0): `device` for single-device handles,
 * device_ids[(frame_index / batch) % n_devices] for multi-device ones.  g1s_diff_push_frame_device on a multi-device
 * handle needs the planes of frame k resident on that device. */
int g1s_diff_frame_device(const g1s_diff *d, int64_t frame_index);

/* Timing/launch counters of the device pipeline, for the benchmark harness:
 * out[0] = kernels launched, out[1] = total device ms spent in the fused
 * residual+autocorrelation kernel (CUDA events on the engine stream),
 * out[2] = its launch count, out[3] = device ms of the flat-block kernel,
 * out[4] = its launch count, out[5] = frames fully processed, out[6] = batches that took
 * the int8 tensor-core path (residual kernel + TMA-fed Gram kernel), out[7] = device ms of the
 * residual kernel, out[8] = batches whose planes were 16-byte aligned (128-bit loads), out[9] = device ms of the
 * strict-mode (reference-order) Gram kernel. */
int g1s_diff_get_counters(const g1s_diff *d, double *out, size_t n);

/* Records CUDA event `which` (0 or 1) on the engine's kernel stream; g1s_diff_marks_elapsed_ms returns the
 * device time between them (benchmark timing on the stream the kernels are launched on). */
int g1s_diff_mark(g1s_diff *d, int which);
double g1s_diff_marks_elapsed_ms(g1s_diff *d);

/* ---- per-frame records (the unit exchanged between GPUs) -------------------------
 * One record holds everything the host model needs from one frame pair, all integers
 * except the f32 flatness scores: gram[3][351] int64 (upper triangle over 26 taps:
 * 0..23 AR taps, 24 luma tap x 2^(ss_x+ss_y), 25 centre sample), nobs[3] int64,
 * num_flat int64, luma_sum[nb] u32, rsum[3][nb] i32, rsq[3][nb] u32, score[nb] f32,
 * flat[nb] u8.  g1s_record_layout fills off[0..7] with the byte offsets of those eight
 * arrays in that order and returns the record size. */
size_t g1s_record_layout(int32_t num_blocks, size_t off[8]);
/* Strict mode (gram_order = G1S_GRAM_REF_ORDER) adds gramf[3][351] f64 to the record: the same tap pairs as gram,
 * accumulated term by term as RN(acc + RN(product / 255^2)) in the reference's pixel order, products through the
 * luma tap unscaled.  Byte offset of that array (zeros in EXACT_INT mode). */
size_t g1s_record_gramf_offset(int32_t num_blocks);
size_t g1s_diff_record_bytes(const g1s_diff *d);
typedef void (*g1s_record_fn)(void *user, int64_t frame_index, const void *record, size_t bytes);
/* Called on the caller's thread (inside push/flush/finish) once per frame, in frame
 * order, before the record is folded into the model. */
int g1s_diff_set_record_tap(g1s_diff *d, g1s_record_fn fn, void *user);
/* CONSUMER handles: fold one record (next frame in order) into the model. */
int g1s_diff_consume_record(g1s_diff *d, const void *record, size_t bytes);
/* Same for `count` consecutive frames, record k at records + k * stride_bytes; the per-frame half of
 * the model is evaluated on the handle's host threads, the merge stays in frame order. */
int g1s_diff_consume_records(g1s_diff *d, const void *records, size_t count, size_t stride_bytes);

/* ---- per-frame digests (what scales across GPUs) ------------------------------------
 * The host model splits into a per-frame half (a pure function of one record: AR solve, strength
 * measurements, strength solve) and a sequential merge.  A PRODUCER handle with a digest sink runs
 * the per-frame half itself and writes one fixed-size digest (g1s_digest_bytes(), ~11 KB, independent
 * of the frame size) per retired frame into the caller's buffer, in frame order; the rank that owns
 * the model folds them with g1s_diff_consume_digests.  Same results as exchanging the full records,
 * 27x less traffic and no per-frame work left on the sequential rank but the merge itself. */
size_t g1s_digest_bytes(void);
/* buffer must hold capacity_frames digests and stay valid while it is the sink; resets the count.
 * The sink is a ring: digest k (counted from the reset) lands in slot k % capacity_frames, so a reader
 * that keeps up never has to reset it. */
int g1s_diff_set_digest_sink(g1s_diff *d, void *buffer, size_t capacity_frames);
/* digests written into the sink since it was (re)set (monotonic). */
int64_t g1s_diff_digest_count(const g1s_diff *d);
/* Blocks until at least `frames` frames (counted from creation) have left the device pipeline and, for
 * sinks and taps, been delivered -- without forcing a partially filled batch out as flush does.
 * Returns early if fewer frames are in flight. */
int g1s_diff_wait_retired(g1s_diff *d, int64_t frames);
int g1s_diff_consume_digests(g1s_diff *d, const void *digests, size_t count);
/* Same without the copy: the digests are read in place by the fold thread, so the memory must stay
 * valid and unchanged until g1s_diff_flush on this (CONSUMER) handle has returned. */
int g1s_diff_consume_digests_borrowed(g1s_diff *d, const void *digests, size_t count);
/* Evaluates the per-frame half for one record without touching the model (any handle of the stream's
 * geometry; no device work) and writes its digest. */
int g1s_diff_digest_from_record(g1s_diff *d, const void *record, size_t bytes, void *digest_out);

/* `filmgrn1` writer — src/main.rs:525-530 and 631-696, byte for byte. */
int g1s_write_grain_table(const g1s_segment *segs, size_t n, const char *path);
/* Same, into a caller buffer.  Returns the number of bytes needed (excluding the
 * terminating NUL) or a negative status. */
int64_t g1s_format_grain_table(const g1s_segment *segs, size_t n, char *buf, size_t cap);

int g1s_abi_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * `inspect` (BASELINE configs[0]): AV1 OBU headers -> film grain headers -> grain table.
 * CPU only, exactly like the reference: replaces BitstreamParser::<false>::get_grain_headers
 * (src/parser.rs:120-173, fed by FFmpeg packets there) and aggregate_grain_headers (src/main.rs:713-772) as driven
 * by the `Inspect` arm (src/main.rs:172-196).  A packet is one demuxed sample: whole OBUs of one temporal unit. */
typedef struct g1s_inspect g1s_inspect;

enum g1s_grain_kind { G1S_GRAIN_DISABLE = 0, G1S_GRAIN_COPY_REF_FRAME = 1, G1S_GRAIN_UPDATE = 2 }; /* FilmGrainHeader */

typedef struct g1s_stream_info {
  int32_t have_sequence_header, seq_profile, bit_depth, monochrome, ss_x, ss_y;
  int32_t max_frame_width, max_frame_height, film_grain_params_present;
  int32_t color_primaries, transfer_characteristics, matrix_coefficients, color_range;
  int32_t order_hint_bits, reduced_still_picture_header, reserved_;
  uint64_t packets, obus;
  /* the most recent frame header: coded size (after super-resolution down-scaling) and tile grid */
  int32_t last_frame_width, last_frame_height, last_tile_cols, last_tile_rows;
} g1s_stream_info;

int g1s_inspect_create(g1s_inspect **out);
void g1s_inspect_destroy(g1s_inspect *h);
const char *g1s_inspect_last_error(const g1s_inspect *h);
/* src/parser.rs:136-165: parse every OBU of the packet; one header is recorded per SHOWN frame header. */
int g1s_inspect_push_packet(g1s_inspect *h, const uint8_t *data, size_t size);
/* Demux an IVF file or a Section-5 low-overhead .obu stream and push its packets.  *fps_num / *fps_den <= 0 on entry
 * are replaced by the IVF header's rate / scale (the reference asks FFmpeg: src/main.rs:173). */
int g1s_inspect_push_file(g1s_inspect *h, const char *path, int64_t *fps_num, int64_t *fps_den);
size_t g1s_inspect_num_headers(const g1s_inspect *h);
/* kind: enum g1s_grain_kind; params (may be NULL) is meaningful for G1S_GRAIN_UPDATE, random_seed = grain_seed. */
int g1s_inspect_header(const g1s_inspect *h, size_t i, int32_t *kind, g1s_segment *params);
/* aggregate_grain_headers: one packet per header, time_per_packet = den / num * 1e7 accumulated in f64 and ceiled.
 * *n == 0 with G1S_OK means "no film grain headers found" (src/main.rs:177-183). G1S_E_STATE: cap too small. */
int g1s_inspect_finish(g1s_inspect *h, int64_t fps_num, int64_t fps_den, g1s_segment *out, size_t cap, size_t *n);
int g1s_inspect_stream_info(const g1s_inspect *h, g1s_stream_info *info);
/* `apply` / `remove`: BitstreamParser::<true>::modify_grain_headers (src/parser.rs:175-348) and the
 * header rewriting of src/parser/frame.rs:608-826 / sequence.rs:404-423, without the FFmpeg muxer: one packet in, one
 * packet out.  apply != 0: every frame header that may carry film grain gets the parameters of the table segment with
 * start_time <= packet_ts < end_time (apply_grain = 1, update_grain = 1 on inter frames, seed advanced by
 * DEFAULT_GRAIN_SEED per frame), or apply_grain = 0 when no segment covers it; apply == 0 (table ignored): film grain
 * is removed (film_grain_params_present = 0, grain syntax dropped).  OBU sizes are re-encoded.  packet_ts is in 1e-7 s
 * (src/parser.rs:103-118).  The handle also collects the ORIGINAL grain headers like an inspect handle. */
int g1s_rewrite_create(const g1s_segment *table, size_t n, int apply, g1s_inspect **out);
int g1s_rewrite_packet(g1s_inspect *h, const uint8_t *data, size_t size, uint64_t packet_ts, size_t *out_size);
/* copies the packet produced by the last g1s_rewrite_packet call (G1S_E_STATE: cap < *out_size) */
int g1s_rewrite_take(g1s_inspect *h, uint8_t *out, size_t cap);
int g1s_rewrite_counters(const g1s_inspect *h, uint64_t *frames_with_grain, uint64_t *frames_grain_disabled);
/* `generate` (src/main.rs:247-308): av1_grain::generate_photon_noise_params -- a photon-noise grain segment for a
 * camera ISO setting, frame size and transfer function (luma scaling points only, lag 0); apply it with g1s_rewrite_*.
 * full_range: NoiseGenArgs::full_range, which the reference sets from the stream's colour range (src/main.rs:299); non-zero
 * places the scaling points on the 0..255 scale, zero on 16..235 (that branch is recalled from the crate: unpinned).
 * random_seed < 0 takes DEFAULT_GRAIN_SEED.  G1S_TRANSFER_BT470BG is not reachable from the reference's CLI: it is the
 * curve behind the reference's fixture tests/example-table.tbl and is kept to reproduce it. */
enum g1s_transfer { G1S_TRANSFER_BT1886 = 0, G1S_TRANSFER_SMPTE2084 = 1, G1S_TRANSFER_BT470BG = 2 };
int g1s_generate_photon_noise(uint32_t iso, uint32_t width, uint32_t height, int transfer, int chroma_grain,
                              int full_range, int32_t random_seed, uint64_t start_time, uint64_t end_time,
                              g1s_segment *out);
/* Host-side reduction used by host_narrow (exported for tests): dst[i] = (uint8_t)(src[i] >> shift).  dst is meant to
 * be staging memory nothing reads back soon: on x86-64 with AVX2 / AVX-512BW the bytes go out with streaming stores.
 * g1s_narrow_isa: 0 compiler loop, 1 AVX2, 2 AVX-512BW; g1s_narrow_row_with runs a path at or below that. */
void g1s_narrow_row(uint8_t *dst, const uint16_t *src, int n, int shift);
void g1s_narrow_row_with(uint8_t *dst, const uint16_t *src, int n, int shift, int isa);
int g1s_narrow_isa(void);
/* Test hook: the host model's linear solvers on caller data (util.rs::linsolve semantics, bit for bit).  which 0: the
 * elimination as the model runs it, 1: the tridiagonal fast path of the noise strength systems alone.  Returns 1 solved,
 * -1 the reference's failure (pivot below 1e-16), 0 (which = 1 only) a row swap would be needed: the model then runs the
 * dense elimination.  x must come in as the caller's x (the reference leaves the rows it did not reach untouched). */
int g1s_linsolve_probe(int which, int n, const double *A, const double *b, double *x);
/* Test hook: parse ONE syntax group (named as in the AV1 spec / the reference's functions) from a raw bit buffer;
 * returns bits consumed or a negative status.  Lets tests replay the reference's own unit-test vectors. */
int64_t g1s_obu_probe(const char *what, const uint8_t *data, size_t size, const int64_t *args, size_t nargs,
                      int64_t *out, size_t nout, g1s_segment *seg);

#ifdef __cplusplus
}
#endif
#endif /* G1S_H_ */
