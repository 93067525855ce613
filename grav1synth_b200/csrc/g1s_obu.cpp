// g1s_obu.cpp — AV1 OBU header walk for `grav1synth inspect` (BASELINE configs[0]).
//
// CPU only, no device work: the reference does this on the CPU too.  What is restated (behaviour, not code):
//   packet loop / header collection      /root/reference/src/parser.rs:120-173  (get_grain_headers)
//   OBU header, size, layer filter       /root/reference/src/parser/obu.rs:43-250, 318-379
//   sequence_header_obu                  /root/reference/src/parser/sequence.rs:163-653
//   frame_header_obu, uncompressed_header /root/reference/src/parser/frame.rs:75-699, 865-1991
//   film_grain_params                    /root/reference/src/parser/grain.rs:136-295
//   tile group header (end-of-frame)     /root/reference/src/parser/tile_group.rs:11-52
//   aggregate_grain_headers              /root/reference/src/main.rs:713-772
// i.e. AV1 bitstream spec sections 5.3 (OBU), 5.5 (sequence header), 5.9 (frame header), 5.11.1 (tile group).
// The reference gets packets from FFmpeg; here a packet is whatever the caller pushes (g1s_inspect_push_packet),
// and g1s_inspect_file demuxes IVF and Section-5 ("low overhead") .obu files itself.
//
// Where the reference simplifies the spec harmlessly the same simplification is kept, so that both read the same
// bits (OBU_REDUNDANT_FRAME_HEADER is not parsed).
// Deliberate differences, all where the reference cannot continue or would read bits the stream does not have (each
// is exercised on libaom-encoded streams in tests/test_inspect_libaom.py or on hand-built ones in test_inspect.py):
//  * a standalone OBU_TILE_GROUP is an `unreachable!()` there (obu.rs:215-219); here its header is read to find the
//    end of the frame;
//  * frame_refs_short_signaling runs the spec's set_frame_refs process (7.8), which the reference stubs out;
//  * found_ref takes the frame size of the referenced slot (the reference keeps the sequence maximum);
//  * UpscaledWidth is tracked, so that allow_intrabc and the loop-restoration parameters of a super-resolved frame
//    follow the spec (5.9.2, 5.9.20);
//  * show_existing_frame of a key frame refreshes the reference slots (7.21), which forward key frames rely on;
//  * segmentation features are inherited from the primary reference frame when they are not re-sent (7.20);
//  * ref_order_hint[i] of an error-resilient frame replaces the slot's saved order hint (5.9.2);
//  * with uniform tile spacing the tile counts are ceil(sb / tile size) as in 5.9.15 (the reference floors).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/g1s.h"

namespace {

struct ParseError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------- bit reader (spec 4.10: f(n), ns, su, uvlc, leb128)
struct BitReader {
  const uint8_t *p;
  size_t nbits;
  size_t pos = 0;
  BitReader(const uint8_t *d, size_t nbytes) : p(d), nbits(nbytes * 8) {}
  uint64_t f(unsigned n) {
    if (n > 64) throw ParseError("field wider than 64 bits");
    if (pos + n > nbits) throw ParseError("unexpected end of data inside an OBU");
    uint64_t v = 0;
    for (unsigned i = 0; i < n; ++i, ++pos) v = (v << 1) | ((p[pos >> 3] >> (7 - (pos & 7))) & 1u);
    return v;
  }
  bool flag() { return f(1) != 0; }
  int64_t su(unsigned n) {  // signed, n bits including the sign (spec 4.10.6)
    const int64_t v = (int64_t)f(n);
    const int64_t sign = (int64_t)1 << (n - 1);
    return (v & sign) ? v - 2 * sign : v;
  }
  uint64_t ns(uint64_t n) {  // non-symmetric unsigned (spec 4.10.7)
    unsigned w = 0;
    for (uint64_t x = n; x; x >>= 1) ++w;  // FloorLog2(n) + 1
    const uint64_t m = ((uint64_t)1 << w) - n;
    const uint64_t v = f(w - 1);
    if (v < m) return v;
    return (v << 1) - m + f(1);
  }
  uint32_t uvlc() {
    unsigned lz = 0;
    while (!flag()) ++lz;
    if (lz >= 32) return 0xFFFFFFFFu;
    return (uint32_t)f(lz) + (uint32_t)(((uint64_t)1 << lz) - 1);
  }
  void byte_alignment(bool verify_zero) {
    while (pos & 7) {
      if (f(1) && verify_zero) throw ParseError("non-zero bit inside byte_alignment()");
    }
  }
  size_t bytes_consumed() const { return (pos + 7) >> 3; }
};

// leb128 as the reference reads it (util.rs:49-72): at most 8 bytes
bool read_leb128(const uint8_t *d, size_t n, uint64_t *value, size_t *used) {
  uint64_t v = 0;
  for (unsigned i = 0; i < 8; ++i) {
    if (i >= n) return false;
    v |= (uint64_t)(d[i] & 0x7f) << (i * 7);
    if (!(d[i] & 0x80)) {
      *value = v;
      *used = i + 1;
      return true;
    }
  }
  *value = v;
  *used = 8;
  return true;
}

enum { OBU_SEQUENCE_HEADER = 1, OBU_TEMPORAL_DELIMITER = 2, OBU_FRAME_HEADER = 3, OBU_TILE_GROUP = 4, OBU_FRAME = 6 };
enum { KEY_FRAME = 0, INTER_FRAME = 1, INTRA_ONLY_FRAME = 2, SWITCH_FRAME = 3 };
enum { SELECT_SCREEN_CONTENT_TOOLS = 2, SELECT_INTEGER_MV = 2, PRIMARY_REF_NONE = 7 };
constexpr int REFS_PER_FRAME = 7, NUM_REF_FRAMES = 8, MAX_SEGMENTS = 8, SEG_LVL_MAX = 8;

struct SequenceHeader {
  bool valid = false;
  bool reduced_still_picture_header = false, frame_id_numbers_present = false;
  int additional_frame_id_len_minus_1 = 0, delta_frame_id_len_minus_2 = 0;
  bool film_grain_params_present = false;
  int force_screen_content_tools = 0, force_integer_mv = 0, order_hint_bits = 0;
  int frame_width_bits_minus_1 = 0, frame_height_bits_minus_1 = 0;
  uint32_t max_frame_width_minus_1 = 0, max_frame_height_minus_1 = 0;
  bool has_decoder_model = false;
  int buffer_delay_length_minus_1 = 0, buffer_removal_time_length_minus_1 = 0,
      frame_presentation_time_length_minus_1 = 0;
  bool has_timing_info = false, equal_picture_interval = false;
  int operating_points_cnt_minus_1 = 0;
  uint16_t operating_point_idc[32] = {0};
  bool decoder_model_present_for_op[32] = {false};
  uint16_t cur_operating_point_idc = 0;
  bool enable_ref_frame_mvs = false, enable_warped_motion = false, enable_superres = false, enable_cdef = false,
       enable_restoration = false, use_128x128_superblock = false;
  int seq_profile = 0, bit_depth = 8, num_planes = 3, ss_x = 1, ss_y = 1;
  bool separate_uv_delta_q = false;
  int color_primaries = 2, transfer_characteristics = 2, matrix_coefficients = 2, color_range = 0;
};

struct TileInfo {
  uint32_t tile_cols = 1, tile_rows = 1, tile_cols_log2 = 0, tile_rows_log2 = 0;
};

struct GrainHeader {
  int kind = G1S_GRAIN_DISABLE;
  g1s_segment params;  // start/end unused here
  GrainHeader() { std::memset(&params, 0, sizeof params); }
};

struct FrameHeader {
  bool show_frame = false, show_existing_frame = false;
  GrainHeader grain;
  TileInfo tile_info;
  // for the rewriter (apply / remove): where film_grain_params() starts, whether it may be present, the frame type
  size_t grain_pos = 0;
  bool film_grain_allowed = false;
  int frame_type = 0;
};

// MSB-first bit sink for rewritten headers
struct BitWriter {
  std::vector<uint8_t> bytes;
  size_t nbits = 0;
  void put(uint64_t v, unsigned n) {
    for (unsigned i = n; i-- > 0;) {
      if ((nbits & 7) == 0) bytes.push_back(0);
      if ((v >> i) & 1) bytes.back() |= (uint8_t)(0x80u >> (nbits & 7));
      ++nbits;
    }
  }
  void copy_bits(const uint8_t *src, size_t count) {
    for (size_t i = 0; i < count; ++i) put((src[i >> 3] >> (7 - (i & 7))) & 1u, 1);
  }
};

constexpr uint16_t DEFAULT_GRAIN_SEED = 10956;  // av1_grain::DEFAULT_GRAIN_SEED (frame.rs:3, :637-640)

// write_film_grain_bits (frame.rs:700-826): apply_grain = 1 and every parameter of `p`, update_grain = 1 on inter frames
void write_film_grain_bits(BitWriter &bw, const g1s_segment &p, int frame_type, bool monochrome, int ss_x, int ss_y) {
  bw.put(1, 1);
  bw.put(p.random_seed, 16);
  if (frame_type == INTER_FRAME) bw.put(1, 1);
  bw.put(p.num_y_points, 4);
  for (int i = 0; i < p.num_y_points; ++i) bw.put(p.scaling_points_y[i][0], 8), bw.put(p.scaling_points_y[i][1], 8);
  const bool csfl = monochrome ? false : p.chroma_scaling_from_luma != 0;
  if (!monochrome) bw.put(csfl, 1);
  int ncb = 0, ncr = 0;
  if (!(monochrome || csfl || (ss_x == 1 && ss_y == 1 && p.num_y_points == 0))) {
    ncb = p.num_cb_points, ncr = p.num_cr_points;
    bw.put((unsigned)ncb, 4);
    for (int i = 0; i < ncb; ++i) bw.put(p.scaling_points_cb[i][0], 8), bw.put(p.scaling_points_cb[i][1], 8);
    bw.put((unsigned)ncr, 4);
    for (int i = 0; i < ncr; ++i) bw.put(p.scaling_points_cr[i][0], 8), bw.put(p.scaling_points_cr[i][1], 8);
  }
  bw.put((unsigned)(p.scaling_shift - 8) & 3, 2);
  bw.put(p.ar_coeff_lag & 3u, 2);
  const int lag = p.ar_coeff_lag & 3;
  const int num_pos_luma = 2 * lag * (lag + 1);
  int num_pos_chroma = num_pos_luma;
  if (p.num_y_points > 0) {
    for (int i = 0; i < num_pos_luma; ++i) bw.put((uint8_t)(p.ar_coeffs_y[i] + 128), 8);
    num_pos_chroma = num_pos_luma + 1;
  }
  if (csfl || ncb > 0)
    for (int i = 0; i < num_pos_chroma; ++i) bw.put((uint8_t)(p.ar_coeffs_cb[i] + 128), 8);
  if (csfl || ncr > 0)
    for (int i = 0; i < num_pos_chroma; ++i) bw.put((uint8_t)(p.ar_coeffs_cr[i] + 128), 8);
  bw.put((unsigned)(p.ar_coeff_shift - 6) & 3, 2);
  bw.put(p.grain_scale_shift & 3u, 2);
  if (ncb > 0) bw.put(p.cb_mult, 8), bw.put(p.cb_luma_mult, 8), bw.put(p.cb_offset & 0x1FFu, 9);
  if (ncr > 0) bw.put(p.cr_mult, 8), bw.put(p.cr_luma_mult, 8), bw.put(p.cr_offset & 0x1FFu, 9);
  bw.put(p.overlap_flag ? 1 : 0, 1);
  bw.put(p.clip_to_restricted_range ? 1 : 0, 1);
}

// FilmGrainParams equality as the reference defines it (grain.rs:83-105): everything but the seed.
bool same_grain(const g1s_segment &a, const g1s_segment &b) {
  auto pts = [](const uint8_t(*x)[2], int nx, const uint8_t(*y)[2], int ny) {
    return nx == ny && std::memcmp(x, y, (size_t)nx * 2) == 0;
  };
  // num_ar_coeffs_plus1 is the count + 1, or 0 for "follow the lag" (2 * lag * (lag + 1), +1 for chroma)
  const int lag_n = 2 * a.ar_coeff_lag * (a.ar_coeff_lag + 1);
  auto co = [&](const int8_t *x, int px, const int8_t *y, int py, int follow, int cap) {
    if (px != py) return false;
    const int n = std::min(px ? px - 1 : follow, cap);
    return std::memcmp(x, y, (size_t)n) == 0;
  };
  return pts(a.scaling_points_y, a.num_y_points, b.scaling_points_y, b.num_y_points) &&
         pts(a.scaling_points_cb, a.num_cb_points, b.scaling_points_cb, b.num_cb_points) &&
         pts(a.scaling_points_cr, a.num_cr_points, b.scaling_points_cr, b.num_cr_points) &&
         a.scaling_shift == b.scaling_shift && a.ar_coeff_lag == b.ar_coeff_lag &&
         co(a.ar_coeffs_y, a.num_ar_coeffs_plus1[0], b.ar_coeffs_y, b.num_ar_coeffs_plus1[0], lag_n, G1S_NUM_Y_COEFFS) &&
         co(a.ar_coeffs_cb, a.num_ar_coeffs_plus1[1], b.ar_coeffs_cb, b.num_ar_coeffs_plus1[1], lag_n + 1, G1S_NUM_UV_COEFFS) &&
         co(a.ar_coeffs_cr, a.num_ar_coeffs_plus1[2], b.ar_coeffs_cr, b.num_ar_coeffs_plus1[2], lag_n + 1, G1S_NUM_UV_COEFFS) &&
         a.ar_coeff_shift == b.ar_coeff_shift &&
         a.cb_mult == b.cb_mult && a.cb_luma_mult == b.cb_luma_mult && a.cb_offset == b.cb_offset &&
         a.cr_mult == b.cr_mult && a.cr_luma_mult == b.cr_luma_mult && a.cr_offset == b.cr_offset &&
         a.chroma_scaling_from_luma == b.chroma_scaling_from_luma && a.grain_scale_shift == b.grain_scale_shift &&
         a.overlap_flag == b.overlap_flag && a.clip_to_restricted_range == b.clip_to_restricted_range;
}

// ---------------------------------------------------------------- film_grain_params (spec 5.9.30, grain.rs:136-295)
GrainHeader film_grain_params(BitReader &br, bool allowed, int frame_type, bool monochrome, int ss_x, int ss_y) {
  GrainHeader h;
  if (!allowed) return h;
  if (!br.flag()) return h;  // apply_grain
  g1s_segment &p = h.params;
  p.random_seed = (uint16_t)br.f(16);
  const bool update_grain = frame_type == INTER_FRAME ? br.flag() : true;
  if (!update_grain) {
    br.f(3);  // film_grain_params_ref_idx
    h.kind = G1S_GRAIN_COPY_REF_FRAME;
    return h;
  }
  h.kind = G1S_GRAIN_UPDATE;
  auto points = [&](uint8_t(*dst)[2], int cap, const char *what) -> int {
    const int n = (int)br.f(4);
    if (n > cap) throw ParseError(std::string("too many scaling points for ") + what);
    for (int i = 0; i < n; ++i) {
      dst[i][0] = (uint8_t)br.f(8);
      dst[i][1] = (uint8_t)br.f(8);
    }
    return n;
  };
  p.num_y_points = (uint8_t)points(p.scaling_points_y, G1S_NUM_Y_POINTS, "luma");
  p.chroma_scaling_from_luma = monochrome ? 0 : (uint8_t)br.flag();
  if (monochrome || p.chroma_scaling_from_luma || (ss_x == 1 && ss_y == 1 && p.num_y_points == 0)) {
    p.num_cb_points = p.num_cr_points = 0;
  } else {
    p.num_cb_points = (uint8_t)points(p.scaling_points_cb, G1S_NUM_UV_POINTS, "Cb");
    p.num_cr_points = (uint8_t)points(p.scaling_points_cr, G1S_NUM_UV_POINTS, "Cr");
  }
  p.scaling_shift = (uint8_t)(br.f(2) + 8);
  p.ar_coeff_lag = (uint8_t)br.f(2);
  const int num_pos_luma = 2 * p.ar_coeff_lag * (p.ar_coeff_lag + 1);
  int num_pos_chroma = num_pos_luma;
  int ny = 0, ncb = 0, ncr = 0;
  if (p.num_y_points > 0) {
    for (int i = 0; i < num_pos_luma; ++i) p.ar_coeffs_y[ny++] = (int8_t)((int)br.f(8) - 128);
    num_pos_chroma = num_pos_luma + 1;
  }
  // the reference keeps a single 0 where the bitstream carries no chroma coefficients (grain.rs:232-233, 241-242)
  if (p.chroma_scaling_from_luma || p.num_cb_points > 0) {
    for (int i = 0; i < num_pos_chroma; ++i) p.ar_coeffs_cb[ncb++] = (int8_t)((int)br.f(8) - 128);
  } else {
    p.ar_coeffs_cb[ncb++] = 0;
  }
  if (p.chroma_scaling_from_luma || p.num_cr_points > 0) {
    for (int i = 0; i < num_pos_chroma; ++i) p.ar_coeffs_cr[ncr++] = (int8_t)((int)br.f(8) - 128);
  } else {
    p.ar_coeffs_cr[ncr++] = 0;
  }
  p.num_ar_coeffs_plus1[0] = (uint8_t)(ny + 1);
  p.num_ar_coeffs_plus1[1] = (uint8_t)(ncb + 1);
  p.num_ar_coeffs_plus1[2] = (uint8_t)(ncr + 1);
  p.ar_coeff_shift = (uint8_t)(br.f(2) + 6);
  p.grain_scale_shift = (uint8_t)br.f(2);
  if (p.num_cb_points > 0) {
    p.cb_mult = (uint8_t)br.f(8);
    p.cb_luma_mult = (uint8_t)br.f(8);
    p.cb_offset = (uint16_t)br.f(9);
  }
  if (p.num_cr_points > 0) {
    p.cr_mult = (uint8_t)br.f(8);
    p.cr_luma_mult = (uint8_t)br.f(8);
    p.cr_offset = (uint16_t)br.f(9);
  }
  p.overlap_flag = (uint8_t)br.flag();
  p.clip_to_restricted_range = (uint16_t)br.flag();
  return h;
}

// ---------------------------------------------------------------- sequence header (spec 5.5, sequence.rs:163-653)
void color_config(BitReader &br, SequenceHeader &s) {
  const bool high_bitdepth = br.flag();
  if (s.seq_profile == 2 && high_bitdepth)
    s.bit_depth = br.flag() ? 12 : 10;
  else
    s.bit_depth = high_bitdepth ? 10 : 8;
  const bool monochrome = s.seq_profile == 1 ? false : br.flag();
  s.num_planes = monochrome ? 1 : 3;
  if (br.flag()) {  // color_description_present_flag
    s.color_primaries = (int)br.f(8);
    s.transfer_characteristics = (int)br.f(8);
    s.matrix_coefficients = (int)br.f(8);
  } else {
    s.color_primaries = s.transfer_characteristics = s.matrix_coefficients = 2;  // unspecified
  }
  if (monochrome) {
    s.color_range = (int)br.f(1);
    s.ss_x = s.ss_y = 1;
    s.separate_uv_delta_q = false;
    return;
  }
  if (s.color_primaries == 1 && s.transfer_characteristics == 13 && s.matrix_coefficients == 0) {  // sRGB
    s.color_range = 1;
    s.ss_x = s.ss_y = 0;
  } else {
    s.color_range = (int)br.f(1);
    if (s.seq_profile == 0) {
      s.ss_x = s.ss_y = 1;
    } else if (s.seq_profile == 1) {
      s.ss_x = s.ss_y = 0;
    } else if (s.bit_depth == 12) {
      s.ss_x = (int)br.f(1);
      s.ss_y = s.ss_x ? (int)br.f(1) : 0;
    } else {
      s.ss_x = 1;
      s.ss_y = 0;
    }
    if (s.ss_x && s.ss_y) br.f(2);  // chroma_sample_position
  }
  s.separate_uv_delta_q = br.flag();
}

SequenceHeader parse_sequence_header(BitReader &br) {
  SequenceHeader s;
  s.seq_profile = (int)br.f(3);
  br.flag();  // still_picture
  s.reduced_still_picture_header = br.flag();
  if (s.reduced_still_picture_header) {
    br.f(5);  // seq_level_idx[0]
    s.operating_points_cnt_minus_1 = 0;
    s.operating_point_idc[0] = 0;
    s.decoder_model_present_for_op[0] = false;
  } else {
    if (br.flag()) {  // timing_info_present_flag
      s.has_timing_info = true;
      br.f(32);  // num_units_in_display_tick
      br.f(32);  // time_scale
      s.equal_picture_interval = br.flag();
      if (s.equal_picture_interval) br.uvlc();  // num_ticks_per_picture_minus_1
      if (br.flag()) {                          // decoder_model_info_present_flag
        s.has_decoder_model = true;
        s.buffer_delay_length_minus_1 = (int)br.f(5);
        br.f(32);  // num_units_in_decoding_tick
        s.buffer_removal_time_length_minus_1 = (int)br.f(5);
        s.frame_presentation_time_length_minus_1 = (int)br.f(5);
      }
    }
    const bool initial_display_delay_present = br.flag();
    s.operating_points_cnt_minus_1 = (int)br.f(5);
    for (int i = 0; i <= s.operating_points_cnt_minus_1; ++i) {
      s.operating_point_idc[i] = (uint16_t)br.f(12);
      const int seq_level_idx = (int)br.f(5);
      if (seq_level_idx > 7) br.flag();  // seq_tier
      if (s.has_decoder_model) {
        s.decoder_model_present_for_op[i] = br.flag();
        if (s.decoder_model_present_for_op[i]) {
          const unsigned n = (unsigned)s.buffer_delay_length_minus_1 + 1;
          br.f(n);  // decoder_buffer_delay
          br.f(n);  // encoder_buffer_delay
          br.flag();  // low_delay_mode_flag
        }
      }
      if (initial_display_delay_present && br.flag()) br.f(4);  // initial_display_delay_minus_1
    }
  }
  s.cur_operating_point_idc = s.operating_point_idc[0];  // choose_operating_point() == 0
  s.frame_width_bits_minus_1 = (int)br.f(4);
  s.frame_height_bits_minus_1 = (int)br.f(4);
  s.max_frame_width_minus_1 = (uint32_t)br.f((unsigned)s.frame_width_bits_minus_1 + 1);
  s.max_frame_height_minus_1 = (uint32_t)br.f((unsigned)s.frame_height_bits_minus_1 + 1);
  s.frame_id_numbers_present = s.reduced_still_picture_header ? false : br.flag();
  if (s.frame_id_numbers_present) {
    s.delta_frame_id_len_minus_2 = (int)br.f(4);
    s.additional_frame_id_len_minus_1 = (int)br.f(3);
  }
  s.use_128x128_superblock = br.flag();
  br.flag();  // enable_filter_intra
  br.flag();  // enable_intra_edge_filter
  if (s.reduced_still_picture_header) {
    s.force_screen_content_tools = SELECT_SCREEN_CONTENT_TOOLS;
    s.force_integer_mv = SELECT_INTEGER_MV;
    s.order_hint_bits = 0;
  } else {
    br.flag();  // enable_interintra_compound
    br.flag();  // enable_masked_compound
    s.enable_warped_motion = br.flag();
    br.flag();  // enable_dual_filter
    const bool enable_order_hint = br.flag();
    if (enable_order_hint) {
      br.flag();  // enable_jnt_comp
      s.enable_ref_frame_mvs = br.flag();
    }
    const bool seq_choose_screen_content_tools = br.flag();
    s.force_screen_content_tools = seq_choose_screen_content_tools ? SELECT_SCREEN_CONTENT_TOOLS : (int)br.f(1);
    if (s.force_screen_content_tools > 0) {
      const bool seq_choose_integer_mv = br.flag();
      s.force_integer_mv = seq_choose_integer_mv ? SELECT_INTEGER_MV : (int)br.f(1);
    } else {
      s.force_integer_mv = SELECT_INTEGER_MV;
    }
    s.order_hint_bits = enable_order_hint ? (int)br.f(3) + 1 : 0;
  }
  s.enable_superres = br.flag();
  s.enable_cdef = br.flag();
  s.enable_restoration = br.flag();
  color_config(br, s);
  s.film_grain_params_present = br.flag();
  s.valid = true;
  return s;
}

// ---------------------------------------------------------------- pieces of uncompressed_header (spec 5.9.x)
struct Dimensions {
  uint32_t width, height;
};

void superres_params(BitReader &br, bool enable_superres, Dimensions &frame, Dimensions &upscaled) {
  const bool use_superres = enable_superres ? br.flag() : false;
  const uint32_t denom = use_superres ? (uint32_t)br.f(3) + 9 : 8;
  upscaled.width = frame.width;
  frame.width = (upscaled.width * 8 + denom / 2) / denom;
}

Dimensions frame_size(BitReader &br, bool override_flag, const SequenceHeader &s, uint32_t *upscaled_width) {
  Dimensions d;
  if (override_flag) {
    d.width = (uint32_t)br.f((unsigned)s.frame_width_bits_minus_1 + 1) + 1;
    d.height = (uint32_t)br.f((unsigned)s.frame_height_bits_minus_1 + 1) + 1;
  } else {
    d.width = s.max_frame_width_minus_1 + 1;
    d.height = s.max_frame_height_minus_1 + 1;
  }
  Dimensions up = d;
  superres_params(br, s.enable_superres, d, up);
  *upscaled_width = up.width;
  return d;
}

void render_size(BitReader &br) {
  if (br.flag()) {  // render_and_frame_size_different
    br.f(16);
    br.f(16);
  }
}

uint32_t tile_log2(uint32_t blk, uint32_t target) {
  uint32_t k = 0;
  while (((uint64_t)blk << k) < target) ++k;
  return k;
}

TileInfo tile_info(BitReader &br, bool use_128, uint32_t mi_cols, uint32_t mi_rows) {
  const uint32_t sb_cols = use_128 ? (mi_cols + 31) >> 5 : (mi_cols + 15) >> 4;
  const uint32_t sb_rows = use_128 ? (mi_rows + 31) >> 5 : (mi_rows + 15) >> 4;
  const uint32_t sb_size = (use_128 ? 5u : 4u) + 2;
  const uint32_t max_tile_width_sb = 4096u >> sb_size;
  const uint32_t max_tile_area_sb = (4096u * 2304u) >> (2 * sb_size);
  const uint32_t min_log2_tile_cols = tile_log2(max_tile_width_sb, sb_cols);
  const uint32_t max_log2_tile_cols = tile_log2(1, std::min<uint32_t>(sb_cols, 64));
  const uint32_t max_log2_tile_rows = tile_log2(1, std::min<uint32_t>(sb_rows, 64));
  const uint32_t min_log2_tiles = std::max(min_log2_tile_cols, tile_log2(max_tile_area_sb, sb_rows * sb_cols));
  TileInfo t;
  if (br.flag()) {  // uniform_tile_spacing_flag
    t.tile_cols_log2 = min_log2_tile_cols;
    while (t.tile_cols_log2 < max_log2_tile_cols && br.flag()) ++t.tile_cols_log2;
    const uint32_t tile_width_sb = (sb_cols + (1u << t.tile_cols_log2) - 1) >> t.tile_cols_log2;
    // spec 5.9.15 counts the tile starts, i.e. ceil(sbCols / tileWidthSb).  The reference floors (frame.rs:1106, :1121),
    // which loses the last, narrower tile (7 superblock columns at log2 = 1 are 2 tiles, not 1) and with it the
    // end-of-frame detection in tile_group_obu(); found by fuzzing resized libaom streams.
    t.tile_cols = (sb_cols + tile_width_sb - 1) / tile_width_sb;
    const uint32_t min_log2_tile_rows = min_log2_tiles > t.tile_cols_log2 ? min_log2_tiles - t.tile_cols_log2 : 0;
    t.tile_rows_log2 = min_log2_tile_rows;
    while (t.tile_rows_log2 < max_log2_tile_rows && br.flag()) ++t.tile_rows_log2;
    const uint32_t tile_height_sb = (sb_rows + (1u << t.tile_rows_log2) - 1) >> t.tile_rows_log2;
    t.tile_rows = (sb_rows + tile_height_sb - 1) / tile_height_sb;
  } else {
    uint32_t widest = 0, start = 0, n = 0;
    while (start < sb_cols) {
      const uint32_t size_sb = (uint32_t)br.ns(std::min(sb_cols - start, max_tile_width_sb)) + 1;
      widest = std::max(widest, size_sb);
      start += size_sb;
      ++n;
    }
    t.tile_cols = n;
    const uint32_t max_tile_height_sb = std::max<uint32_t>(max_tile_area_sb / widest, 1);
    start = 0;
    n = 0;
    while (start < sb_rows) {
      start += (uint32_t)br.ns(std::min(sb_rows - start, max_tile_height_sb)) + 1;
      ++n;
    }
    t.tile_rows = n;
    t.tile_cols_log2 = tile_log2(1, t.tile_cols);
    t.tile_rows_log2 = tile_log2(1, t.tile_rows);
  }
  if (t.tile_cols == 0 || t.tile_rows == 0) throw ParseError("tile_info: empty tile grid");
  if (t.tile_cols_log2 > 0 || t.tile_rows_log2 > 0) {
    br.f(t.tile_rows_log2 + t.tile_cols_log2);  // context_update_tile_id
    br.f(2);                                    // tile_size_bytes_minus_1
  }
  return t;
}

struct QuantParams {
  int base_q_idx = 0;
  int64_t y_dc = 0, u_dc = 0, u_ac = 0, v_dc = 0, v_ac = 0;
};

int64_t read_delta_q(BitReader &br) { return br.flag() ? br.su(7) : 0; }

QuantParams quantization_params(BitReader &br, int num_planes, bool separate_uv_delta_q) {
  QuantParams q;
  q.base_q_idx = (int)br.f(8);
  q.y_dc = read_delta_q(br);
  if (num_planes > 1) {
    const bool diff_uv_delta = separate_uv_delta_q ? br.flag() : false;
    q.u_dc = read_delta_q(br);
    q.u_ac = read_delta_q(br);
    if (diff_uv_delta) {
      q.v_dc = read_delta_q(br);
      q.v_ac = read_delta_q(br);
    } else {
      q.v_dc = q.u_dc;
      q.v_ac = q.u_ac;
    }
  }
  if (br.flag()) {  // using_qmatrix
    br.f(4);        // qm_y
    br.f(4);        // qm_u
    if (separate_uv_delta_q) br.f(4);  // qm_v
  }
  return q;
}

struct Segmentation {
  bool enabled = false;
  bool has[MAX_SEGMENTS][SEG_LVL_MAX] = {{false}};
  int16_t value[MAX_SEGMENTS][SEG_LVL_MAX] = {{0}};
};

// `previous` = the segmentation parameters saved with the primary reference frame (load_previous(), spec 7.20): they
// stay in force when segmentation is enabled without segmentation_update_data.  (The reference starts from an empty
// set there, frame.rs:1326-1395; the only use of the values is the lossless test, so the bits read are the same
// unless a lossless stream relies on inherited ALT_Q features.)
Segmentation segmentation_params(BitReader &br, int primary_ref_frame, const Segmentation *previous = nullptr) {
  static const int kBits[SEG_LVL_MAX] = {8, 6, 6, 6, 6, 3, 0, 0};
  static const bool kSigned[SEG_LVL_MAX] = {true, true, true, true, true, false, false, false};
  static const int kMax[SEG_LVL_MAX] = {255, 63, 63, 63, 63, 7, 0, 0};
  Segmentation sg;
  sg.enabled = br.flag();
  if (!sg.enabled) return sg;
  bool update_data = true;
  if (primary_ref_frame != PRIMARY_REF_NONE) {
    if (br.flag()) br.flag();  // segmentation_update_map, segmentation_temporal_update
    update_data = br.flag();
  }
  if (!update_data && previous) {
    std::memcpy(sg.has, previous->has, sizeof sg.has);
    std::memcpy(sg.value, previous->value, sizeof sg.value);
  }
  if (update_data) {
    for (int i = 0; i < MAX_SEGMENTS; ++i)
      for (int j = 0; j < SEG_LVL_MAX; ++j) {
        if (!br.flag()) continue;  // feature_enabled
        int64_t v;
        if (kSigned[j]) {
          v = br.su(1 + (unsigned)kBits[j]);
          v = std::max<int64_t>(-kMax[j], std::min<int64_t>(kMax[j], v));
        } else {
          v = (int64_t)br.f((unsigned)kBits[j]);
          v = std::max<int64_t>(0, std::min<int64_t>(kMax[j], v));
        }
        sg.has[i][j] = true;
        sg.value[i][j] = (int16_t)v;
      }
  }
  return sg;
}

bool delta_q_params(BitReader &br, int base_q_idx) {
  const bool present = base_q_idx > 0 ? br.flag() : false;
  if (present) br.f(2);  // delta_q_res
  return present;
}

void delta_lf_params(BitReader &br, bool delta_q_present, bool allow_intrabc) {
  if (!delta_q_present) return;
  const bool present = allow_intrabc ? false : br.flag();
  if (present) {
    br.f(2);   // delta_lf_res
    br.flag();  // delta_lf_multi
  }
}

int get_qindex_ignoring_delta(int segment_id, int base_q_idx, const Segmentation &sg) {
  if (sg.enabled && sg.has[segment_id][0]) {
    const int q = base_q_idx + sg.value[segment_id][0];
    return std::max(0, std::min(255, q));
  }
  return base_q_idx;
}

void loop_filter_params(BitReader &br, bool coded_lossless, bool allow_intrabc, int num_planes) {
  if (coded_lossless || allow_intrabc) return;
  const int l0 = (int)br.f(6), l1 = (int)br.f(6);
  if (num_planes > 1 && (l0 > 0 || l1 > 0)) {
    br.f(6);
    br.f(6);
  }
  br.f(3);          // loop_filter_sharpness
  if (br.flag()) {  // loop_filter_delta_enabled
    if (br.flag()) {  // loop_filter_delta_update
      for (int i = 0; i < 8; ++i)
        if (br.flag()) br.su(7);
      for (int i = 0; i < 2; ++i)
        if (br.flag()) br.su(7);
    }
  }
}

void cdef_params(BitReader &br, bool coded_lossless, bool allow_intrabc, bool enable_cdef, int num_planes) {
  if (coded_lossless || allow_intrabc || !enable_cdef) return;
  br.f(2);  // cdef_damping_minus_3
  const int cdef_bits = (int)br.f(2);
  for (int i = 0; i < (1 << cdef_bits); ++i) {
    br.f(4);
    br.f(2);
    if (num_planes > 1) {
      br.f(4);
      br.f(2);
    }
  }
}

void lr_params(BitReader &br, bool all_lossless, bool allow_intrabc, const SequenceHeader &s) {
  if (all_lossless || allow_intrabc || !s.enable_restoration) return;
  bool uses_lr = false, uses_chroma_lr = false;
  for (int i = 0; i < s.num_planes; ++i) {
    if (br.f(2) != 0) {
      uses_lr = true;
      if (i > 0) uses_chroma_lr = true;
    }
  }
  if (!uses_lr) return;
  if (s.use_128x128_superblock) {
    br.flag();  // lr_unit_shift
  } else if (br.flag()) {
    br.flag();  // lr_unit_extra_shift
  }
  if (s.ss_x && s.ss_y && uses_chroma_lr) br.flag();  // lr_uv_shift
}

int64_t get_relative_dist(int64_t a, int64_t b, int order_hint_bits) {
  if (order_hint_bits == 0) return 0;
  const int64_t diff = a - b;
  const int64_t m = (int64_t)1 << (order_hint_bits - 1);
  return (diff & (m - 1)) - (diff & m);
}

bool skip_mode_allowed(bool frame_is_intra, bool reference_select, int order_hint_bits, uint64_t order_hint,
                       const uint64_t *ref_order_hint, const int *ref_frame_idx) {
  if (frame_is_intra || !reference_select || order_hint_bits == 0) return false;
  int forward_idx = -1, backward_idx = -1;
  int64_t forward_hint = -1, backward_hint = -1;
  for (int i = 0; i < REFS_PER_FRAME; ++i) {
    const int64_t ref_hint = (int64_t)ref_order_hint[ref_frame_idx[i]];
    const int64_t d = get_relative_dist(ref_hint, (int64_t)order_hint, order_hint_bits);
    if (d < 0) {
      if (forward_idx < 0 || get_relative_dist(ref_hint, forward_hint, order_hint_bits) > 0) {
        forward_idx = i;
        forward_hint = ref_hint;
      }
    } else if (d > 0 && (backward_idx < 0 || get_relative_dist(ref_hint, backward_hint, order_hint_bits) < 0)) {
      backward_idx = i;
      backward_hint = ref_hint;
    }
  }
  if (forward_idx < 0) return false;
  if (backward_idx >= 0) return true;
  int second_forward_idx = -1;
  int64_t second_forward_hint = -1;
  for (int i = 0; i < REFS_PER_FRAME; ++i) {
    const int64_t ref_hint = (int64_t)ref_order_hint[ref_frame_idx[i]];
    if (get_relative_dist(ref_hint, forward_hint, order_hint_bits) < 0 &&
        (second_forward_idx < 0 || get_relative_dist(ref_hint, second_forward_hint, order_hint_bits) > 0)) {
      second_forward_idx = i;
      second_forward_hint = ref_hint;
    }
  }
  return second_forward_idx >= 0;
}

// decode_subexp (spec 5.9.28): only the number of bits consumed matters to this walk, the value is returned for tests
int32_t decode_subexp(BitReader &br, int32_t num_syms) {
  int32_t i = 0, mk = 0;
  const int32_t k = 3;
  for (;;) {
    const int32_t b2 = i ? k + i - 1 : k;
    const int32_t a = 1 << b2;
    if (num_syms <= mk + 3 * a) return (int32_t)br.ns((uint64_t)(num_syms - mk)) + mk;
    if (br.flag()) {  // subexp_more_bits
      ++i;
      mk += a;
    } else {
      return (int32_t)br.f((unsigned)b2) + mk;
    }
  }
}

int32_t inverse_recenter(int32_t r, int32_t v) {
  if (v > 2 * r) return v;
  if (v & 1) return r - ((v + 1) >> 1);
  return r + (v >> 1);
}

int32_t decode_unsigned_subexp_with_ref(BitReader &br, int32_t mx, int32_t r) {
  const int32_t v = decode_subexp(br, mx);
  if ((r << 1) <= mx) return inverse_recenter(r, v);
  return mx - 1 - inverse_recenter(mx - 1 - r, v);
}

int32_t decode_signed_subexp_with_ref(BitReader &br, int32_t low, int32_t high, int32_t r) {
  return decode_unsigned_subexp_with_ref(br, high - low, r - low) + low;
}

void read_global_param(BitReader &br, bool allow_high_precision_mv, int type, int idx) {
  int abs_bits = 12, prec_bits = 15;  // GM_ABS_ALPHA_BITS, GM_ALPHA_PREC_BITS
  if (idx < 2) {
    if (type == 1) {  // TRANSLATION
      abs_bits = 9 - (allow_high_precision_mv ? 0 : 1);
      prec_bits = 3 - (allow_high_precision_mv ? 0 : 1);
    } else {
      abs_bits = 12;
      prec_bits = 6;
    }
  }
  const int prec_diff = 16 - prec_bits;
  const int32_t sub = (idx % 3 == 2) ? (1 << prec_bits) : 0;
  const int32_t prev = (idx % 3 == 2) ? (1 << 16) : 0;  // the reference always starts from the identity model
  const int32_t mx = 1 << abs_bits;
  const int32_t r = (prev >> prec_diff) - sub;
  decode_signed_subexp_with_ref(br, -mx, mx + 1, r);
}

void global_motion_params(BitReader &br, bool frame_is_intra, bool allow_high_precision_mv) {
  if (frame_is_intra) return;
  for (int ref = 1; ref <= 7; ++ref) {
    int type = 0;  // IDENTITY
    if (br.flag()) {  // is_global
      if (br.flag())
        type = 2;  // ROTZOOM
      else
        type = br.flag() ? 1 : 3;  // TRANSLATION : AFFINE
    }
    if (type >= 2) {
      read_global_param(br, allow_high_precision_mv, type, 2);
      read_global_param(br, allow_high_precision_mv, type, 3);
      if (type == 3) {
        read_global_param(br, allow_high_precision_mv, type, 4);
        read_global_param(br, allow_high_precision_mv, type, 5);
      }
    }
    if (type >= 1) {
      read_global_param(br, allow_high_precision_mv, type, 0);
      read_global_param(br, allow_high_precision_mv, type, 1);
    }
  }
}

}  // namespace

// ---------------------------------------------------------------- the parser object behind the C ABI
struct g1s_inspect {
  std::string err;
  SequenceHeader seq;
  bool seen_frame_header = false;
  bool have_frame_header = false;
  TileInfo cur_tile_info;  // layout of the frame whose tile groups are being walked
  int ref_frame_idx[REFS_PER_FRAME] = {0};
  uint64_t ref_order_hint[NUM_REF_FRAMES] = {0};
  uint64_t big_ref_order_hint[NUM_REF_FRAMES] = {0};
  bool big_ref_valid[NUM_REF_FRAMES] = {false};
  // sizes saved with every reference slot (spec 7.20): found_ref takes the frame size from the referenced slot
  uint32_t ref_upscaled_width[NUM_REF_FRAMES] = {0}, ref_frame_height[NUM_REF_FRAMES] = {0};
  int ref_frame_type[NUM_REF_FRAMES] = {0};
  Segmentation ref_segmentation[NUM_REF_FRAMES];
  std::vector<GrainHeader> headers;  // one per shown frame header, in stream order (parser.rs:155-158)
  uint64_t packets = 0, obus = 0;
  uint32_t last_frame_width = 0, last_frame_height = 0;
  // rewriter (BitstreamParser::<true>, parser.rs:74-101): `write` mirrors every OBU into packet_out; `have_table`
  // is incoming_grain_header.is_some() (apply) vs None (remove)
  bool write = false, have_table = false;
  std::vector<g1s_segment> table;
  std::vector<uint8_t> packet_out;
  uint64_t packet_ts = 0;
  uint64_t frames_with_grain = 0, frames_grain_disabled = 0;
  // the rewritten header of the current frame as a standalone frame_header_obu() payload: repeats of the header inside
  // the frame (OBU_FRAME_HEADER again, OBU_REDUNDANT_FRAME_HEADER) must stay bit-identical to it (spec 7.5)
  std::vector<uint8_t> current_header_payload;
  std::vector<uint8_t> rewrite_frame_payload(const uint8_t *payload, size_t obu_size, const FrameHeader &fh,
                                             size_t header_end_bits, bool is_frame_obu);

  FrameHeader uncompressed_header(BitReader &br, bool has_ext, int temporal_id, int spatial_id, bool verify_alignment);
  // returns true and fills `out` when the header belongs to a shown frame
  bool parse_frame_header(BitReader &br, bool has_ext, int tid, int sid, bool verify_alignment, FrameHeader *out);
  void tile_group_header(BitReader &br, const TileInfo &ti);
  void set_frame_refs(int last_frame_idx, int gold_frame_idx, uint64_t order_hint, int order_hint_bits);
  void parse_packet(const uint8_t *data, size_t size);
};

// AV1 spec 7.8, "set frame refs process": with frame_refs_short_signaling only LAST and GOLDEN are coded, the other
// five references follow from the order hints of the eight slots.
void g1s_inspect::set_frame_refs(int last_frame_idx, int gold_frame_idx, uint64_t order_hint, int order_hint_bits) {
  enum { LAST = 0, LAST2 = 1, LAST3 = 2, GOLDEN = 3, BWDREF = 4, ALTREF2 = 5, ALTREF = 6 };
  bool used[NUM_REF_FRAMES] = {false};
  int64_t shifted[NUM_REF_FRAMES];
  for (int i = 0; i < REFS_PER_FRAME; ++i) ref_frame_idx[i] = -1;
  ref_frame_idx[LAST] = last_frame_idx;
  ref_frame_idx[GOLDEN] = gold_frame_idx;
  used[last_frame_idx] = used[gold_frame_idx] = true;
  const int64_t cur = (int64_t)1 << (order_hint_bits - 1);
  for (int i = 0; i < NUM_REF_FRAMES; ++i)
    shifted[i] = cur + get_relative_dist((int64_t)big_ref_order_hint[i], (int64_t)order_hint, order_hint_bits);
  auto find = [&](bool backward, bool latest) {
    int ref = -1;
    int64_t best = 0;
    for (int i = 0; i < NUM_REF_FRAMES; ++i) {
      const int64_t hint = shifted[i];
      if (used[i] || (backward ? hint < cur : hint >= cur)) continue;
      if (ref < 0 || (latest ? hint >= best : hint < best)) {
        ref = i;
        best = hint;
      }
    }
    return ref;
  };
  auto take = [&](int slot, int ref) {
    if (ref >= 0) {
      ref_frame_idx[slot] = ref;
      used[ref] = true;
    }
  };
  take(ALTREF, find(true, true));     // the backward reference furthest in the future
  take(BWDREF, find(true, false));    // the closest backward reference
  take(ALTREF2, find(true, false));   // the next closest
  static const int kOrder[REFS_PER_FRAME - 2] = {LAST2, LAST3, BWDREF, ALTREF2, ALTREF};
  for (int slot : kOrder)
    if (ref_frame_idx[slot] < 0) take(slot, find(false, true));  // remaining: forward references, latest first
  int ref = -1;
  int64_t earliest = 0;
  for (int i = 0; i < NUM_REF_FRAMES; ++i)
    if (ref < 0 || shifted[i] < earliest) {
      ref = i;
      earliest = shifted[i];
    }
  for (int i = 0; i < REFS_PER_FRAME; ++i)
    if (ref_frame_idx[i] < 0) ref_frame_idx[i] = ref;
}

FrameHeader g1s_inspect::uncompressed_header(BitReader &br, bool has_ext, int temporal_id, int spatial_id,
                                             bool verify_alignment) {
  if (!seq.valid) throw ParseError("frame header before any sequence header");
  const SequenceHeader &s = seq;
  const int id_len = s.frame_id_numbers_present ? s.additional_frame_id_len_minus_1 + s.delta_frame_id_len_minus_2 + 3 : 0;
  FrameHeader fh;
  int frame_type = KEY_FRAME;
  bool show_frame = true, showable_frame = true, error_resilient_mode = false;
  if (!s.reduced_still_picture_header) {
    if (br.flag()) {  // show_existing_frame
      const int shown_slot = (int)br.f(3);  // frame_to_show_map_idx
      // spec 5.9.2: temporal_point_info() precedes display_frame_id here too (the reference skips it, and then fails
      // its alignment check on such streams)
      if (s.has_decoder_model && !(s.has_timing_info && s.equal_picture_interval))
        br.f((unsigned)s.frame_presentation_time_length_minus_1 + 1);
      if (id_len) br.f((unsigned)id_len);
      if (!have_frame_header) throw ParseError("show_existing_frame before any frame header");
      if (ref_frame_type[shown_slot] == KEY_FRAME && big_ref_valid[shown_slot]) {
        // spec 7.21: showing a key frame refreshes every slot with it (the reference skips this; later frames' skip-mode
        // decision reads these order hints)
        for (int i = 0; i < NUM_REF_FRAMES; ++i) {
          big_ref_order_hint[i] = big_ref_order_hint[shown_slot];
          ref_upscaled_width[i] = ref_upscaled_width[shown_slot];
          ref_frame_height[i] = ref_frame_height[shown_slot];
          ref_frame_type[i] = KEY_FRAME;
          ref_segmentation[i] = ref_segmentation[shown_slot];
          big_ref_valid[i] = true;
        }
      }
      if (verify_alignment) br.byte_alignment(true);
      fh.show_frame = true;
      fh.show_existing_frame = true;
      fh.grain.kind = G1S_GRAIN_COPY_REF_FRAME;
      fh.tile_info = cur_tile_info;
      return fh;
    }
    frame_type = (int)br.f(2);
    show_frame = br.flag();
    if (show_frame && s.has_decoder_model && !(s.has_timing_info && s.equal_picture_interval))
      br.f((unsigned)s.frame_presentation_time_length_minus_1 + 1);  // temporal_point_info
    showable_frame = show_frame ? frame_type != KEY_FRAME : br.flag();
    error_resilient_mode = (frame_type == SWITCH_FRAME || (frame_type == KEY_FRAME && show_frame)) ? true : br.flag();
  }
  const bool frame_is_intra = frame_type == KEY_FRAME || frame_type == INTRA_ONLY_FRAME;
  if (frame_type == KEY_FRAME && show_frame) {
    for (int i = 0; i < NUM_REF_FRAMES; ++i) {
      big_ref_valid[i] = false;
      big_ref_order_hint[i] = 0;
    }
  }
  const bool disable_cdf_update = br.flag();
  const bool allow_screen_content_tools =
      s.force_screen_content_tools == SELECT_SCREEN_CONTENT_TOOLS ? br.flag() : s.force_screen_content_tools == 1;
  // spec 5.9.2: the FRAME-level force_integer_mv (coded only under SELECT_INTEGER_MV) gates allow_high_precision_mv
  // below; the reference reads the bit and then tests the sequence-level value (frame.rs:469), which shifts every
  // later field of such an inter frame by one bit
  int force_integer_mv = 0;
  if (allow_screen_content_tools)
    force_integer_mv = s.force_integer_mv == SELECT_INTEGER_MV ? (int)br.flag() : s.force_integer_mv;
  if (frame_is_intra) force_integer_mv = 1;
  if (s.frame_id_numbers_present) br.f((unsigned)id_len);                               // current_frame_id
  const bool frame_size_override_flag =
      frame_type == SWITCH_FRAME ? true : (s.reduced_still_picture_header ? false : br.flag());
  const uint64_t order_hint = br.f((unsigned)s.order_hint_bits);
  const int primary_ref_frame = (frame_is_intra || error_resilient_mode) ? PRIMARY_REF_NONE : (int)br.f(3);
  if (s.has_decoder_model) {
    if (br.flag()) {  // buffer_removal_time_present_flag
      for (int op = 0; op <= s.operating_points_cnt_minus_1; ++op) {
        if (!s.decoder_model_present_for_op[op]) continue;
        const uint16_t idc = s.operating_point_idc[op];
        const int tid = has_ext ? temporal_id : 0, sid = has_ext ? spatial_id : 0;
        const bool in_t = (idc >> tid) & 1, in_s = (idc >> (sid + 8)) & 1;
        if (idc == 0 || (in_t && in_s)) br.f((unsigned)s.buffer_removal_time_length_minus_1 + 1);
      }
    }
  }
  bool allow_intrabc = false;
  const unsigned refresh_frame_flags =
      (frame_type == SWITCH_FRAME || (frame_type == KEY_FRAME && show_frame)) ? 0xFFu : (unsigned)br.f(8);
  if ((!frame_is_intra || refresh_frame_flags != 0xFFu) && error_resilient_mode && s.order_hint_bits > 0) {
    for (int i = 0; i < NUM_REF_FRAMES; ++i) {
      // spec 5.9.2: the signalled hint REPLACES the slot's saved hint (and invalidates the slot when it differs).  The
      // reference saves the previous signalled value instead (frame.rs:355-362), which zeroes the saved hints the
      // first time an error-resilient inter frame arrives and then mis-decides skip_mode_present.
      const uint64_t cur = br.f((unsigned)s.order_hint_bits);  // ref_order_hint[i]
      if (cur != big_ref_order_hint[i]) big_ref_valid[i] = false;
      big_ref_order_hint[i] = cur;
      ref_order_hint[i] = cur;
    }
  }
  bool allow_high_precision_mv = false, use_ref_frame_mvs = false;
  Dimensions fsize, upscaled;
  uint32_t true_upscaled_width = 0;  // UpscaledWidth of the spec, saved with the reference slots
  if (frame_is_intra) {
    fsize = frame_size(br, frame_size_override_flag, s, &true_upscaled_width);
    upscaled = fsize;
    render_size(br);
    // spec 5.9.2: allow_intrabc is coded only when UpscaledWidth == FrameWidth.  (The reference compares the
    // post-superres width with itself, frame.rs:389-399, and would read a bit that a super-resolved frame lacks.)
    if (allow_screen_content_tools && true_upscaled_width == fsize.width) allow_intrabc = br.flag();
  } else {
    bool frame_refs_short_signaling = false;
    if (s.order_hint_bits > 0) {
      frame_refs_short_signaling = br.flag();
      if (frame_refs_short_signaling) {
        const int last_frame_idx = (int)br.f(3);
        const int gold_frame_idx = (int)br.f(3);
        // The reference stubs set_frame_refs() out and leaves every index at 0 (frame.rs:415-437, :941); the indices
        // decide whether skip_mode_present is coded, so the spec's process (7.8) is run here instead.
        set_frame_refs(last_frame_idx, gold_frame_idx, order_hint, s.order_hint_bits);
      }
    }
    for (int i = 0; i < REFS_PER_FRAME; ++i) {
      if (!frame_refs_short_signaling) ref_frame_idx[i] = (int)br.f(3);  // else derived above
      // spec 5.9.2: delta_frame_id_minus_1 is coded for every reference whenever frame ids are present, short
      // signaling or not
      if (s.frame_id_numbers_present) br.f((unsigned)s.delta_frame_id_len_minus_2 + 2);
    }
    if (frame_size_override_flag && !error_resilient_mode) {
      int found_ref = -1;
      for (int i = 0; i < REFS_PER_FRAME && found_ref < 0; ++i)
        if (br.flag()) found_ref = i;
      if (found_ref >= 0) {
        // The reference keeps the sequence maximum here (frame.rs:456-470 with :949-975), which reads the same bits
        // unless the tile-column / tile-row limits differ between the two sizes; the slot's own size is used instead.
        const int slot = ref_frame_idx[found_ref];
        fsize = Dimensions{ref_upscaled_width[slot], ref_frame_height[slot]};
        if (fsize.width == 0 || fsize.height == 0)
          fsize = Dimensions{s.max_frame_width_minus_1 + 1, s.max_frame_height_minus_1 + 1};
        upscaled = fsize;
        superres_params(br, s.enable_superres, fsize, upscaled);
        true_upscaled_width = upscaled.width;
      } else {
        fsize = frame_size(br, frame_size_override_flag, s, &true_upscaled_width);
        upscaled = Dimensions{s.max_frame_width_minus_1 + 1, s.max_frame_height_minus_1 + 1};
        render_size(br);
      }
    } else {
      fsize = frame_size(br, frame_size_override_flag, s, &true_upscaled_width);
      upscaled = fsize;
      render_size(br);
    }
    allow_high_precision_mv = force_integer_mv ? false : br.flag();
    if (!br.flag()) br.f(2);  // is_filter_switchable, interpolation_filter
    br.flag();                // is_motion_mode_switchable
    use_ref_frame_mvs = (error_resilient_mode || !s.enable_ref_frame_mvs) ? false : br.flag();
  }
  (void)use_ref_frame_mvs;
  const uint32_t upscaled_width_for_refs = true_upscaled_width;
  last_frame_width = fsize.width;
  last_frame_height = fsize.height;
  const uint32_t mi_cols = 2 * ((fsize.width + 7) >> 3), mi_rows = 2 * ((fsize.height + 7) >> 3);
  if (!(s.reduced_still_picture_header || disable_cdf_update)) br.flag();  // disable_frame_end_update_cdf
  fh.tile_info = tile_info(br, s.use_128x128_superblock, mi_cols, mi_rows);
  const QuantParams q = quantization_params(br, s.num_planes, s.separate_uv_delta_q);
  const Segmentation sg = segmentation_params(
      br, primary_ref_frame, primary_ref_frame == PRIMARY_REF_NONE ? nullptr : &ref_segmentation[ref_frame_idx[primary_ref_frame]]);
  const bool delta_q_present = delta_q_params(br, q.base_q_idx);
  delta_lf_params(br, delta_q_present, allow_intrabc);
  bool coded_lossless = true;
  for (int seg = 0; seg < MAX_SEGMENTS && coded_lossless; ++seg) {
    const int qindex = get_qindex_ignoring_delta(seg, q.base_q_idx, sg);
    coded_lossless = qindex == 0 && q.y_dc == 0 && q.u_ac == 0 && q.u_dc == 0 && q.v_ac == 0 && q.v_dc == 0;
  }
  (void)upscaled;
  const bool all_lossless = coded_lossless && fsize.width == true_upscaled_width;  // spec: FrameWidth == UpscaledWidth
  loop_filter_params(br, coded_lossless, allow_intrabc, s.num_planes);
  cdef_params(br, coded_lossless, allow_intrabc, s.enable_cdef, s.num_planes);
  lr_params(br, all_lossless, allow_intrabc, s);
  if (!coded_lossless) br.flag();  // tx_mode_select
  const bool reference_select = frame_is_intra ? false : br.flag();
  if (skip_mode_allowed(frame_is_intra, reference_select, s.order_hint_bits, order_hint, big_ref_order_hint,
                        ref_frame_idx))
    br.flag();  // skip_mode_present
  if (!(frame_is_intra || error_resilient_mode || !s.enable_warped_motion)) br.flag();  // allow_warped_motion
  br.flag();                                                                           // reduced_tx_set
  global_motion_params(br, frame_is_intra, allow_high_precision_mv);
  const bool film_grain_allowed = show_frame || showable_frame;
  fh.grain_pos = br.pos;
  fh.film_grain_allowed = film_grain_allowed;
  fh.frame_type = frame_type;
  fh.grain = film_grain_params(br, s.film_grain_params_present && film_grain_allowed, frame_type, s.num_planes == 1,
                               s.ss_x, s.ss_y);
  for (int i = 0; i < NUM_REF_FRAMES; ++i) {
    if ((refresh_frame_flags >> i) & 1) {
      big_ref_valid[i] = true;
      big_ref_order_hint[i] = order_hint;
      ref_upscaled_width[i] = upscaled_width_for_refs;
      ref_frame_height[i] = fsize.height;
      ref_frame_type[i] = frame_type;
      ref_segmentation[i] = sg;
    }
  }
  if (verify_alignment) br.byte_alignment(true);
  fh.show_frame = show_frame;
  fh.show_existing_frame = false;
  return fh;
}

bool g1s_inspect::parse_frame_header(BitReader &br, bool has_ext, int tid, int sid, bool verify_alignment,
                                     FrameHeader *out) {
  if (seen_frame_header) return false;  // a repeat inside the same frame: nothing is read (frame.rs:117-119)
  seen_frame_header = true;
  const FrameHeader fh = uncompressed_header(br, has_ext, tid, sid, verify_alignment);
  if (fh.show_existing_frame) seen_frame_header = false;
  *out = fh;  // hidden frames are returned too (their tile layout is needed), but only shown ones count
  return fh.show_frame;
}

void g1s_inspect::tile_group_header(BitReader &br, const TileInfo &ti) {
  const uint32_t num_tiles = ti.tile_cols * ti.tile_rows;
  const bool start_and_end_present = num_tiles > 1 ? br.flag() : false;
  uint32_t tg_end = num_tiles - 1;
  if (!(num_tiles == 1 || !start_and_end_present)) {
    const unsigned tile_bits = ti.tile_cols_log2 + ti.tile_rows_log2;
    br.f(tile_bits);  // tg_start
    tg_end = (uint32_t)br.f(tile_bits);
  }
  br.byte_alignment(true);
  if (tg_end == num_tiles - 1) seen_frame_header = false;
}

// New payload of an OBU_FRAME / OBU_FRAME_HEADER whose header was just parsed (frame.rs:608-676): the bits before
// film_grain_params() are kept, the grain syntax is replaced (or dropped), the rest is re-aligned.
std::vector<uint8_t> g1s_inspect::rewrite_frame_payload(const uint8_t *payload, size_t obu_size, const FrameHeader &fh,
                                                        size_t header_end_bits, bool is_frame_obu) {
  BitWriter bw;
  bw.copy_bits(payload, fh.grain_pos);
  if (have_table && fh.film_grain_allowed) {  // sequence_header.new_film_grain_state && film_grain_allowed
    g1s_segment *seg = nullptr;
    for (g1s_segment &t : table)
      if (t.start_time <= packet_ts && packet_ts < t.end_time) {
        seg = &t;
        break;
      }
    if (seg) {
      seg->random_seed = (uint16_t)(seg->random_seed + DEFAULT_GRAIN_SEED);  // wrapping_add, kept for the next frame
      write_film_grain_bits(bw, *seg, fh.frame_type, seq.num_planes == 1, seq.ss_x, seq.ss_y);
      ++frames_with_grain;
    } else {
      bw.put(0, 1);  // apply_grain = 0
      ++frames_grain_disabled;
    }
  }
  std::vector<uint8_t> out;
  if (is_frame_obu) {
    out = bw.bytes;  // byte_alignment(): the partial byte is already zero-padded
    const size_t tile_off = (header_end_bits + 7) >> 3;
    out.insert(out.end(), payload + tile_off, payload + obu_size);
  }
  bw.put(1, 1);  // trailing_bits()
  current_header_payload = bw.bytes;
  if (!is_frame_obu) out = current_header_payload;
  return out;
}

void g1s_inspect::parse_packet(const uint8_t *data, size_t size) {
  ++packets;
  packet_out.clear();
  size_t off = 0;
  while (off < size) {
    ++obus;
    const uint8_t b0 = data[off];
    if (b0 & 0x80) throw ParseError("obu_forbidden_bit is set");
    if (b0 & 0x01) throw ParseError("obu_reserved_1bit is set");
    const int type = (b0 >> 3) & 0xF;
    const bool has_ext = (b0 >> 2) & 1, has_size = (b0 >> 1) & 1;
    size_t hdr = 1;
    int tid = 0, sid = 0;
    if (has_ext) {
      if (off + 1 >= size) throw ParseError("OBU extension byte missing");
      tid = data[off + 1] >> 5;
      sid = (data[off + 1] >> 3) & 3;
      hdr = 2;
    }
    size_t obu_size;
    if (has_size) {
      uint64_t v;
      size_t used;
      if (!read_leb128(data + off + hdr, size - off - hdr, &v, &used)) throw ParseError("truncated obu_size");
      hdr += used;
      obu_size = (size_t)v;
    } else {
      obu_size = size - off - hdr;  // the OBU runs to the end of the packet
    }
    if (off + hdr + obu_size > size) throw ParseError("OBU larger than its packet");
    const uint8_t *payload = data + off + hdr;
    const size_t obu_hdr_bytes = has_ext ? 2 : 1;
    const uint8_t *obu_start = data + off;
    off += hdr + obu_size;
    std::vector<uint8_t> new_payload;
    bool replaced = false;
    // write mode: the OBU goes out again when this scope ends, with `new_payload` instead of its payload if replaced
    struct Emit {
      g1s_inspect *self;
      const uint8_t *obu_start, *payload;
      size_t obu_hdr_bytes, obu_size;
      bool has_size;
      std::vector<uint8_t> *np;
      bool *replaced;
      ~Emit() {
        if (!self->write) return;
        std::vector<uint8_t> &o = self->packet_out;
        o.insert(o.end(), obu_start, obu_start + obu_hdr_bytes);
        const uint8_t *body = *replaced ? np->data() : payload;
        const size_t n = *replaced ? np->size() : obu_size;
        if (has_size) {
          size_t v = n;
          do {
            uint8_t b = v & 0x7f;
            v >>= 7;
            o.push_back((uint8_t)(b | (v ? 0x80 : 0)));
          } while (v);
        }
        o.insert(o.end(), body, body + n);
      }
    } emit{this, obu_start, payload, obu_hdr_bytes, obu_size, has_size, &new_payload, &replaced};

    // operating point 0 only (obu.rs:92-116)
    if (type != OBU_SEQUENCE_HEADER && type != OBU_TEMPORAL_DELIMITER && has_ext && seq.valid) {
      const uint16_t idc = seq.cur_operating_point_idc;
      if (idc != 0 && !(((idc >> tid) & 1) && ((idc >> (sid + 8)) & 1))) continue;
    }
    switch (type) {
      case OBU_SEQUENCE_HEADER: {
        BitReader br(payload, obu_size);
        seq = parse_sequence_header(br);
        if (write) {  // film_grain_params_present follows what is being written (sequence.rs:404-423)
          new_payload.assign(payload, payload + obu_size);
          const size_t bit = br.pos - 1;
          if (have_table)
            new_payload[bit >> 3] |= (uint8_t)(0x80u >> (bit & 7));
          else
            new_payload[bit >> 3] &= (uint8_t)~(0x80u >> (bit & 7));
          replaced = true;
        }
        break;
      }
      case OBU_TEMPORAL_DELIMITER:
        seen_frame_header = false;
        break;
      case OBU_FRAME: {
        BitReader br(payload, obu_size);
        FrameHeader fh;
        const bool was_new = !seen_frame_header;
        const bool shown = parse_frame_header(br, has_ext, tid, sid, true, &fh);
        if (was_new) {
          if (shown) headers.push_back(fh.grain);
          // The tile group that follows is walked with THIS frame's tile layout.  (The reference uses the last SHOWN
          // header's layout for a hidden frame, frame.rs:86-89 with parser.rs:157; identical unless the layout changes
          // between a hidden frame and the shown frame before it.)
          if (!fh.show_existing_frame) cur_tile_info = fh.tile_info;
          have_frame_header = true;
          if (write && !fh.show_existing_frame) {
            new_payload = rewrite_frame_payload(payload, obu_size, fh, br.pos, true);
            replaced = true;
          }
        } else if (!have_frame_header) {
          throw ParseError("tile data before any frame header");
        }
        if (!fh.show_existing_frame) tile_group_header(br, cur_tile_info);
        break;
      }
      case OBU_FRAME_HEADER: {
        BitReader br(payload, obu_size);
        FrameHeader fh;
        const bool was_new = !seen_frame_header;
        const bool shown = parse_frame_header(br, has_ext, tid, sid, false, &fh);
        if (was_new) {
          if (shown) headers.push_back(fh.grain);
          if (!fh.show_existing_frame) cur_tile_info = fh.tile_info;
          have_frame_header = true;
          if (write && !fh.show_existing_frame) {
            new_payload = rewrite_frame_payload(payload, obu_size, fh, br.pos, false);
            replaced = true;
          }
          if (fh.show_existing_frame) current_header_payload.clear();
        } else if (write && !current_header_payload.empty()) {
          new_payload = current_header_payload;  // a repeat of the current frame's header
          replaced = true;
        }
        break;
      }
      case 7: {  // OBU_REDUNDANT_FRAME_HEADER: parsed by nobody (obu.rs:236-249), kept equal to the rewritten header
        if (write && seen_frame_header && !current_header_payload.empty()) {
          new_payload = current_header_payload;
          replaced = true;
        }
        break;
      }
      case OBU_TILE_GROUP: {
        if (!have_frame_header) throw ParseError("tile group before any frame header");
        BitReader br(payload, obu_size);
        tile_group_header(br, cur_tile_info);
        break;
      }
      default:
        break;  // metadata, padding, tile list, redundant frame header, reserved: skipped
    }
  }
}

// ---------------------------------------------------------------- aggregate_grain_headers (main.rs:713-772)
static std::vector<g1s_segment> aggregate(const std::vector<GrainHeader> &hs, int64_t fps_num, int64_t fps_den) {
  const double time_per_packet = (double)fps_den / (double)fps_num * 10000000.0;
  uint64_t start = 0;
  double end_f = time_per_packet;
  uint64_t end = (uint64_t)std::ceil(end_f);
  std::vector<g1s_segment> acc;
  for (const GrainHeader &h : hs) {
    const bool prev_has_grain = !acc.empty() && acc.back().end_time == start;
    if (prev_has_grain) {
      if (h.kind == G1S_GRAIN_COPY_REF_FRAME) {
        acc.back().end_time = end;
      } else if (h.kind == G1S_GRAIN_UPDATE) {
        if (same_grain(h.params, acc.back())) {
          acc.back().end_time = end;
        } else {
          g1s_segment s = h.params;
          s.start_time = start;
          s.end_time = end;
          acc.push_back(s);
        }
      }  // Disable: nothing, the run ends here
    } else if (h.kind == G1S_GRAIN_UPDATE) {
      g1s_segment s = h.params;
      s.start_time = start;
      s.end_time = end;
      acc.push_back(s);
    }
    start = end;
    end_f += time_per_packet;
    end = (uint64_t)std::ceil(end_f);
  }
  return acc;
}

// ---------------------------------------------------------------- C ABI
extern "C" {

int g1s_inspect_create(g1s_inspect **out) {
  if (!out) return G1S_E_ARG;
  *out = new (std::nothrow) g1s_inspect();
  return *out ? G1S_OK : G1S_E_NOMEM;
}

void g1s_inspect_destroy(g1s_inspect *h) { delete h; }

const char *g1s_inspect_last_error(const g1s_inspect *h) { return h ? h->err.c_str() : "null handle"; }

int g1s_inspect_push_packet(g1s_inspect *h, const uint8_t *data, size_t size) {
  if (!h || (!data && size)) return G1S_E_ARG;
  try {
    h->parse_packet(data, size);
  } catch (const std::exception &e) {
    char where[64];
    std::snprintf(where, sizeof where, " (packet %llu)", (unsigned long long)h->packets - 1);
    h->err = std::string(e.what()) + where;
    return G1S_E_STREAM;
  }
  return G1S_OK;
}

size_t g1s_inspect_num_headers(const g1s_inspect *h) { return h ? h->headers.size() : 0; }

int g1s_inspect_header(const g1s_inspect *h, size_t i, int32_t *kind, g1s_segment *params) {
  if (!h || i >= h->headers.size()) return G1S_E_ARG;
  if (kind) *kind = h->headers[i].kind;
  if (params) *params = h->headers[i].params;
  return G1S_OK;
}

int g1s_inspect_finish(g1s_inspect *h, int64_t fps_num, int64_t fps_den, g1s_segment *out, size_t cap, size_t *n) {
  if (!h || !n || fps_num <= 0 || fps_den <= 0) return G1S_E_ARG;
  const std::vector<g1s_segment> segs = aggregate(h->headers, fps_num, fps_den);
  *n = segs.size();
  if (segs.size() > cap || (!out && !segs.empty())) return G1S_E_STATE;
  if (!segs.empty()) std::memcpy(out, segs.data(), segs.size() * sizeof(g1s_segment));
  return G1S_OK;
}

// ---- apply / remove: BitstreamParser::<true>::modify_grain_headers (parser.rs:175-348) without the FFmpeg muxer
int g1s_rewrite_create(const g1s_segment *table, size_t n, int apply, g1s_inspect **out) {
  if (!out || (apply && (!table || n == 0))) return G1S_E_ARG;
  g1s_inspect *h = new (std::nothrow) g1s_inspect();
  if (!h) return G1S_E_NOMEM;
  h->write = true;
  h->have_table = apply != 0;
  if (apply) h->table.assign(table, table + n);
  *out = h;
  return G1S_OK;
}

int g1s_rewrite_packet(g1s_inspect *h, const uint8_t *data, size_t size, uint64_t packet_ts, size_t *out_size) {
  if (!h || !h->write || (!data && size) || !out_size) return G1S_E_ARG;
  h->packet_ts = packet_ts;
  const int rc = g1s_inspect_push_packet(h, data, size);
  *out_size = rc == G1S_OK ? h->packet_out.size() : 0;
  return rc;
}

int g1s_rewrite_take(g1s_inspect *h, uint8_t *out, size_t cap) {
  if (!h || !h->write || (!out && !h->packet_out.empty())) return G1S_E_ARG;
  if (cap < h->packet_out.size()) return G1S_E_STATE;
  if (!h->packet_out.empty()) std::memcpy(out, h->packet_out.data(), h->packet_out.size());
  return G1S_OK;
}

int g1s_rewrite_counters(const g1s_inspect *h, uint64_t *with_grain, uint64_t *disabled) {
  if (!h) return G1S_E_ARG;
  if (with_grain) *with_grain = h->frames_with_grain;
  if (disabled) *disabled = h->frames_grain_disabled;
  return G1S_OK;
}

// ---- generate: av1_grain::generate_photon_noise_params as called at src/main.rs:288-303 (crate source absent; this is
// the algorithm of libaom examples/photon_noise_table.c, which the crate ports, in f32 like both).  The BT.470BG branch
// exists because the reference's only fixture, tests/example-table.tbl, is an output of that generator with gamma 2.8
// (e.g. 1920x1080 at ISO 750): tests reproduce it byte for byte.  The BT.1886 and PQ curves are the crate's two
// transfer functions, restated from its published formulae (unpinned); its `full_range` argument is not modelled.
namespace {
struct Transfer {
  int id;
  float to_linear(float x) const {
    if (id == G1S_TRANSFER_SMPTE2084) {
      const float m1 = 2610.f / 16384.f, m2 = 128.f * 2523.f / 4096.f, c1 = 3424.f / 4096.f, c2 = 32.f * 2413.f / 4096.f,
                  c3 = 32.f * 2392.f / 4096.f;
      const float p = std::pow(x, 1.f / m2);
      return std::pow(std::fmax(0.f, p - c1) / (c2 - c3 * p), 1.f / m1);
    }
    if (id == G1S_TRANSFER_BT470BG) return std::pow(x, 2.8f);
    return bt1886_alpha() * std::pow(std::fmax(0.f, x + bt1886_beta()), 2.4f) / 203.f;
  }
  float from_linear(float x) const {
    if (id == G1S_TRANSFER_SMPTE2084) {
      const float m1 = 2610.f / 16384.f, m2 = 128.f * 2523.f / 4096.f, c1 = 3424.f / 4096.f, c2 = 32.f * 2413.f / 4096.f,
                  c3 = 32.f * 2392.f / 4096.f;
      if (x < 1.1920929e-7f) return 0.f;
      const float p = std::pow(x, m1);
      return std::pow((c1 + c2 * p) / (1.f + c3 * p), m2);
    }
    if (id == G1S_TRANSFER_BT470BG) return std::pow(x, 1.f / 2.8f);
    return std::pow(x * 203.f / bt1886_alpha(), 1.f / 2.4f) - bt1886_beta();
  }
  // ITU-R BT.1886 with Lw = 203 cd/m2, Lb = 0.1 cd/m2: L = alpha * max(x + beta, 0)^2.4
  static float inv_white() { return std::pow(203.f, 1.f / 2.4f); }
  static float inv_black() { return std::pow(0.1f, 1.f / 2.4f); }
  static float bt1886_alpha() { return std::pow(inv_white() - inv_black(), 2.4f); }
  static float bt1886_beta() { return inv_black() / (inv_white() - inv_black()); }
};
}  // namespace

int g1s_generate_photon_noise(uint32_t iso, uint32_t width, uint32_t height, int transfer, int chroma_grain,
                              int full_range, int32_t random_seed, uint64_t start_time, uint64_t end_time,
                              g1s_segment *out) {
  if (!out || iso == 0 || width == 0 || height == 0 || transfer < 0 || transfer > G1S_TRANSFER_BT470BG) return G1S_E_ARG;
  const Transfer tf{transfer};
  std::memset(out, 0, sizeof *out);
  const float kPhotonsPerLxSPerUm2 = 11260.f, kEffectiveQuantumEfficiency = 0.20f, kPhotoResponseNonUniformity = 0.005f,
              kInputReferredReadNoise = 1.5f;
  const float mid_tone_exposure = 10.f / (float)iso;                                  // lx.s on an 18 % card
  const float pixel_area_um2 = (36000 * 24000.f) / ((float)width * (float)height);   // 35 mm sensor
  const float mid_tone_electrons = kEffectiveQuantumEfficiency * kPhotonsPerLxSPerUm2 * mid_tone_exposure * pixel_area_um2;
  const float max_electrons = mid_tone_electrons / tf.to_linear(0.5f);
  out->num_y_points = G1S_NUM_Y_POINTS;
  for (int i = 0; i < G1S_NUM_Y_POINTS; ++i) {
    const float x = (float)i / (G1S_NUM_Y_POINTS - 1.f);
    const float linear = tf.to_linear(x);
    const float electrons = max_electrons * linear;
    // read noise, photon shot noise and photo-response non-uniformity added in quadrature (electrons rms)
    const float noise = std::sqrt(kInputReferredReadNoise * kInputReferredReadNoise + electrons +
                                  kPhotoResponseNonUniformity * kPhotoResponseNonUniformity * electrons * electrons);
    const float linear_noise = noise / max_electrons;
    const float lo = std::fmax(0.f, linear - 2 * linear_noise), hi = std::fmin(1.f, linear + 2 * linear_noise);
    const float slope = (tf.from_linear(hi) - tf.from_linear(lo)) / (hi - lo);
    const float encoded_noise = linear_noise * slope;
    // NoiseGenArgs::full_range (src/main.rs:299: color_range == JPEG).  Full range: code values and noise on the 0..255
    // scale (the form libaom's photon_noise_table.c has, pinned by the reference's fixture).  Limited range: the same
    // curve on the 16..235 scale -- recalled from the crate, which is absent here, so this branch is parity-unpinned.
    const float span = full_range ? 255.f : 219.f, base = full_range ? 0.f : 16.f;
    out->scaling_points_y[i][0] = (uint8_t)std::round(base + span * x);
    out->scaling_points_y[i][1] = (uint8_t)std::fmin(255.f, std::round(span * 7.88f * encoded_noise));
  }
  out->start_time = start_time;
  out->end_time = end_time;
  out->scaling_shift = 8;
  out->ar_coeff_lag = 0;
  out->ar_coeff_shift = 6;
  out->num_ar_coeffs_plus1[0] = 1;  // no luma coefficients, a single 0 per chroma plane (as a parsed lag-0 header)
  out->num_ar_coeffs_plus1[1] = out->num_ar_coeffs_plus1[2] = 2;
  out->overlap_flag = 1;
  out->chroma_scaling_from_luma = chroma_grain ? 1 : 0;
  out->random_seed = random_seed < 0 ? DEFAULT_GRAIN_SEED : (uint16_t)random_seed;
  out->clip_to_restricted_range = 1;
  return G1S_OK;
}

int g1s_inspect_stream_info(const g1s_inspect *h, g1s_stream_info *info) {
  if (!h || !info) return G1S_E_ARG;
  std::memset(info, 0, sizeof *info);
  const SequenceHeader &s = h->seq;
  info->have_sequence_header = s.valid;
  info->seq_profile = s.seq_profile;
  info->bit_depth = s.bit_depth;
  info->monochrome = s.num_planes == 1;
  info->ss_x = s.ss_x;
  info->ss_y = s.ss_y;
  info->max_frame_width = (int32_t)s.max_frame_width_minus_1 + 1;
  info->max_frame_height = (int32_t)s.max_frame_height_minus_1 + 1;
  info->film_grain_params_present = s.film_grain_params_present;
  info->color_primaries = s.color_primaries;
  info->transfer_characteristics = s.transfer_characteristics;
  info->matrix_coefficients = s.matrix_coefficients;
  info->color_range = s.color_range;
  info->order_hint_bits = s.order_hint_bits;
  info->reduced_still_picture_header = s.reduced_still_picture_header;
  info->packets = h->packets;
  info->obus = h->obus;
  info->last_frame_width = (int32_t)h->last_frame_width;
  info->last_frame_height = (int32_t)h->last_frame_height;
  info->last_tile_cols = (int32_t)h->cur_tile_info.tile_cols;
  info->last_tile_rows = (int32_t)h->cur_tile_info.tile_rows;
  return G1S_OK;
}

// IVF ("DKIF", 32-byte file header, 12-byte frame headers) or a Section-5 low-overhead OBU stream split at temporal
// delimiters.  fps 0/0 takes the IVF header's rate/scale.
int g1s_inspect_push_file(g1s_inspect *h, const char *path, int64_t *fps_num, int64_t *fps_den) {
  if (!h || !path) return G1S_E_ARG;
  FILE *f = std::fopen(path, "rb");
  if (!f) {
    h->err = std::string("cannot open ") + path;
    return G1S_E_IO;
  }
  std::vector<uint8_t> buf;
  uint8_t tmp[1 << 16];
  size_t got;
  while ((got = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
  std::fclose(f);
  auto le32 = [&](size_t o) { return (uint32_t)buf[o] | (uint32_t)buf[o + 1] << 8 | (uint32_t)buf[o + 2] << 16 | (uint32_t)buf[o + 3] << 24; };
  if (buf.size() >= 32 && std::memcmp(buf.data(), "DKIF", 4) == 0) {
    const size_t hdr_len = (size_t)buf[6] | (size_t)buf[7] << 8;
    if (std::memcmp(buf.data() + 8, "AV01", 4) != 0) {
      h->err = "IVF stream is not AV01";
      return G1S_E_STREAM;
    }
    if (fps_num && fps_den && (*fps_num <= 0 || *fps_den <= 0)) {
      *fps_num = le32(16);
      *fps_den = le32(20);
    }
    size_t off = hdr_len < 32 ? 32 : hdr_len;
    while (off + 12 <= buf.size()) {
      const size_t sz = le32(off);
      off += 12;
      if (off + sz > buf.size()) {
        h->err = "truncated IVF frame";
        return G1S_E_STREAM;
      }
      const int rc = g1s_inspect_push_packet(h, buf.data() + off, sz);
      if (rc != G1S_OK) return rc;
      off += sz;
    }
    return G1S_OK;
  }
  // Section 5 stream: every OBU carries a size; a packet is the run of OBUs from one temporal delimiter to the next
  size_t off = 0, start = 0;
  bool first = true;
  while (off < buf.size()) {
    const uint8_t b0 = buf[off];
    const int type = (b0 >> 3) & 0xF;
    const bool has_ext = (b0 >> 2) & 1, has_size = (b0 >> 1) & 1;
    if ((b0 & 0x81) || !has_size) {
      h->err = "not an IVF file and not a low-overhead OBU stream with size fields";
      return G1S_E_STREAM;
    }
    uint64_t v;
    size_t used;
    const size_t hdr = 1 + (has_ext ? 1 : 0);
    if (off + hdr > buf.size() || !read_leb128(buf.data() + off + hdr, buf.size() - off - hdr, &v, &used) ||
        off + hdr + used + v > buf.size()) {
      h->err = "truncated OBU stream";
      return G1S_E_STREAM;
    }
    if (type == OBU_TEMPORAL_DELIMITER && !first) {
      const int rc = g1s_inspect_push_packet(h, buf.data() + start, off - start);
      if (rc != G1S_OK) return rc;
      start = off;
    }
    first = false;
    off += hdr + used + (size_t)v;
  }
  if (off > start) return g1s_inspect_push_packet(h, buf.data() + start, off - start);
  return G1S_OK;
}

// Test hook: one syntax element group on a raw bit buffer; returns the number of bits consumed (or a negative status).
// Used by tests/test_inspect.py to replay the reference's unit-test vectors (frame.rs / grain.rs / sequence.rs tests).
int64_t g1s_obu_probe(const char *what, const uint8_t *data, size_t size, const int64_t *a, size_t na, int64_t *out,
                      size_t nout, g1s_segment *seg) {
  if (!what || (!data && size)) return G1S_E_ARG;
  auto arg = [&](size_t i) -> int64_t { return i < na ? a[i] : 0; };
  auto put = [&](size_t i, int64_t v) {
    if (out && i < nout) out[i] = v;
  };
  try {
    BitReader br(data, size);
    const std::string w = what;
    if (w == "film_grain_params") {  // allowed, frame_type, monochrome, ss_x, ss_y -> kind
      const GrainHeader h = film_grain_params(br, arg(0) != 0, (int)arg(1), arg(2) != 0, (int)arg(3), (int)arg(4));
      put(0, h.kind);
      if (seg) *seg = h.params;
    } else if (w == "tile_info") {  // use_128, mi_cols, mi_rows -> cols, rows, cols_log2, rows_log2
      const TileInfo t = tile_info(br, arg(0) != 0, (uint32_t)arg(1), (uint32_t)arg(2));
      put(0, t.tile_cols), put(1, t.tile_rows), put(2, t.tile_cols_log2), put(3, t.tile_rows_log2);
    } else if (w == "quantization_params") {  // num_planes, separate_uv -> base, ydc, udc, uac, vdc, vac
      const QuantParams q = quantization_params(br, (int)arg(0), arg(1) != 0);
      put(0, q.base_q_idx), put(1, q.y_dc), put(2, q.u_dc), put(3, q.u_ac), put(4, q.v_dc), put(5, q.v_ac);
    } else if (w == "segmentation_params") {  // primary_ref_frame -> enabled, then has/value of [seg a1][feature a2]
      const Segmentation sg = segmentation_params(br, (int)arg(0));
      put(0, sg.enabled), put(1, sg.has[arg(1) & 7][arg(2) & 7]), put(2, sg.value[arg(1) & 7][arg(2) & 7]);
    } else if (w == "delta_q_params") {
      put(0, delta_q_params(br, (int)arg(0)));
    } else if (w == "delta_lf_params") {
      delta_lf_params(br, arg(0) != 0, arg(1) != 0);
    } else if (w == "loop_filter_params") {
      loop_filter_params(br, arg(0) != 0, arg(1) != 0, (int)arg(2));
    } else if (w == "cdef_params") {
      cdef_params(br, arg(0) != 0, arg(1) != 0, arg(2) != 0, (int)arg(3));
    } else if (w == "lr_params") {  // all_lossless, intrabc, enable_restoration, use_128, num_planes, ss_x, ss_y
      SequenceHeader s;
      s.enable_restoration = arg(2) != 0, s.use_128x128_superblock = arg(3) != 0, s.num_planes = (int)arg(4);
      s.ss_x = (int)arg(5), s.ss_y = (int)arg(6);
      lr_params(br, arg(0) != 0, arg(1) != 0, s);
    } else if (w == "skip_mode_params") {  // intra, reference_select, order_hint_bits, order_hint, hints[8], idx[7]
      uint64_t hints[8];
      int idx[7];
      for (int i = 0; i < 8; ++i) hints[i] = (uint64_t)arg(4 + i);
      for (int i = 0; i < 7; ++i) idx[i] = (int)arg(12 + i) & 7;
      const bool allowed = skip_mode_allowed(arg(0) != 0, arg(1) != 0, (int)arg(2), (uint64_t)arg(3), hints, idx);
      if (allowed) br.flag();
      put(0, allowed);
    } else if (w == "global_motion_params") {
      global_motion_params(br, arg(0) != 0, arg(1) != 0);
    } else if (w == "decode_subexp") {
      put(0, decode_subexp(br, (int32_t)arg(0)));
    } else if (w == "decode_signed_subexp_with_ref") {
      put(0, decode_signed_subexp_with_ref(br, (int32_t)arg(0), (int32_t)arg(1), (int32_t)arg(2)));
    } else if (w == "decode_unsigned_subexp_with_ref") {
      put(0, decode_unsigned_subexp_with_ref(br, (int32_t)arg(0), (int32_t)arg(1)));
    } else if (w == "inverse_recenter") {
      put(0, inverse_recenter((int32_t)arg(0), (int32_t)arg(1)));
    } else if (w == "get_relative_dist") {
      put(0, get_relative_dist(arg(0), arg(1), (int)arg(2)));
    } else if (w == "ns") {
      put(0, (int64_t)br.ns((uint64_t)arg(0)));
    } else if (w == "su") {
      put(0, br.su((unsigned)arg(0)));
    } else if (w == "uvlc") {
      put(0, (int64_t)br.uvlc());
    } else if (w == "leb128") {
      uint64_t v;
      size_t used;
      if (!read_leb128(data, size, &v, &used)) return G1S_E_STREAM;
      put(0, (int64_t)v);
      return (int64_t)used * 8;
    } else if (w == "sequence_header") {
      const SequenceHeader s = parse_sequence_header(br);
      put(0, s.seq_profile), put(1, s.bit_depth), put(2, s.num_planes), put(3, s.ss_x), put(4, s.ss_y);
      put(5, s.film_grain_params_present), put(6, s.order_hint_bits), put(7, s.max_frame_width_minus_1 + 1);
      put(8, s.max_frame_height_minus_1 + 1), put(9, s.reduced_still_picture_header);
      put(10, s.operating_points_cnt_minus_1), put(11, s.cur_operating_point_idc), put(12, s.has_decoder_model);
      put(13, s.frame_id_numbers_present), put(14, s.force_screen_content_tools), put(15, s.force_integer_mv);
    } else {
      return G1S_E_ARG;
    }
    return (int64_t)br.pos;
  } catch (const std::exception &) {
    return G1S_E_STREAM;
  }
}

}  // extern "C"
