// Host side of the grain-estimation path: everything av1-grain's NoiseModel does AFTER
// the per-pixel sums exist (the "tiny f64 solves" of SURVEY.md section 8a rows a7, a9-a12).
// Input is one integer FrameRecord per frame pair, produced by the CUDA kernels; this
// file never touches pixels.  Strictly sequential across frames, as in the reference.
#pragma once
#include <functional>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/g1s.h"

namespace g1s {

struct StreamGeometry {
  int width = 0, height = 0, ss_x = 1, ss_y = 1, planes = 3;
  int nbw = 0, nbh = 0, nb = 0;
};

// View over one per-frame record in host memory (see RecordLayout in g1s_kernels.h).
struct FrameRecordView {
  const int64_t *gram;       // [3][351]
  const int64_t *nobs;       // [3]
  int64_t num_flat;
  const uint32_t *luma_sum;  // [nb]
  const int32_t *rsum;       // [3][nb]
  const uint32_t *rsq;       // [3][nb]
  const uint8_t *flat;       // [nb]
  const double *gramf = nullptr;  // [3][351] strict mode: reference-order f64 sums of product / 255^2 (else null)
};

enum class NoiseStatus { Ok, DifferentType, Error };

class LinearSystem {
 public:
  explicit LinearSystem(int n = 0) { reset(n); }
  void reset(int n);
  void clear();
  bool solve();  // Gaussian elimination on copies of A and b (EquationSystem::solve)
  void add(const LinearSystem &o);
  void copy_from(const LinearSystem &o);
  int n = 0;
  std::vector<double> A, b, x;
};

class StrengthSolver {
 public:
  StrengthSolver();
  void clear();
  void add_measurement(double block_mean, double noise_std);
  bool solve();         // = bump_b() + solve_bumped(), the reference's NoiseStrengthSolver::solve
  void bump_b();        // its in-place side effect on b alone
  bool solve_bumped();  // its elimination alone, on the current b
  double value_at(double intensity) const;
  double bin_center(int i) const;
  void add(const StrengthSolver &o);
  // (intensity, strength) points after greedy simplification (fit_piecewise)
  std::vector<std::pair<double, double>> fit_piecewise(int max_points) const;
  LinearSystem eqns;
  int num_bins;
  int num_equations = 0;
  double total = 0;
  double min_intensity = 0, max_intensity = 255;

 private:
  double bin_index(double v) const;
  void update_residual(const std::vector<std::pair<double, double>> &pts, std::vector<double> &residual, int start,
                       int end) const;
};

struct ChannelState {
  explicit ChannelState(int n = 24) : eqns(n) {}
  LinearSystem eqns;
  double ar_gain = 1.0;
  int64_t num_observations = 0;
  StrengthSolver strength;
  bool solve_ar(bool is_chroma);
};

// Everything NoiseModel::update derives from ONE frame before it looks at the combined state: the
// three "latest" channel states.  A pure function of the frame's record, so the frames of a batch
// are evaluated in parallel (host thread pool) and only `fold` runs in frame order.
struct LatestFrame {
  LatestFrame() : ch{ChannelState(24), ChannelState(25), ChannelState(25)} {}
  ChannelState ch[3];
  bool enough_flat = false;
  int channels = 0;       // channels evaluated (stops at the first failing one)
  int fail_channel = -1;  // channel whose evaluation hit NoiseStatus::Error, -1 if none
  const char *fail_text = nullptr;
  // scratch reused across frames (per-block bin data shared by the three channels)
  std::vector<int> bin0;
  std::vector<int> contrib;  // blocks that add a strength measurement (flat, more than block_size samples), in block order
  std::vector<double> frac, mean, strength;

  // Fixed-size little-endian image of the state above (the "digest"): what a producer rank sends to the
  // rank that owns the sequential model, 11.1 KB per frame whatever the frame size (symmetric systems are sent
  // as their upper triangle, the tridiagonal strength system as two diagonals).
  static constexpr size_t kDigestDoubles = 4 + 3 * (325 + 25 + 25 + 2 + 20 + 20 + 20 + 20 + 2);
  void to_digest(double *out) const;
  void from_digest(const double *in);
};

class NoiseModel {
 public:
  explicit NoiseModel(const StreamGeometry &g);
  // NoiseModel::update = compute_latest (thread-safe, const) + fold (sequential)
  void compute_latest(const FrameRecordView &rec, LatestFrame &out) const;
  NoiseStatus fold(const LatestFrame &lf);
  // fold() over a run of frames with the per-frame solves of the combined luma state off the calling thread.
  // What fold does to luma is a chain only through SUMS (state_k = state_k-1 + frame_k, rounded adds in frame order);
  // the solves of state_k feed nothing but the is_different verdict on frame k+1.  So, assuming no frame of the run
  // differs: (1) the sums of every prefix, sequentially; (2) their solves, independent of each other, through `par`;
  // (3) the verdicts in order.  Frames up to the first that differs (or whose combined solve fails) are committed --
  // bit for bit the state fold() would have left -- and their count is returned; the caller hands the next frame to
  // fold().  Every frame must be "plain": enough flat blocks, no failed channel.
  using ParallelFor = std::function<void(int, const std::function<void(int)> &)>;
  int fold_run(const LatestFrame *frames, int count, const ParallelFor &par);
  static bool plain(const LatestFrame &lf, int planes) { return lf.enough_flat && lf.fail_channel < 0 && lf.channels == planes; }
  int planes() const { return g_.planes; }
  NoiseStatus update(const FrameRecordView &rec);
  void save_latest();
  // Deferred work of fold(): the chroma solves of the combined state (see fold).  Must run before the
  // combined chroma solutions are read (grain_parameters).
  void settle();
  void grain_parameters(uint64_t start_ts, uint64_t end_ts, g1s_segment *seg) const;
  const std::string &last_error() const { return err_; }
  ChannelState combined[3];

 private:
  void load_equations(int c, const FrameRecordView &rec, ChannelState &st) const;
  void add_strength_measurements(int c, const FrameRecordView &rec, LatestFrame &lf) const;
  bool is_different(const LatestFrame &lf) const;
  static bool differs(const LatestFrame &lf, const double *combined_ar_x, const double *combined_strength_x);
  StreamGeometry g_;
  std::string err_;
  LatestFrame scratch_;        // used by update()
  const LatestFrame *last_ = nullptr;  // the frame save_latest() copies from
  bool stale_[3] = {false, false, false};  // combined[c] has sums newer than its solutions
  bool same_blocks_;           // luma and chroma contribute the same blocks to the strength solver
  std::vector<int> cnt_luma_, cnt_chroma_;  // samples per frame-clipped block
  std::vector<double> inv_luma_, inv_chroma_;  // exact reciprocals where the count is a power of two
  std::vector<int> odd_luma_, odd_chroma_;  // blocks whose count is not (frame-edge slivers)
};

// DiffGenerator minus the pixels: frame counter, timestamps, segment list.
class DiffSequencer {
 public:
  DiffSequencer(int64_t fps_num, int64_t fps_den, const StreamGeometry &g);
  void consume(const FrameRecordView &rec);
  void consume_latest(const LatestFrame &lf);  // same, for a frame whose latest state is already evaluated
  // consume_latest over consecutive frames; long runs of plain frames go through NoiseModel::fold_run
  void consume_latest_batch(const LatestFrame *frames, int count, const NoiseModel::ParallelFor &par);
  std::vector<g1s_segment> finish();
  int64_t frames() const { return frame_count_; }
  NoiseModel &model() { return model_; }

 private:
  void after_update(NoiseStatus st);
  int64_t fps_num_, fps_den_;
  int64_t frame_count_ = 0;
  uint64_t prev_timestamp_ = 0;
  NoiseModel model_;
  std::vector<g1s_segment> table_;
};

// (A^T A)^-1 of FlatBlockFinder::new, row-major 3x3.
void flat_block_ata_inv(double out[9]);

// `filmgrn1` text, src/main.rs:525-530 + 631-696.
std::string format_grain_table(const g1s_segment *segs, size_t n);

}  // namespace g1s
