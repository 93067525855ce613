// The engine behind include/g1s.h: owns the device buffers, the CUDA stream and the
// host noise model, and turns pushed frame pairs into grain-table segments.
//
// Data flow per batch ("slot") of B frame pairs:
//   caller planes --memcpy--> pinned staging --cudaMemcpyAsync--> HBM frame store
//   HBM frames --flat_features / flat_select / gram kernels--> HBM per-frame records
//   records --cudaMemcpyAsync--> pinned host --DiffSequencer (f64 solves, in frame order)
// Three slots rotate, so the caller fills slot k+1 while the device works on slot k and
// the records of slot k-1 are folded into the model.  There is no CPU fallback: every
// device failure is reported as G1S_E_CUDA.
#include <cuda.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "../../include/g1s.h"
#include "g1s_filters.h"
#include "g1s_kernels.h"
#include "g1s_model.h"
#include "g1s_pool.h"

using namespace g1s;

namespace {

thread_local std::string g_create_error;

// A block of digests on its way through the fold thread: as received, and unpacked into model states
struct DigestBlock {
  std::vector<double> raw;
  std::vector<LatestFrame> frames;
};

constexpr int kSlots = 4;

// The sequential half of the host model (merge into the combined state, segment cuts) runs on its own
// thread, strictly in submission order, so neither the kernels' launches nor the caller wait for it.
class FoldQueue {
 public:
  FoldQueue() : th_([this] { loop(); }) {}
  ~FoldQueue() {
    {
      std::lock_guard<std::mutex> l(m_);
      stop_ = true;
    }
    cv_.notify_all();
    th_.join();
  }
  uint64_t push(std::function<void()> fn) {
    std::lock_guard<std::mutex> l(m_);
    q_.push_back(std::move(fn));
    cv_.notify_all();
    return ++submitted_;
  }
  void wait(uint64_t ticket) {
    std::unique_lock<std::mutex> l(m_);
    done_cv_.wait(l, [&] { return done_ >= ticket; });
  }
  void wait_all() { wait(submitted_snapshot()); }

 private:
  uint64_t submitted_snapshot() {
    std::lock_guard<std::mutex> l(m_);
    return submitted_;
  }
  void loop() {
    for (;;) {
      std::function<void()> fn;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        fn = std::move(q_.front());
        q_.pop_front();
      }
      fn();
      {
        std::lock_guard<std::mutex> l(m_);
        ++done_;
      }
      done_cv_.notify_all();
    }
  }
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  std::deque<std::function<void()>> q_;
  uint64_t submitted_ = 0, done_ = 0;
  bool stop_ = false;
  std::thread th_;
};

struct PlaneGeom {
  int w, h;          // storage size in samples
  size_t pitch[2];   // [0] source, [1] denoised: bytes per row in the HBM frame store
  size_t off[2];     // byte offset inside one frame-pair slot entry
};

struct Slot {
  uint8_t *d_frames = nullptr;   // B * pair_bytes
  uint8_t *h_frames = nullptr;   // pinned mirror
  FrameDesc *d_descs = nullptr;
  FrameDesc *h_descs = nullptr;  // pinned
  int8_t *d_res = nullptr;           // B * ResidualStore::frame_bytes: s8 residual + luma-tap planes (tensor-core path)
  CUtensorMap *d_rmaps = nullptr;    // [B][kResidualMaps] TMA descriptors of d_res, built once
  void *d_plan = nullptr;            // gram_plan_kernel's unit descriptors of the batch
  int *d_plan_counts = nullptr;      // [B][3] units per frame and plane
  uint8_t *d_records = nullptr;
  uint8_t *h_records = nullptr;  // pinned
  double *d_digests = nullptr;   // [B][kDigestDoubles]: latest_kernel's output (device-side per-frame model half)
  double *h_digests = nullptr;   // pinned
  bool records_read = false;     // the batch's records were copied back too (record tap / host model)
  cudaEvent_t done = nullptr, k0_beg = nullptr, k0_end = nullptr, kr_beg = nullptr, k1_beg = nullptr, k1_end = nullptr,
              ks_beg = nullptr, ks_end = nullptr, copied = nullptr, computed = nullptr;
  int count = 0;                 // frames staged
  int host_frames = 0;           // of which need the H2D copy (contiguous prefix is not required)
  bool in_flight = false;
  std::vector<LatestFrame> latest;  // per-frame model state of the batch, owned until its fold job is done
  uint64_t fold_ticket = 0;
};

}  // namespace

struct g1s_diff {
  g1s_diff_config cfg{};
  Geometry geom{};
  StreamGeometry sgeom{};
  RecordLayout rl{};
  FlatConsts fc{};
  PlaneGeom pg[3]{};
  size_t pair_bytes = 0;
  int batch = 1;
  // cfg.host_narrow: samples wider than 8 bits are reduced (truncating >>, frame_into_u8) while they are staged, so
  // the device, and the PCIe link, only ever see 8-bit planes; host_* keep what the caller's planes look like
  bool narrow = false;
  int host_bytes[2] = {1, 1}, host_shift[2] = {0, 0};
  cudaStream_t stream = nullptr;       // kernels of even batches (and the benchmark marks)
  cudaStream_t more[3] = {nullptr, nullptr, nullptr};  // kernel streams of the other batches in flight: consecutive
                                       // batches overlap on the device, so the FP64-bound flat-block finder of one
                                       // runs beside the HBM- / tensor-bound residual and Gram kernels of another
  int nstreams = 1;                    // G1S_STREAMS, default 3
  cudaEvent_t marks[2] = {nullptr, nullptr}, join = nullptr;
  uint64_t submitted = 0;
  cudaStream_t copy_stream = nullptr;  // per-frame host->device copies, overlapping the staging of the next frame
  cudaStream_t d2h_stream = nullptr;   // record read-back, overlapping the kernels of the next batch
  Slot slots[kSlots];
  int cur = 0;            // slot being filled
  int oldest = 0;         // oldest slot possibly in flight
  std::unique_ptr<DiffSequencer> seq;
  std::unique_ptr<HostPool> pool;
  std::unique_ptr<HostPool> fold_pool;  // the fold thread's helpers (NoiseModel::fold_run); nobody else uses it
  NoiseModel::ParallelFor fold_par;
  std::unique_ptr<FoldQueue> folder;
  std::vector<LatestFrame> latest;  // scratch for the single-record entry points
  std::string err;
  bool finished = false;
  int64_t pushed = 0;
  int64_t retired = 0;
  g1s_record_fn tap = nullptr;
  void *tap_user = nullptr;
  std::mutex digest_mu;    // consumer handles: recycled copies of incoming digest blocks
  std::vector<std::shared_ptr<DigestBlock>> digest_free;  // recycled
  // source filters (g1s_diff_set_source_filters): the chain runs on the copy stream between the upload of the raw
  // source planes (a small ring, reused in stream order) and the kernels of the path
  std::unique_ptr<SourceFilters> filters;
  std::vector<g1s_filter_op> filter_ops;  // kept for the children of a multi-device handle
  int raw_w = 0, raw_h = 0;
  size_t raw_pitch[3] = {0, 0, 0}, raw_off[3] = {0, 0, 0}, raw_bytes = 0;
  static constexpr int kRaw = 3;
  uint8_t *raw_host[kRaw] = {nullptr, nullptr, nullptr}, *raw_dev[kRaw] = {nullptr, nullptr, nullptr};
  cudaEvent_t raw_ev[kRaw] = {nullptr, nullptr, nullptr};
  int raw_next = 0;
  // multi-device handles (cfg.n_devices >= 2): one PRODUCER child per device, this handle owns the model
  std::vector<g1s_diff *> kids;
  std::vector<double *> kid_sinks;   // pinned digest rings the children write into
  std::vector<int64_t> kid_frames;   // frames dealt to each child
  size_t kid_sink_cap = 0;           // digests per ring (a multiple of the batch)
  int64_t next_batch = 0;            // next global batch whose digests are folded
  double *sink = nullptr;  // digest sink (producer ranks)
  size_t sink_cap = 0, sink_count = 0;
  // cuTensorMapEncodeTiled, fetched through the runtime so libcuda is not a link dependency
  typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn encode_tiled = nullptr;
  double tma_batches = 0, vector_batches = 0;
  ResidualStore rstore{};
  bool tensor_path = false;  // residual_kernel + gram_imma_kernel available for this stream
  bool device_model = false;  // latest_kernel evaluates the per-frame model half: only digests are read back
  // counters
  double kernels_launched = 0, k1_ms = 0, k1_launches = 0, k0_ms = 0, k0_launches = 0, frames_done = 0, kr_ms = 0,
         ks_ms = 0;
};

namespace {

#define CU_TRY(d, expr)                                                                         \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      (d)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                            \
      return G1S_E_CUDA;                                                                        \
    }                                                                                           \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Handles own streams, events and memory of ONE device; a caller (or a multi-device parent) may have made another
// device current since the last call.
inline bool owns_device(const g1s_diff *d) { return d->stream != nullptr; }
#define G1S_ON_DEVICE(d)                                                    \
  do {                                                                      \
    if (owns_device(d)) CU_TRY(d, cudaSetDevice((d)->cfg.device));          \
  } while (0)

FrameRecordView view_of(const g1s_diff *d, const uint8_t *rec) {
  FrameRecordView v;
  v.gram = reinterpret_cast<const int64_t *>(rec + d->rl.off_gram);
  v.nobs = reinterpret_cast<const int64_t *>(rec + d->rl.off_nobs);
  v.num_flat = *reinterpret_cast<const int64_t *>(rec + d->rl.off_num_flat);
  v.luma_sum = reinterpret_cast<const uint32_t *>(rec + d->rl.off_luma_sum);
  v.rsum = reinterpret_cast<const int32_t *>(rec + d->rl.off_rsum);
  v.rsq = reinterpret_cast<const uint32_t *>(rec + d->rl.off_rsq);
  v.flat = rec + d->rl.off_flat;
  v.gramf = d->cfg.gram_order == G1S_GRAM_REF_ORDER ? reinterpret_cast<const double *>(rec + d->rl.off_gramf) : nullptr;
  return v;
}

// Per-frame model evaluation in parallel, then tap + merge in frame order.
// Per-frame model half in parallel on the host pool, taps / digests on the caller's thread in frame
// order, then the sequential merge is queued for the fold thread.  `store` must stay untouched until the
// returned ticket is done (0: nothing queued).
uint64_t fold_records(g1s_diff *d, const uint8_t *recs, int count, size_t stride, std::vector<LatestFrame> &store,
                      const double *digests = nullptr, bool have_records = true) {
  const bool model = d->cfg.mode != G1S_MODE_PRODUCER;
  constexpr size_t K = LatestFrame::kDigestDoubles;
  if (model || (d->sink && !digests)) {
    if ((int)store.size() < count) store.resize(count);
    if (digests) {
      // the device evaluated the per-frame half (latest_kernel): rebuild the states from its digests
      for (int i = 0; i < count; ++i) store[i].from_digest(digests + K * i);
    } else {
      const NoiseModel &nm = d->seq->model();
      d->pool->parallel_for(count, [&](int i) { nm.compute_latest(view_of(d, recs + (size_t)i * stride), store[i]); });
    }
  }
  for (int i = 0; i < count; ++i) {
    if (d->tap && have_records) d->tap(d->tap_user, d->retired, recs + (size_t)i * stride, d->rl.bytes);
    if (d->sink && d->sink_cap) {
      double *slot = d->sink + K * (d->sink_count++ % d->sink_cap);
      if (digests) std::memcpy(slot, digests + K * i, sizeof(double) * K);
      else store[i].to_digest(slot);
    }
    d->retired++;
  }
  if (!model) return 0;
  DiffSequencer *seq = d->seq.get();
  LatestFrame *frames = store.data();
  const NoiseModel::ParallelFor *par = &d->fold_par;
  return d->folder->push([seq, frames, count, par] { seq->consume_latest_batch(frames, count, *par); });
}

// Builds (once per slot) the TMA descriptors of the slot's residual planes: per frame of the batch the s8
// residual of Y, Cb, Cr and the two luma-tap halves.  Extents are the LOOP extents (floor for chroma), so
// everything beyond reads as zero, like the reference's frame clipping.
bool build_residual_maps(g1s_diff *d, Slot &s, std::vector<CUtensorMap> &host) {
  const Geometry &g = d->geom;
  const ResidualStore &rs = d->rstore;
  int box[kResidualMaps][2];
  gram_imma_tma_boxes(box);
  host.assign((size_t)d->batch * kResidualMaps, CUtensorMap{});
  for (int i = 0; i < d->batch; ++i) {
    int8_t *fb = s.d_res + (size_t)i * rs.frame_bytes;
    for (int k = 0; k < (g.planes == 3 ? kResidualMaps : 1); ++k) {
      const bool luma = k == 0;
      const cuuint64_t dims[2] = {(cuuint64_t)(luma ? g.width : g.width >> 1), (cuuint64_t)(luma ? g.height : g.height >> 1)};
      const cuuint64_t strides[1] = {(cuuint64_t)(luma ? rs.pitch_l : rs.pitch_c)};
      const cuuint32_t bx[2] = {(cuuint32_t)box[k][0], (cuuint32_t)box[k][1]};
      const cuuint32_t estr[2] = {1, 1};
      void *base = fb + (k < 3 ? rs.off_res[k] : rs.off_tap);
      const CUresult r = d->encode_tiled(&host[(size_t)i * kResidualMaps + k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims,
                                         strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return false;
    }
  }
  return true;
}

int submit(g1s_diff *d, Slot &s) {
  if (s.count == 0) return G1S_OK;
  const int which = (int)(d->submitted++ % (uint64_t)d->nstreams);
  cudaStream_t st = which ? d->more[which - 1] : d->stream;
  if (s.host_frames > 0) {  // frames were sent one by one on the copy stream as they were pushed
    CU_TRY(d, cudaEventRecord(s.copied, d->copy_stream));
    CU_TRY(d, cudaStreamWaitEvent(st, s.copied, 0));
  }
  CU_TRY(d, cudaMemcpyAsync(s.d_descs, s.h_descs, sizeof(FrameDesc) * s.count, cudaMemcpyHostToDevice, st));
  CU_TRY(d, cudaMemsetAsync(s.d_records, 0, d->rl.bytes * s.count, st));
  int gram_launches = 1;
  if (d->tensor_path) {
    // Tensor-core path: the streaming residual kernel goes first (it needs no flat flags) and leaves, beside the s8
    // residual planes, the 8-bit source luma the flat-block finder reads: half the bytes of a 10-bit frame, aligned,
    // and fresh in L2.  128-bit loads need 16-byte aligned rows; otherwise the residual kernel uses scalar loads.
    bool aligned = !std::getenv("G1S_SCALAR_LOADS");
    for (int i = 0; i < s.count && aligned; ++i)
      for (int c = 0; c < d->geom.planes; ++c) {
        const FrameDesc &fd = s.h_descs[i];
        if (((uintptr_t)fd.src[c] | (uintptr_t)fd.den[c] | fd.src_stride[c] | fd.den_stride[c]) & 15) aligned = false;
      }
    ResidualStore rs = d->rstore;
    rs.base = s.d_res;
    CU_TRY(d, cudaEventRecord(s.kr_beg, st));
    launch_residual(s.d_descs, s.count, d->geom, rs, s.d_records, d->rl, aligned, st);
    CU_TRY(d, cudaEventRecord(s.k0_beg, st));
    launch_flat_features(s.d_descs, s.count, d->geom, d->fc, s.d_records, d->rl, true, st,
                         reinterpret_cast<const uint8_t *>(s.d_res) + rs.off_y8, rs.frame_bytes, rs.pitch_l);
    CU_TRY(d, cudaEventRecord(s.k0_end, st));
    launch_flat_select(s.count, d->geom, s.d_records, d->rl, st);
    CU_TRY(d, cudaEventRecord(s.k1_beg, st));
    launch_gram_plan(s.count, d->geom, s.d_records, d->rl, s.d_plan, s.d_plan_counts, st);
    launch_gram_imma(s.count, d->geom, s.d_records, d->rl, s.d_rmaps, s.d_plan, s.d_plan_counts, st);
    CU_TRY(d, cudaEventRecord(s.k1_end, st));
    launch_gram_generic(s.d_descs, s.count, d->geom, s.d_records, d->rl, /*only_overflow=*/true, st);
    gram_launches = 4;
    d->tma_batches += 1;
    if (aligned) d->vector_batches += 1;
  } else {
    CU_TRY(d, cudaEventRecord(s.k0_beg, st));
    bool luma16 = true;  // 16-byte aligned source-luma rows -> 128-bit loads in the flat-block kernel
    for (int i = 0; i < s.count && luma16; ++i)
      if (((uintptr_t)s.h_descs[i].src[0] | s.h_descs[i].src_stride[0]) & 15) luma16 = false;
    launch_flat_features(s.d_descs, s.count, d->geom, d->fc, s.d_records, d->rl, luma16, st);
    CU_TRY(d, cudaEventRecord(s.k0_end, st));
    launch_flat_select(s.count, d->geom, s.d_records, d->rl, st);
    CU_TRY(d, cudaEventRecord(s.kr_beg, st));
    CU_TRY(d, cudaEventRecord(s.k1_beg, st));
    launch_gram_generic(s.d_descs, s.count, d->geom, s.d_records, d->rl, /*only_overflow=*/false, st);
    CU_TRY(d, cudaEventRecord(s.k1_end, st));
  }
  if (d->cfg.gram_order == G1S_GRAM_REF_ORDER) {
    // strict mode: the same sums once more, term by term in the reference's order and rounding (f64 chains)
    CU_TRY(d, cudaEventRecord(s.ks_beg, st));
    launch_gram_strict(s.d_descs, s.count, d->geom, s.d_records, d->rl, st);
    CU_TRY(d, cudaEventRecord(s.ks_end, st));
    gram_launches += 1;
  }
  if (d->device_model) {
    launch_latest(s.count, d->geom, s.d_records, d->rl, d->cfg.gram_order == G1S_GRAM_REF_ORDER, s.d_digests,
                  (int)LatestFrame::kDigestDoubles, st);
    gram_launches += 1;
  }
  CU_TRY(d, cudaGetLastError());
  // the read-back rides its own stream so the next batch's kernels start right behind this batch's; with the
  // per-frame model on the device only the digests (11 KB per frame) come back, the records only for a record tap
  CU_TRY(d, cudaEventRecord(s.computed, st));
  CU_TRY(d, cudaStreamWaitEvent(d->d2h_stream, s.computed, 0));
  s.records_read = !d->device_model || d->tap != nullptr;
  if (d->device_model)
    CU_TRY(d, cudaMemcpyAsync(s.h_digests, s.d_digests, sizeof(double) * LatestFrame::kDigestDoubles * s.count,
                              cudaMemcpyDeviceToHost, d->d2h_stream));
  if (s.records_read)
    CU_TRY(d, cudaMemcpyAsync(s.h_records, s.d_records, d->rl.bytes * s.count, cudaMemcpyDeviceToHost, d->d2h_stream));
  CU_TRY(d, cudaEventRecord(s.done, d->d2h_stream));
  s.in_flight = true;
  d->kernels_launched += 2 + gram_launches;
  d->k0_launches += 1;
  d->k1_launches += 1;
  return G1S_OK;
}

// Waits for a slot's device work and folds its records into the model (frame order).
int retire(g1s_diff *d, Slot &s) {
  if (!s.in_flight) return G1S_OK;
  CU_TRY(d, cudaEventSynchronize(s.done));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, s.k0_beg, s.k0_end) == cudaSuccess) d->k0_ms += ms;
  if (cudaEventElapsedTime(&ms, s.k1_beg, s.k1_end) == cudaSuccess) d->k1_ms += ms;
  if (cudaEventElapsedTime(&ms, s.kr_beg, d->tensor_path ? s.k0_beg : s.k1_beg) == cudaSuccess) d->kr_ms += ms;
  if (d->cfg.gram_order == G1S_GRAM_REF_ORDER && cudaEventElapsedTime(&ms, s.ks_beg, s.ks_end) == cudaSuccess)
    d->ks_ms += ms;
  d->folder->wait(s.fold_ticket);  // the slot's previous batch must have left the fold thread
  s.fold_ticket = fold_records(d, s.h_records, s.count, d->rl.bytes, s.latest, d->device_model ? s.h_digests : nullptr,
                               s.records_read);
  d->frames_done += s.count;
  s.in_flight = false;
  s.count = 0;
  s.host_frames = 0;
  return G1S_OK;
}

int rotate(g1s_diff *d) {
  int rc = submit(d, d->slots[d->cur]);
  if (rc != G1S_OK) return rc;
  d->cur = (d->cur + 1) % kSlots;
  // the slot we are about to fill must be free; retire in submission order
  while (d->slots[d->cur].in_flight) {
    rc = retire(d, d->slots[d->oldest]);
    if (rc != G1S_OK) return rc;
    d->oldest = (d->oldest + 1) % kSlots;
  }
  return G1S_OK;
}

int drain(g1s_diff *d) {
  int rc = submit(d, d->slots[d->cur]);
  if (rc != G1S_OK) return rc;
  if (d->slots[d->cur].in_flight) d->cur = (d->cur + 1) % kSlots;
  for (int k = 0; k < kSlots; ++k) {
    rc = retire(d, d->slots[d->oldest]);
    if (rc != G1S_OK) return rc;
    d->oldest = (d->oldest + 1) % kSlots;
  }
  d->oldest = d->cur;
  d->folder->wait_all();
  return G1S_OK;
}

int check_frames(g1s_diff *d, const g1s_frame *s, const g1s_frame *n) {
  if (!s || !n) {
    d->err = "null frame";
    return G1S_E_ARG;
  }
  if (d->filters || !d->filter_ops.empty()) {
    // the source is filtered before the comparison: it must have the size the chain was configured for, the
    // denoised frame the size the chain produces (= the handle's)
    if (s->width != d->raw_w || s->height != d->raw_h || n->width != d->cfg.width || n->height != d->cfg.height) {
      char b[200];
      std::snprintf(b, sizeof b, "Luma dimensions do not match: filtered source %dx%d (from %dx%d), denoised %dx%d",
                    d->cfg.width, d->cfg.height, s->width, s->height, n->width, n->height);
      d->err = b;
      return G1S_E_DIMS;
    }
  } else if (s->width != n->width || s->height != n->height) {
    char b[160];
    std::snprintf(b, sizeof b, "Luma dimensions do not match: source %dx%d, denoised %dx%d", s->width, s->height,
                  n->width, n->height);
    d->err = b;
    return G1S_E_DIMS;
  }
  if (n->width != d->cfg.width || n->height != d->cfg.height) {
    d->err = "frame size differs from the size the handle was created for";
    return G1S_E_ARG;
  }
  for (int c = 0; c < d->geom.planes; ++c)
    if (!s->plane[c] || !n->plane[c]) {
      d->err = "missing plane";
      return G1S_E_ARG;
    }
  return G1S_OK;
}


// Queues `count` digests (copied) for the fold thread, in order.  The fold thread unpacks a block into model states
// with its helpers, then merges them (DiffSequencer::consume_latest_batch).
void fold_digests_copy(g1s_diff *d, const double *src, size_t count) {
  std::shared_ptr<DigestBlock> blk;
  {
    std::lock_guard<std::mutex> lk(d->digest_mu);
    if (!d->digest_free.empty()) {
      blk = std::move(d->digest_free.back());
      d->digest_free.pop_back();
    }
  }
  if (!blk) blk = std::make_shared<DigestBlock>();
  blk->raw.assign(src, src + LatestFrame::kDigestDoubles * count);
  DiffSequencer *seq = d->seq.get();
  const NoiseModel::ParallelFor *par = &d->fold_par;
  d->folder->push([blk, seq, count, d, par] {
    constexpr size_t K = LatestFrame::kDigestDoubles;
    constexpr size_t kStep = 256;  // states held unpacked at a time (25 KB each)
    if (blk->frames.size() < std::min(count, kStep)) blk->frames.resize(std::min(count, kStep));
    LatestFrame *frames = blk->frames.data();
    const double *raw = blk->raw.data();
    for (size_t at = 0; at < count; at += kStep) {
      const int n = (int)std::min(kStep, count - at);
      (*par)(n, [&](int i) { frames[i].from_digest(raw + K * (at + i)); });
      seq->consume_latest_batch(frames, n, *par);
    }
    std::lock_guard<std::mutex> lk(d->digest_mu);
    if (d->digest_free.size() < 8) d->digest_free.push_back(blk);
  });
  d->retired += (int64_t)count;
  d->frames_done += (double)count;
}

// ---- multi-device handles -------------------------------------------------------------------------------------
// Frames are dealt to the children in batches, round-robin: frame k belongs to child (k / batch) % n.  Every child is
// a PRODUCER handle on its own device (kernels + the per-frame half of the model) that writes one digest per frame
// into its pinned ring; this handle folds the digests in global frame order.  Same table as one device.
int multi_fail(g1s_diff *d, int kid, int rc) {
  d->err = "device " + std::to_string(d->cfg.device_ids[kid]) + ": " + g1s_diff_last_error(d->kids[kid]);
  return (rc == G1S_E_CUDA && kid > 0) ? G1S_E_NCCL : rc;
}

// Folds every global batch whose digests are complete; with `final` the (possibly short) remaining batches too.
void multi_collect(g1s_diff *d, bool final) {
  const int n = (int)d->kids.size();
  const int64_t B = d->batch;
  for (;;) {
    const int kid = (int)(d->next_batch % n);
    const int64_t j = d->next_batch / n;  // the child's own batch index
    const int64_t have = g1s_diff_digest_count(d->kids[kid]);
    int64_t cnt = std::min<int64_t>(B, d->kid_frames[kid] - j * B);
    if (cnt <= 0) break;                       // nothing dealt to this batch yet
    if (cnt < B && !final) break;              // the batch is still being filled
    if (have < j * B + cnt) break;             // its kernels / model half have not retired yet
    const size_t at = (size_t)((j * B) % (int64_t)d->kid_sink_cap);
    fold_digests_copy(d, d->kid_sinks[kid] + at * LatestFrame::kDigestDoubles, (size_t)cnt);
    d->next_batch++;
  }
}

int multi_push(g1s_diff *d, const g1s_frame *source, const g1s_frame *denoised, bool device_frames) {
  const int n = (int)d->kids.size();
  const int kid = (int)((d->pushed / d->batch) % n);
  if (cudaSetDevice(d->cfg.device_ids[kid]) != cudaSuccess) {
    d->err = "cudaSetDevice failed for a device of this handle";
    return kid > 0 ? G1S_E_NCCL : G1S_E_CUDA;
  }
  const int rc = device_frames ? g1s_diff_push_frame_device(d->kids[kid], source, denoised)
                               : g1s_diff_push_frame(d->kids[kid], source, denoised);
  if (rc != G1S_OK) return multi_fail(d, kid, rc);
  d->kid_frames[kid]++;
  d->pushed++;
  multi_collect(d, false);
  return G1S_OK;
}

int multi_drain(g1s_diff *d) {
  for (size_t k = 0; k < d->kids.size(); ++k) {
    if (cudaSetDevice(d->cfg.device_ids[k]) != cudaSuccess) return k > 0 ? G1S_E_NCCL : G1S_E_CUDA;
    const int rc = g1s_diff_flush(d->kids[k]);
    if (rc != G1S_OK) return multi_fail(d, (int)k, rc);
  }
  multi_collect(d, true);
  d->folder->wait_all();
  return G1S_OK;
}

}  // namespace

extern "C" {

int g1s_abi_version(void) { return G1S_ABI_VERSION; }

int g1s_diff_create(const g1s_diff_config *cfg, g1s_diff **out) {
  if (!cfg || !out) {
    g_create_error = "null argument";
    return G1S_E_ARG;
  }
  *out = nullptr;
  if (cfg->width <= 0 || cfg->height <= 0 || cfg->fps_num <= 0 || cfg->fps_den <= 0 || cfg->src_bit_depth < 8 ||
      cfg->src_bit_depth > 16 || cfg->den_bit_depth < 8 || cfg->den_bit_depth > 16 || cfg->ss_x < 0 ||
      cfg->ss_x > 1 || cfg->ss_y < 0 || cfg->ss_y > 1) {
    g_create_error = "unsupported configuration (bit depths 8..16, subsampling log2 0..1, positive size and fps)";
    return G1S_E_ARG;
  }
  if (cfg->mode < G1S_MODE_FULL || cfg->mode > G1S_MODE_CONSUMER) {
    g_create_error = "unknown mode";
    return G1S_E_ARG;
  }
  if (cfg->gram_order != G1S_GRAM_EXACT_INT && cfg->gram_order != G1S_GRAM_REF_ORDER) {
    g_create_error = "unknown gram_order";
    return G1S_E_ARG;
  }
  const bool multi = cfg->n_devices >= 2;
  if (multi && (cfg->n_devices > G1S_MAX_DEVICES || cfg->mode != G1S_MODE_FULL)) {
    g_create_error = "multi-device handles: 2..8 devices, mode G1S_MODE_FULL";
    return G1S_E_ARG;
  }
  const bool consumer = cfg->mode == G1S_MODE_CONSUMER || multi;  // a multi-device handle itself only owns the model
  int ndev = 0;
  cudaError_t ce = consumer ? cudaSuccess : cudaGetDeviceCount(&ndev);
  if (!consumer && (ce != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev)) {
    g_create_error = std::string("no usable CUDA device (this engine has no CPU fallback): ") +
                     (ce != cudaSuccess ? cudaGetErrorString(ce) : "device ordinal out of range");
    return G1S_E_CUDA;
  }
  std::unique_ptr<g1s_diff> d(new g1s_diff);
  d->cfg = *cfg;
  auto fail = [&](int code) {
    g_create_error = d->err;
    g1s_diff_destroy(d.release());
    return code;
  };
#define CU_NEW(expr)                                                     \
  do {                                                                   \
    cudaError_t e_ = (expr);                                             \
    if (e_ != cudaSuccess) {                                             \
      d->err = std::string(#expr) + ": " + cudaGetErrorString(e_);       \
      return fail(e_ == cudaErrorMemoryAllocation ? G1S_E_NOMEM : G1S_E_CUDA); \
    }                                                                    \
  } while (0)
  if (!consumer) CU_NEW(cudaSetDevice(cfg->device));

  Geometry &g = d->geom;
  g.width = cfg->width;
  g.height = cfg->height;
  g.ss_x = cfg->ss_x;
  g.ss_y = cfg->ss_y;
  g.planes = cfg->monochrome ? 1 : 3;
  g.src_shift = cfg->src_bit_depth - 8;
  g.den_shift = cfg->den_bit_depth - 8;
  g.src_bytes = cfg->src_bit_depth > 8 ? 2 : 1;
  g.den_bytes = cfg->den_bit_depth > 8 ? 2 : 1;
  d->host_bytes[0] = g.src_bytes, d->host_bytes[1] = g.den_bytes;
  d->host_shift[0] = g.src_shift, d->host_shift[1] = g.den_shift;
  d->narrow = cfg->host_narrow != 0 && !consumer && (g.src_bytes == 2 || g.den_bytes == 2);
  if (d->narrow) {  // from here on this is an 8-bit stream for the frame store and the kernels
    g.src_shift = g.den_shift = 0;
    g.src_bytes = g.den_bytes = 1;
  }
  g.nbw = (g.width + kBlock - 1) / kBlock;
  g.nbh = (g.height + kBlock - 1) / kBlock;
  g.nb = g.nbw * g.nbh;
  d->sgeom.width = g.width;
  d->sgeom.height = g.height;
  d->sgeom.ss_x = g.ss_x;
  d->sgeom.ss_y = g.ss_y;
  d->sgeom.planes = g.planes;
  d->sgeom.nbw = g.nbw;
  d->sgeom.nbh = g.nbh;
  d->sgeom.nb = g.nb;
  d->rl = RecordLayout::make(g.nb);
  flat_block_ata_inv(d->fc.ata_inv);
  d->seq.reset(new DiffSequencer(cfg->fps_num, cfg->fps_den, d->sgeom));
  {
    int threads = cfg->host_threads;
    if (const char *e = std::getenv("G1S_HOST_THREADS")) threads = std::atoi(e);
    if (threads <= 0) threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));  // staging copies want them all
    d->pool.reset(new HostPool(threads));
    d->folder.reset(new FoldQueue());
    // the sequential half of the model hands the solves of the combined state to a few helpers of its own
    int helpers = (int)std::min(6u, std::max(1u, std::thread::hardware_concurrency() / 4));
    if (const char *e = std::getenv("G1S_FOLD_THREADS")) helpers = std::max(1, std::atoi(e));
    d->fold_pool.reset(new HostPool(helpers));
    HostPool *fp = d->fold_pool.get();
    d->fold_par = [fp](int n, const std::function<void(int)> &fn) { fp->parallel_for(n, fn); };
  }

  size_t off = 0;
  for (int c = 0; c < g.planes; ++c) {
    PlaneGeom &p = d->pg[c];
    p.w = c ? (g.width + g.ss_x) >> g.ss_x : g.width;
    p.h = c ? (g.height + g.ss_y) >> g.ss_y : g.height;
    const int bytes[2] = {g.src_bytes, g.den_bytes};
    for (int k = 0; k < 2; ++k) {
      p.pitch[k] = align_up((size_t)p.w * bytes[k], 16);
      p.off[k] = off;
      off += align_up(p.pitch[k] * p.h, 256);
    }
  }
  d->pair_bytes = off;
  // Default batch: at most ~2 GiB of frame pairs per slot and 64 frames (measured at 4K 10-bit: 15.9 / 20.9 / 21.2 / 21.4 k
  // frames/s with 10 / 20 / 30 / 40 frames per launch; the per-frame model kernel, one CTA per frame, gains most);
  // within that, the size whose flat-block launch (one thread per block, 384 resident threads per SM) fills its last wave best.
  int batch = cfg->batch_frames;
  if (batch <= 0) {
    const int cap = (int)std::max<size_t>(1, std::min<size_t>(64, ((size_t)2 << 30) / d->pair_bytes));
    int sms = 148;
    if (!consumer) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    const double wave = 384.0 * sms;
    double best = -1;
    batch = cap;
    for (int b = std::max(1, cap / 2); b <= cap; ++b) {
      const double waves = (double)b * g.nb / wave, eff = waves / std::ceil(waves);
      if (eff >= best) best = eff, batch = b;
    }
  }
  d->batch = batch = std::min(batch, 64);  // the Gram kernel's warps keep a batch's unit counts in two registers per lane

  if (multi) {
    // one PRODUCER child per device; their batch size is this handle's dealing unit
    g1s_diff_config kc = *cfg;
    kc.n_devices = 0;
    kc.mode = G1S_MODE_PRODUCER;
    if (kc.model_placement == G1S_MODEL_AUTO)  // about 3 host cores per GPU keep the per-frame model half fed
      kc.model_placement = 3u * (unsigned)cfg->n_devices > std::max(2u, std::thread::hardware_concurrency()) / 2 ? G1S_MODEL_DEVICE
                                                                                                                : G1S_MODEL_HOST;
    for (int i = 0; i < cfg->n_devices; ++i) {
      kc.device = cfg->device_ids[i];
      g1s_diff *kid = nullptr;
      const int rc = g1s_diff_create(&kc, &kid);
      if (rc != G1S_OK) {
        d->err = "device " + std::to_string(kc.device) + ": " + g_create_error;
        return fail((rc == G1S_E_CUDA && i > 0) ? G1S_E_NCCL : rc);
      }
      d->kids.push_back(kid);
      d->batch = kid->batch;
      d->kid_sink_cap = (size_t)kid->batch * 8;
      double *ring = nullptr;
      CU_NEW(cudaMallocHost(&ring, d->kid_sink_cap * LatestFrame::kDigestDoubles * sizeof(double)));
      d->kid_sinks.push_back(ring);
      d->kid_frames.push_back(0);
      g1s_diff_set_digest_sink(kid, ring, d->kid_sink_cap);
    }
    *out = d.release();
    return G1S_OK;
  }
  if (consumer) {
    *out = d.release();
    return G1S_OK;
  }
  {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      d->encode_tiled = reinterpret_cast<g1s_diff::EncodeTiledFn>(fn);
    else
      (void)cudaGetLastError();
  }
  CU_NEW(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  {
    const char *e = std::getenv("G1S_STREAMS");
    d->nstreams = e ? std::min(std::max(std::atoi(e), 1), 4) : 3;  // measured: 16.1 / 17.6 / 18.5 k frames/s with 1 / 2 / 3
    for (int i = 1; i < d->nstreams; ++i) CU_NEW(cudaStreamCreateWithFlags(&d->more[i - 1], cudaStreamNonBlocking));
    CU_NEW(cudaEventCreateWithFlags(&d->join, cudaEventDisableTiming));
  }
  CU_NEW(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
  CU_NEW(cudaStreamCreateWithFlags(&d->d2h_stream, cudaStreamNonBlocking));
  // int8 tensor-core path (residual_kernel + gram_imma_kernel): 4:2:0 / monochrome, TMA descriptors available
  d->tensor_path = cfg->gram_kernel == 0 && gram_imma_supported(g) && d->encode_tiled != nullptr;
  d->rstore = ResidualStore::make(g);
  {
    // per-frame model half: host threads by default; on the device when asked, or when this process drives more GPUs
    // than the host has cores for (a multi-device parent passes its decision down as an explicit placement)
    int place = cfg->model_placement;
    if (std::getenv("G1S_HOST_MODEL")) place = G1S_MODEL_HOST;
    if (std::getenv("G1S_DEVICE_MODEL")) place = G1S_MODEL_DEVICE;
    d->device_model = place == G1S_MODEL_DEVICE && latest_supported(g);
  }
  for (Slot &s : d->slots) {
    // the frame store (device) and its pinned mirror (host) are allocated on the first host push:
    // streams whose frames are already in HBM never need them
    if (d->tensor_path) {
      CU_NEW(cudaMalloc(&s.d_res, d->rstore.frame_bytes * batch));
      CU_NEW(cudaMalloc(&s.d_rmaps, sizeof(CUtensorMap) * kResidualMaps * batch));
      CU_NEW(cudaMalloc(&s.d_plan, gram_plan_bytes(batch, g)));
      CU_NEW(cudaMalloc(&s.d_plan_counts, sizeof(int) * (3 * batch + 1)));
      std::vector<CUtensorMap> host;
      if (!build_residual_maps(d.get(), s, host)) {
        d->err = "cuTensorMapEncodeTiled failed for the residual planes";
        return fail(G1S_E_CUDA);
      }
      CU_NEW(cudaMemcpy(s.d_rmaps, host.data(), sizeof(CUtensorMap) * host.size(), cudaMemcpyHostToDevice));
    }
    CU_NEW(cudaMalloc(&s.d_descs, sizeof(FrameDesc) * batch));
    CU_NEW(cudaMallocHost(&s.h_descs, sizeof(FrameDesc) * batch));
    if (d->device_model) {
      CU_NEW(cudaMalloc(&s.d_digests, sizeof(double) * LatestFrame::kDigestDoubles * batch));
      CU_NEW(cudaMallocHost(&s.h_digests, sizeof(double) * LatestFrame::kDigestDoubles * batch));
    }
    CU_NEW(cudaMalloc(&s.d_records, d->rl.bytes * batch));
    CU_NEW(cudaMallocHost(&s.h_records, d->rl.bytes * batch));
    CU_NEW(cudaEventCreate(&s.done));
    CU_NEW(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    CU_NEW(cudaEventCreateWithFlags(&s.computed, cudaEventDisableTiming));
    CU_NEW(cudaEventCreate(&s.k0_beg));
    CU_NEW(cudaEventCreate(&s.k0_end));
    CU_NEW(cudaEventCreate(&s.kr_beg));
    CU_NEW(cudaEventCreate(&s.k1_beg));
    CU_NEW(cudaEventCreate(&s.k1_end));
    CU_NEW(cudaEventCreate(&s.ks_beg));
    CU_NEW(cudaEventCreate(&s.ks_end));
  }
#undef CU_NEW
  *out = d.release();
  return G1S_OK;
}

// util.rs::frame_into_u8 on one row: `(v >> (bit_depth - 8)) as u8`.  The destination is the pinned staging ring, which
// nothing reads before the DMA engine does: where the host has AVX-512BW the bytes go out with non-temporal stores (no
// read-for-ownership of the destination lines: a third less memory traffic on a copy that is memory bound).
extern "C" __attribute__((target_clones("avx2", "default"))) void g1s_narrow_row_plain(uint8_t *dst, const uint16_t *src,
                                                                                       int n, int shift) {
  for (int i = 0; i < n; ++i) dst[i] = (uint8_t)(src[i] >> shift);
}
#if defined(__x86_64__)
__attribute__((target("avx512f,avx512bw"))) static void narrow_row_avx512(uint8_t *dst, const uint16_t *src, int n, int shift) {
  int i = 0;
  while (i < n && ((uintptr_t)(dst + i) & 31)) dst[i] = (uint8_t)(src[i] >> shift), ++i;
  const __m128i sh = _mm_cvtsi32_si128(shift);
  for (; i + 32 <= n; i += 32) {
    const __m512i v = _mm512_srl_epi16(_mm512_loadu_si512(src + i), sh);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm512_cvtepi16_epi8(v));
  }
  for (; i < n; ++i) dst[i] = (uint8_t)(src[i] >> shift);
}
#endif
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void narrow_row_avx2(uint8_t *dst, const uint16_t *src, int n, int shift) {
  int i = 0;
  while (i < n && ((uintptr_t)(dst + i) & 31)) dst[i] = (uint8_t)(src[i] >> shift), ++i;
  const __m128i sh = _mm_cvtsi32_si128(shift);
  const __m256i low = _mm256_set1_epi16(0x00ff);  // `as u8` truncates; the pack below saturates
  for (; i + 32 <= n; i += 32) {
    const __m256i a = _mm256_and_si256(_mm256_srl_epi16(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)), sh), low);
    const __m256i b = _mm256_and_si256(_mm256_srl_epi16(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 16)), sh), low);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_permute4x64_epi64(_mm256_packus_epi16(a, b), 0xd8));
  }
  for (; i < n; ++i) dst[i] = (uint8_t)(src[i] >> shift);
}
#endif
// 2: AVX-512BW, 1: AVX2 (both with streaming stores), 0: the compiler's loop
extern "C" int g1s_narrow_isa(void) {
#if defined(__x86_64__)
  static const int isa = std::getenv("G1S_NO_STREAM_STORES") ? 0
                         : __builtin_cpu_supports("avx512bw") ? 2
                         : __builtin_cpu_supports("avx2")     ? 1
                                                              : 0;
  return isa;
#else
  return 0;
#endif
}
// (isa: a path at or below what g1s_narrow_isa() reports; the tests walk them all)
extern "C" void g1s_narrow_row_with(uint8_t *dst, const uint16_t *src, int n, int shift, int isa) {
#if defined(__x86_64__)
  isa = std::min(isa, g1s_narrow_isa());
  if (isa == 2) return narrow_row_avx512(dst, src, n, shift);
  if (isa == 1) return narrow_row_avx2(dst, src, n, shift);
#endif
  g1s_narrow_row_plain(dst, src, n, shift);
}
extern "C" void g1s_narrow_row(uint8_t *dst, const uint16_t *src, int n, int shift) {
  g1s_narrow_row_with(dst, src, n, shift, 2);
}

int g1s_diff_push_frame(g1s_diff *d, const g1s_frame *source, const g1s_frame *denoised) {
  if (!d) return G1S_E_ARG;
  if (!d->kids.empty()) {
    if (d->finished) {
      d->err = "push after finish";
      return G1S_E_STATE;
    }
    return multi_push(d, source, denoised, false);
  }
  if (d->finished || d->cfg.mode == G1S_MODE_CONSUMER) {
    d->err = d->finished ? "push after finish" : "consumer handles take records, not frames";
    return G1S_E_STATE;
  }
  int rc = check_frames(d, source, denoised);
  if (rc != G1S_OK) return rc;
  G1S_ON_DEVICE(d);
  Slot &s = d->slots[d->cur];
  // the frame store is allocated on the first host push, for every slot at once: a page-locked gigabyte takes a good
  // fraction of a second to allocate, which must not land in the middle of a stream
  if (!s.d_frames)
    for (Slot &o : d->slots)
      if (!o.d_frames) CU_TRY(d, cudaMalloc(&o.d_frames, (size_t)d->batch * d->pair_bytes));
  const size_t base = (size_t)s.count * d->pair_bytes;
  FrameDesc &fd = s.h_descs[s.count];
  std::memset(&fd, 0, sizeof fd);
  const g1s_frame *fr[2] = {source, denoised};
  const int bytes[2] = {d->host_bytes[0], d->host_bytes[1]};  // of the caller's samples

  // Planes that already live in page-locked host memory go to the device directly (one 2-D DMA per plane at
  // PCIe rate, no staging copy, no host cores); the call still returns only when the borrowed memory is no
  // longer needed.  Pageable planes take the staged path below.
  bool pinned = !d->narrow && !d->filters && std::getenv("G1S_NO_DIRECT_H2D") == nullptr;  // narrowing / filtering stage
  for (int c = 0; c < d->geom.planes && pinned; ++c)
    for (int k = 0; k < 2 && pinned; ++k) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, fr[k]->plane[c]) != cudaSuccess || at.type != cudaMemoryTypeHost) {
        (void)cudaGetLastError();
        pinned = false;
      }
    }
  if (pinned) {
    for (int c = 0; c < d->geom.planes; ++c) {
      const PlaneGeom &p = d->pg[c];
      for (int k = 0; k < 2; ++k) {
        const size_t row_bytes = (size_t)p.w * bytes[k];
        if (fr[k]->stride_bytes[c] < row_bytes) {
          d->err = "stride smaller than a row";
          return G1S_E_ARG;
        }
        uint8_t *dev = s.d_frames + base + p.off[k];
        CU_TRY(d, cudaMemcpy2DAsync(dev, p.pitch[k], fr[k]->plane[c], fr[k]->stride_bytes[c], row_bytes, p.h,
                                    cudaMemcpyHostToDevice, d->copy_stream));
        if (k == 0) {
          fd.src[c] = dev;
          fd.src_stride[c] = (uint32_t)p.pitch[k];
        } else {
          fd.den[c] = dev;
          fd.den_stride[c] = (uint32_t)p.pitch[k];
        }
      }
    }
    CU_TRY(d, cudaStreamSynchronize(d->copy_stream));  // the planes are borrowed only until we return
    s.count++;
    s.host_frames++;
    d->pushed++;
    if (s.count == d->batch) return rotate(d);
    return G1S_OK;
  }
  if (!s.h_frames)
    for (Slot &o : d->slots)
      if (!o.h_frames) CU_TRY(d, cudaMallocHost(&o.h_frames, (size_t)d->batch * d->pair_bytes));
  // The borrowed planes are copied into the pinned staging slot by the host threads, in row chunks of
  // about 1 MiB, so the copy runs at memory bandwidth rather than at one core's memcpy speed.
  struct CopyTask {
    uint8_t *dst;
    const uint8_t *src;
    size_t dst_pitch, src_pitch, row_bytes;
    int rows;
    int narrow_shift;  // >= 0: the source rows are uint16 and are reduced to uint8 with this shift while copied
    int width;
  };
  CopyTask tasks[144];
  int ntasks = 0;
  int raw_slot = -1;
  if (d->filters) {
    // the raw source goes up through its own pinned ring; the ring entry is free once its previous upload finished
    raw_slot = d->raw_next++ % g1s_diff::kRaw;
    CU_TRY(d, cudaEventSynchronize(d->raw_ev[raw_slot]));
    for (int c = 0; c < d->geom.planes; ++c) {
      const int rw = c ? (d->raw_w + d->geom.ss_x) >> d->geom.ss_x : d->raw_w;
      const int rh = c ? (d->raw_h + d->geom.ss_y) >> d->geom.ss_y : d->raw_h;
      const size_t row_bytes = (size_t)rw * bytes[0];
      if (source->stride_bytes[c] < row_bytes) {
        d->err = "stride smaller than a row";
        return G1S_E_ARG;
      }
      const int chunks = (int)std::min<size_t>(16, std::max<size_t>(1, (row_bytes * rh) >> 20));
      const int rows_per = (rh + chunks - 1) / chunks;
      for (int r0 = 0; r0 < rh; r0 += rows_per)
        tasks[ntasks++] = {d->raw_host[raw_slot] + d->raw_off[c] + (size_t)r0 * d->raw_pitch[c],
                           static_cast<const uint8_t *>(source->plane[c]) + (size_t)r0 * source->stride_bytes[c],
                           d->raw_pitch[c], source->stride_bytes[c], row_bytes, std::min(rows_per, rh - r0), -1, rw};
    }
  }
  for (int c = 0; c < d->geom.planes; ++c) {
    const PlaneGeom &p = d->pg[c];
    for (int k = 0; k < 2; ++k) {
      const size_t row_bytes = (size_t)p.w * bytes[k];
      const void *devp = s.d_frames + base + p.off[k];
      if (k == 0 && d->filters) {  // written by the filter chain, not copied
        fd.src[c] = devp;
        fd.src_stride[c] = (uint32_t)p.pitch[k];
        continue;
      }
      if (fr[k]->stride_bytes[c] < row_bytes) {
        d->err = "stride smaller than a row";
        return G1S_E_ARG;
      }
      uint8_t *dst = s.h_frames + base + p.off[k];
      const uint8_t *src = static_cast<const uint8_t *>(fr[k]->plane[c]);
      const int chunks = (int)std::min<size_t>(16, std::max<size_t>(1, (row_bytes * p.h) >> 20));
      const int rows_per = (p.h + chunks - 1) / chunks;
      for (int r0 = 0; r0 < p.h; r0 += rows_per)
        tasks[ntasks++] = {dst + (size_t)r0 * p.pitch[k], src + (size_t)r0 * fr[k]->stride_bytes[c], p.pitch[k],
                           fr[k]->stride_bytes[c], row_bytes, std::min(rows_per, p.h - r0),
                           (d->narrow && bytes[k] == 2) ? d->host_shift[k] : -1, p.w};
      const void *dev = s.d_frames + base + p.off[k];
      if (k == 0) {
        fd.src[c] = dev;
        fd.src_stride[c] = (uint32_t)p.pitch[k];
      } else {
        fd.den[c] = dev;
        fd.den_stride[c] = (uint32_t)p.pitch[k];
      }
    }
  }
  d->pool->parallel_for(ntasks, [&](int i) {
    const CopyTask &t = tasks[i];
    if (t.narrow_shift >= 0) {
      for (int y = 0; y < t.rows; ++y)
        g1s_narrow_row(t.dst + (size_t)y * t.dst_pitch, reinterpret_cast<const uint16_t *>(t.src + (size_t)y * t.src_pitch),
                       t.width, t.narrow_shift);
#if defined(__x86_64__)
      _mm_sfence();  // the streaming stores of this task are visible before the task counts as done
#endif
    } else if (t.dst_pitch == t.src_pitch) {
      std::memcpy(t.dst, t.src, t.dst_pitch * (size_t)(t.rows - 1) + t.row_bytes);
    } else {
      for (int y = 0; y < t.rows; ++y) std::memcpy(t.dst + (size_t)y * t.dst_pitch, t.src + (size_t)y * t.src_pitch, t.row_bytes);
    }
  });
  if (d->filters) {
    // denoised planes only (the source region of the staging entry holds nothing), then the raw source and the chain
    for (int c = 0; c < d->geom.planes; ++c) {
      const PlaneGeom &p = d->pg[c];
      CU_TRY(d, cudaMemcpyAsync(s.d_frames + base + p.off[1], s.h_frames + base + p.off[1], p.pitch[1] * (size_t)p.h,
                                cudaMemcpyHostToDevice, d->copy_stream));
    }
    CU_TRY(d, cudaMemcpyAsync(d->raw_dev[raw_slot], d->raw_host[raw_slot], d->raw_bytes, cudaMemcpyHostToDevice,
                              d->copy_stream));
    CU_TRY(d, cudaEventRecord(d->raw_ev[raw_slot], d->copy_stream));
    DevPlanes in, out;
    for (int c = 0; c < d->geom.planes; ++c) {
      in.ptr[c] = d->raw_dev[raw_slot] + d->raw_off[c], in.pitch[c] = d->raw_pitch[c];
      out.ptr[c] = s.d_frames + base + d->pg[c].off[0], out.pitch[c] = d->pg[c].pitch[0];
    }
    if (!d->filters->apply(in, out, d->copy_stream)) {
      d->err = d->filters->error();
      return G1S_E_CUDA;
    }
  } else {
    CU_TRY(d, cudaMemcpyAsync(s.d_frames + base, s.h_frames + base, d->pair_bytes, cudaMemcpyHostToDevice,
                              d->copy_stream));
  }
  s.count++;
  s.host_frames++;
  d->pushed++;
  if (s.count == d->batch) return rotate(d);
  return G1S_OK;
}

int g1s_diff_set_source_filters(g1s_diff *d, const g1s_filter_op *ops, size_t n, int32_t src_width, int32_t src_height) {
  if (!d || (!ops && n) || src_width <= 0 || src_height <= 0) return G1S_E_ARG;
  if (d->pushed != 0 || d->cfg.mode == G1S_MODE_CONSUMER || d->narrow) {
    d->err = "source filters are set once, before the first frame, on a handle that takes frames (and not with host_narrow)";
    return G1S_E_STATE;
  }
  if (!d->kids.empty()) {  // every device filters the frames it is dealt
    for (size_t k = 0; k < d->kids.size(); ++k) {
      cudaSetDevice(d->cfg.device_ids[k]);
      const int rc = g1s_diff_set_source_filters(d->kids[k], ops, n, src_width, src_height);
      if (rc != G1S_OK) return multi_fail(d, (int)k, rc);
    }
    d->filter_ops.assign(ops, ops + n);
    d->raw_w = src_width, d->raw_h = src_height;
    return G1S_OK;
  }
  G1S_ON_DEVICE(d);
  std::unique_ptr<SourceFilters> f(new SourceFilters);
  if (!f->configure(ops, n, src_width, src_height, d->geom.ss_x, d->geom.ss_y, d->geom.planes, d->cfg.src_bit_depth)) {
    d->err = f->error();
    return G1S_E_ARG;
  }
  if (f->out_width() != d->cfg.width || f->out_height() != d->cfg.height) {
    char b[200];
    std::snprintf(b, sizeof b, "Luma dimensions do not match: filtered source %dx%d, denoised %dx%d", f->out_width(),
                  f->out_height(), d->cfg.width, d->cfg.height);
    d->err = b;
    return G1S_E_DIMS;
  }
  d->raw_w = src_width, d->raw_h = src_height;
  size_t off = 0;
  for (int c = 0; c < d->geom.planes; ++c) {
    const int rw = c ? (src_width + d->geom.ss_x) >> d->geom.ss_x : src_width;
    const int rh = c ? (src_height + d->geom.ss_y) >> d->geom.ss_y : src_height;
    d->raw_pitch[c] = align_up((size_t)rw * d->host_bytes[0], 256);
    d->raw_off[c] = off;
    off += align_up(d->raw_pitch[c] * rh, 256);
  }
  d->raw_bytes = off;
  for (int i = 0; i < g1s_diff::kRaw; ++i) {
    CU_TRY(d, cudaMallocHost(&d->raw_host[i], d->raw_bytes));
    CU_TRY(d, cudaMalloc(&d->raw_dev[i], d->raw_bytes));
    CU_TRY(d, cudaEventCreateWithFlags(&d->raw_ev[i], cudaEventDisableTiming));
  }
  d->filter_ops.assign(ops, ops + n);
  d->filters = std::move(f);
  return G1S_OK;
}

int g1s_resize_table(int32_t alg, int32_t src, int32_t dst, int32_t *left, float *coef, int32_t max_taps) {
  if (alg < G1S_RESIZE_HERMITE || alg > G1S_RESIZE_SPLINE36 || src <= 0 || dst <= 0 || !left || !coef) return G1S_E_ARG;
  std::vector<int> l;
  std::vector<float> c;
  int taps = 0;
  build_resize_table(alg, src, dst, l, c, taps);
  if (taps > max_taps) return G1S_E_STATE;
  std::memcpy(left, l.data(), sizeof(int) * l.size());
  std::memcpy(coef, c.data(), sizeof(float) * c.size());
  return taps;
}

int g1s_diff_push_frame_device(g1s_diff *d, const g1s_frame *source, const g1s_frame *denoised) {
  if (!d) return G1S_E_ARG;
  if (!d->kids.empty()) {
    if (d->finished) {
      d->err = "push after finish";
      return G1S_E_STATE;
    }
    return multi_push(d, source, denoised, true);  // the planes must live on g1s_diff_frame_device(d, frame index)
  }
  if (d->finished || d->cfg.mode == G1S_MODE_CONSUMER) {
    d->err = d->finished ? "push after finish" : "consumer handles take records, not frames";
    return G1S_E_STATE;
  }
  if (d->narrow) {
    d->err = "host_narrow handles reduce samples while staging host frames; device frames would bypass that";
    return G1S_E_STATE;
  }
  int rc = check_frames(d, source, denoised);
  if (rc != G1S_OK) return rc;
  G1S_ON_DEVICE(d);
  Slot &s = d->slots[d->cur];
  FrameDesc &fd = s.h_descs[s.count];
  std::memset(&fd, 0, sizeof fd);
  for (int c = 0; c < d->geom.planes; ++c) {
    fd.src[c] = source->plane[c];
    fd.den[c] = denoised->plane[c];
    fd.src_stride[c] = (uint32_t)source->stride_bytes[c];
    fd.den_stride[c] = (uint32_t)denoised->stride_bytes[c];
  }
  if (d->filters) {
    // the filtered source needs a home: this slot's frame store; the denoised planes are still used in place
    if (!s.d_frames)
      for (Slot &o : d->slots)
        if (!o.d_frames) CU_TRY(d, cudaMalloc(&o.d_frames, (size_t)d->batch * d->pair_bytes));
    const size_t base = (size_t)s.count * d->pair_bytes;
    DevPlanes in, out;
    for (int c = 0; c < d->geom.planes; ++c) {
      in.ptr[c] = static_cast<uint8_t *>(const_cast<void *>(source->plane[c])), in.pitch[c] = source->stride_bytes[c];
      out.ptr[c] = s.d_frames + base + d->pg[c].off[0], out.pitch[c] = d->pg[c].pitch[0];
      fd.src[c] = out.ptr[c];
      fd.src_stride[c] = (uint32_t)out.pitch[c];
    }
    if (!d->filters->apply(in, out, d->copy_stream)) {
      d->err = d->filters->error();
      return G1S_E_CUDA;
    }
    s.host_frames++;  // submit() makes the kernels wait for the copy stream
  }
  s.count++;
  d->pushed++;
  if (s.count == d->batch) return rotate(d);
  return G1S_OK;
}

int g1s_diff_flush(g1s_diff *d) {
  if (!d) return G1S_E_ARG;
  if (!d->kids.empty()) return multi_drain(d);
  if (d->cfg.mode == G1S_MODE_CONSUMER) {
    d->folder->wait_all();
    return G1S_OK;
  }
  G1S_ON_DEVICE(d);
  return drain(d);
}

size_t g1s_record_layout(int32_t num_blocks, size_t off[8]) {
  const RecordLayout rl = RecordLayout::make(num_blocks);
  if (off) {
    const size_t v[8] = {rl.off_gram, rl.off_nobs, rl.off_num_flat, rl.off_luma_sum,
                         rl.off_rsum, rl.off_rsq,  rl.off_score,    rl.off_flat};
    std::memcpy(off, v, sizeof v);
  }
  return rl.bytes;
}

size_t g1s_record_gramf_offset(int32_t num_blocks) { return RecordLayout::make(num_blocks).off_gramf; }

size_t g1s_diff_record_bytes(const g1s_diff *d) { return d ? d->rl.bytes : 0; }

int g1s_diff_set_record_tap(g1s_diff *d, g1s_record_fn fn, void *user) {
  if (!d) return G1S_E_ARG;
  if (!d->kids.empty()) {
    d->err = "multi-device handles exchange digests, not records: no record tap";
    return G1S_E_STATE;
  }
  d->tap = fn;
  d->tap_user = user;
  return G1S_OK;
}

int g1s_diff_consume_record(g1s_diff *d, const void *record, size_t bytes) {
  if (!d || !record) return G1S_E_ARG;
  if (d->cfg.mode != G1S_MODE_CONSUMER || d->finished) {
    d->err = "consume_record needs an unfinished CONSUMER handle";
    return G1S_E_STATE;
  }
  if (bytes != d->rl.bytes) {
    d->err = "record size does not match this stream's geometry";
    return G1S_E_ARG;
  }
  auto blk = std::make_shared<std::vector<LatestFrame>>(1);
  const uint64_t t = fold_records(d, static_cast<const uint8_t *>(record), 1, bytes, *blk);
  d->folder->push([blk] {});  // keeps the block alive until the fold job before it has run
  (void)t;
  d->pushed++;
  d->frames_done += 1;
  return G1S_OK;
}

int g1s_diff_consume_records(g1s_diff *d, const void *records, size_t count, size_t stride_bytes) {
  if (!d || (!records && count)) return G1S_E_ARG;
  if (d->cfg.mode != G1S_MODE_CONSUMER || d->finished) {
    d->err = "consume_records needs an unfinished CONSUMER handle";
    return G1S_E_STATE;
  }
  if (stride_bytes < d->rl.bytes) {
    d->err = "record stride smaller than this stream's record size";
    return G1S_E_ARG;
  }
  auto blk = std::make_shared<std::vector<LatestFrame>>(count);
  fold_records(d, static_cast<const uint8_t *>(records), (int)count, stride_bytes, *blk);
  d->folder->push([blk] {});  // keeps the block alive until the fold job before it has run
  d->pushed += (int64_t)count;
  d->frames_done += (double)count;
  return G1S_OK;
}

int g1s_diff_finish(g1s_diff *d, g1s_segment *out, size_t cap, size_t *n) {
  if (!d || !n) return G1S_E_ARG;
  if (d->cfg.mode == G1S_MODE_PRODUCER) {
    d->err = "producer handles have no model; finish the CONSUMER handle";
    return G1S_E_STATE;
  }
  G1S_ON_DEVICE(d);
  int rc = !d->kids.empty() ? multi_drain(d) : (d->cfg.mode == G1S_MODE_CONSUMER ? G1S_OK : drain(d));
  if (rc != G1S_OK) return rc;
  d->folder->wait_all();
  std::vector<g1s_segment> segs = d->seq->finish();
  *n = segs.size();
  if (cap < segs.size() || !out) {
    d->err = "segment capacity too small";
    return G1S_E_STATE;
  }
  std::memcpy(out, segs.data(), segs.size() * sizeof(g1s_segment));
  d->finished = true;
  return G1S_OK;
}

void g1s_diff_destroy(g1s_diff *d) {
  if (!d) return;
  for (size_t k = 0; k < d->kids.size(); ++k) {
    cudaSetDevice(d->cfg.device_ids[k]);
    g1s_diff_destroy(d->kids[k]);
  }
  for (double *r : d->kid_sinks)
    if (r) cudaFreeHost(r);
  d->folder.reset();  // runs the queued folds to completion, then joins
  if (owns_device(d)) cudaSetDevice(d->cfg.device);
  if (d->copy_stream) cudaStreamSynchronize(d->copy_stream);
  if (d->d2h_stream) cudaStreamSynchronize(d->d2h_stream);
  if (d->stream) cudaStreamSynchronize(d->stream);
  for (cudaStream_t m : d->more)
    if (m) cudaStreamSynchronize(m);
  for (Slot &s : d->slots) {
    if (s.d_frames) cudaFree(s.d_frames);
    if (s.h_frames) cudaFreeHost(s.h_frames);
    if (s.d_res) cudaFree(s.d_res);
    if (s.d_rmaps) cudaFree(s.d_rmaps);
    if (s.d_plan) cudaFree(s.d_plan);
    if (s.d_plan_counts) cudaFree(s.d_plan_counts);
    if (s.d_descs) cudaFree(s.d_descs);
    if (s.h_descs) cudaFreeHost(s.h_descs);
    if (s.d_digests) cudaFree(s.d_digests);
    if (s.h_digests) cudaFreeHost(s.h_digests);
    if (s.d_records) cudaFree(s.d_records);
    if (s.h_records) cudaFreeHost(s.h_records);
    for (cudaEvent_t e : {s.done, s.k0_beg, s.k0_end, s.kr_beg, s.k1_beg, s.k1_end, s.ks_beg, s.ks_end, s.copied, s.computed})
      if (e) cudaEventDestroy(e);
  }
  for (int i = 0; i < g1s_diff::kRaw; ++i) {
    if (d->raw_host[i]) cudaFreeHost(d->raw_host[i]);
    if (d->raw_dev[i]) cudaFree(d->raw_dev[i]);
    if (d->raw_ev[i]) cudaEventDestroy(d->raw_ev[i]);
  }
  d->filters.reset();
  for (cudaEvent_t e : d->marks)
    if (e) cudaEventDestroy(e);
  if (d->join) cudaEventDestroy(d->join);
  if (d->stream) cudaStreamDestroy(d->stream);
  for (cudaStream_t m : d->more)
    if (m) cudaStreamDestroy(m);
  if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
  if (d->d2h_stream) cudaStreamDestroy(d->d2h_stream);
  delete d;
}

const char *g1s_diff_last_error(const g1s_diff *d) { return d ? d->err.c_str() : g_create_error.c_str(); }

int64_t g1s_diff_frames_pushed(const g1s_diff *d) { return d ? d->pushed : 0; }

int g1s_diff_batch_frames(const g1s_diff *d) { return d ? d->batch : 0; }
int g1s_diff_model_on_device(const g1s_diff *d) {
  if (!d) return 0;
  return d->kids.empty() ? (d->device_model ? 1 : 0) : g1s_diff_model_on_device(d->kids[0]);
}

int g1s_diff_frame_device(const g1s_diff *d, int64_t frame_index) {
  if (!d || frame_index < 0) return -1;
  if (d->kids.empty()) return d->cfg.device;
  return d->cfg.device_ids[(frame_index / d->batch) % (int64_t)d->kids.size()];
}

int g1s_diff_get_counters(const g1s_diff *d, double *out, size_t n) {
  if (!d || !out) return G1S_E_ARG;
  if (!d->kids.empty()) {  // sums over the devices (frames_done: frames folded into the model)
    double acc[10] = {0}, one[10];
    for (g1s_diff *k : d->kids) {
      g1s_diff_get_counters(k, one, 10);
      for (int i = 0; i < 10; ++i) acc[i] += one[i];
    }
    acc[5] = d->frames_done;
    for (size_t i = 0; i < n && i < 10; ++i) out[i] = acc[i];
    return G1S_OK;
  }
  const double v[10] = {d->kernels_launched, d->k1_ms,       d->k1_launches, d->k0_ms,          d->k0_launches,
                        d->frames_done,      d->tma_batches, d->kr_ms,       d->vector_batches, d->ks_ms};
  for (size_t i = 0; i < n && i < 10; ++i) out[i] = v[i];
  return G1S_OK;
}

int g1s_diff_mark(g1s_diff *d, int which) {
  if (d && !d->kids.empty()) {
    for (size_t k = 0; k < d->kids.size(); ++k) {
      cudaSetDevice(d->cfg.device_ids[k]);
      const int rc = g1s_diff_mark(d->kids[k], which);
      if (rc != G1S_OK) return rc;
    }
    return G1S_OK;
  }
  if (!d || which < 0 || which > 1 || !d->stream) return G1S_E_ARG;
  G1S_ON_DEVICE(d);
  if (!d->marks[which]) CU_TRY(d, cudaEventCreate(&d->marks[which]));
  for (cudaStream_t m : d->more)  // the mark covers every kernel stream
    if (m) {
      CU_TRY(d, cudaEventRecord(d->join, m));
      CU_TRY(d, cudaStreamWaitEvent(d->stream, d->join, 0));
    }
  CU_TRY(d, cudaEventRecord(d->marks[which], d->stream));
  return G1S_OK;
}

double g1s_diff_marks_elapsed_ms(g1s_diff *d) {
  if (d && !d->kids.empty()) {  // the slowest device
    double worst = -1.0;
    for (size_t k = 0; k < d->kids.size(); ++k) {
      cudaSetDevice(d->cfg.device_ids[k]);
      worst = std::max(worst, g1s_diff_marks_elapsed_ms(d->kids[k]));
    }
    return worst;
  }
  if (!d || !d->marks[0] || !d->marks[1]) return -1.0;
  float ms = -1.0f;
  if (cudaEventSynchronize(d->marks[1]) != cudaSuccess || cudaEventElapsedTime(&ms, d->marks[0], d->marks[1]) != cudaSuccess)
    return -1.0;
  return (double)ms;
}

size_t g1s_digest_bytes(void) { return LatestFrame::kDigestDoubles * sizeof(double); }

int g1s_diff_set_digest_sink(g1s_diff *d, void *buffer, size_t capacity_frames) {
  if (!d) return G1S_E_ARG;
  d->sink = static_cast<double *>(buffer);
  d->sink_cap = buffer ? capacity_frames : 0;
  d->sink_count = 0;
  return G1S_OK;
}

int64_t g1s_diff_digest_count(const g1s_diff *d) { return d ? (int64_t)d->sink_count : 0; }

int g1s_diff_wait_retired(g1s_diff *d, int64_t frames) {
  if (!d) return G1S_E_ARG;
  if (d->cfg.mode == G1S_MODE_CONSUMER) return G1S_OK;
  G1S_ON_DEVICE(d);
  while (d->retired < frames && d->slots[d->oldest].in_flight) {
    const int rc = retire(d, d->slots[d->oldest]);
    if (rc != G1S_OK) return rc;
    d->oldest = (d->oldest + 1) % kSlots;
  }
  return G1S_OK;
}

int g1s_diff_consume_digests(g1s_diff *d, const void *digests, size_t count) {
  if (!d || (!digests && count)) return G1S_E_ARG;
  if (d->cfg.mode != G1S_MODE_CONSUMER || d->finished) {
    d->err = "consume_digests needs an unfinished CONSUMER handle";
    return G1S_E_STATE;
  }
  // the digests are copied (the caller's buffer is free on return) and folded asynchronously, in order; the
  // copies live in recycled buffers (a fresh 10+ MB allocation per exchange is mostly page faults)
  fold_digests_copy(d, static_cast<const double *>(digests), count);
  d->pushed += (int64_t)count;
  return G1S_OK;
}

int g1s_diff_consume_digests_borrowed(g1s_diff *d, const void *digests, size_t count) {
  if (!d || (!digests && count)) return G1S_E_ARG;
  if (d->cfg.mode != G1S_MODE_CONSUMER || d->finished) {
    d->err = "consume_digests needs an unfinished CONSUMER handle";
    return G1S_E_STATE;
  }
  const double *src = static_cast<const double *>(digests);
  DiffSequencer *seq = d->seq.get();
  const NoiseModel::ParallelFor *par = &d->fold_par;
  d->folder->push([src, seq, count, par] {
    constexpr size_t K = LatestFrame::kDigestDoubles;
    constexpr size_t kStep = 256;
    static thread_local std::vector<LatestFrame> frames;  // the fold thread's
    if (frames.size() < std::min(count, kStep)) frames.resize(std::min(count, kStep));
    LatestFrame *fr = frames.data();
    for (size_t at = 0; at < count; at += kStep) {
      const int n = (int)std::min(kStep, count - at);
      (*par)(n, [&](int i) { fr[i].from_digest(src + K * (at + i)); });
      seq->consume_latest_batch(fr, n, *par);
    }
  });
  d->retired += (int64_t)count;
  d->pushed += (int64_t)count;
  d->frames_done += (double)count;
  return G1S_OK;
}

int g1s_diff_digest_from_record(g1s_diff *d, const void *record, size_t bytes, void *digest_out) {
  if (!d || !record || !digest_out) return G1S_E_ARG;
  if (bytes != d->rl.bytes) {
    d->err = "record size does not match this stream's geometry";
    return G1S_E_ARG;
  }
  LatestFrame lf;
  d->seq->model().compute_latest(view_of(d, static_cast<const uint8_t *>(record)), lf);
  lf.to_digest(static_cast<double *>(digest_out));
  return G1S_OK;
}

int64_t g1s_format_grain_table(const g1s_segment *segs, size_t n, char *buf, size_t cap) {
  if (!segs && n) return G1S_E_ARG;
  const std::string s = format_grain_table(segs, n);
  if (buf && cap) {
    const size_t k = std::min(cap - 1, s.size());
    std::memcpy(buf, s.data(), k);
    buf[k] = 0;
  }
  return (int64_t)s.size();
}

int g1s_write_grain_table(const g1s_segment *segs, size_t n, const char *path) {
  if ((!segs && n) || !path) return G1S_E_ARG;
  const std::string s = format_grain_table(segs, n);
  FILE *f = std::fopen(path, "wb");
  if (!f) return G1S_E_IO;
  const bool ok = std::fwrite(s.data(), 1, s.size(), f) == s.size();
  return (std::fclose(f) == 0 && ok) ? G1S_OK : G1S_E_IO;
}

}  // extern "C"
