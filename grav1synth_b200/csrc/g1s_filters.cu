// Source-frame filters of the `diff` command on the device: crop and resize, applied to the SOURCE frame only,
// before the diff (/root/reference/src/main.rs:621-624 -> FilterChain::apply, src/filters.rs:112-182, which calls
// video_resize::{crop, resize}).
//
// Parity status: UNPINNED.  The arithmetic lives in the un-vendored crate video-resize 0.2.0 (Cargo.lock), absent
// from this box like av1-grain, and no executable relative of it exists here.  What is restated is the published
// algorithm that crate ports (zimg's separable resampler): the five kernels the reference names -- hermite,
// catmullrom, mitchell (bicubic b/c = 0/0, 0/0.5, 1/3/1/3, support 2), lanczos (3 taps a side), spline36 (support 3)
// -- and the filter construction  scale = dst/src, step = min(scale, 1), support/step taps centred on
// (i + 0.5)/scale, positions mirrored at the frame edge, weights normalised to 1.  Choices the crate's source would
// settle and that are made here explicitly (tests/test_filters_gpu.py pins the device to oracle/resize_oracle.py, a
// numpy statement of exactly these rules):
//   * weights are computed in f64 on the host and rounded to f32; a sample is  acc = RN(acc + RN(w * x))  in f32 over
//     the taps in order, then floor(acc + 0.5) clamped to [0, 2^bit_depth - 1];
//   * horizontal pass first, rounded to the sample type, then the vertical pass;
//   * chroma planes are resized like luma to ((w + ss_x) >> ss_x, (h + ss_y) >> ss_y) with no siting shift.
// crop is exact by construction: a pointer / pitch offset on the device planes (no copy, no kernel).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "g1s_filters.h"

namespace g1s {

namespace {

double poly3(double x, double c0, double c1, double c2, double c3) { return c0 + x * (c1 + x * (c2 + x * c3)); }

double bicubic(double x, double b, double c) {
  x = std::fabs(x);
  if (x < 1.0)
    return poly3(x, (6.0 - 2.0 * b) / 6.0, 0.0, (-18.0 + 12.0 * b + 6.0 * c) / 6.0, (12.0 - 9.0 * b - 6.0 * c) / 6.0);
  if (x < 2.0)
    return poly3(x, (8.0 * b + 24.0 * c) / 6.0, (-12.0 * b - 48.0 * c) / 6.0, (6.0 * b + 30.0 * c) / 6.0,
                 (-b - 6.0 * c) / 6.0);
  return 0.0;
}

double sinc(double x) {
  const double pi = 3.14159265358979323846;
  return x == 0.0 ? 1.0 : std::sin(pi * x) / (pi * x);
}

double kernel_value(int alg, double x) {
  switch (alg) {
    case G1S_RESIZE_HERMITE: return bicubic(x, 0.0, 0.0);
    case G1S_RESIZE_CATMULLROM: return bicubic(x, 0.0, 0.5);
    case G1S_RESIZE_MITCHELL: return bicubic(x, 1.0 / 3.0, 1.0 / 3.0);
    case G1S_RESIZE_LANCZOS: {
      x = std::fabs(x);
      return x < 3.0 ? sinc(x) * sinc(x / 3.0) : 0.0;
    }
    default: {  // spline36
      x = std::fabs(x);
      if (x < 1.0) return poly3(x, 1.0, -3.0 / 209.0, -453.0 / 209.0, 13.0 / 11.0);
      if (x < 2.0) return poly3(x - 1.0, 0.0, -156.0 / 209.0, 270.0 / 209.0, -6.0 / 11.0);
      if (x < 3.0) return poly3(x - 2.0, 0.0, 26.0 / 209.0, -45.0 / 209.0, 1.0 / 11.0);
      return 0.0;
    }
  }
}

int kernel_support(int alg) { return (alg == G1S_RESIZE_LANCZOS || alg == G1S_RESIZE_SPLINE36) ? 3 : 2; }

template <typename T>
__global__ void resize_h_kernel(const T *__restrict__ in, size_t in_pitch, T *__restrict__ out, size_t out_pitch, int rows,
                                int dst_w, const int *__restrict__ left, const float *__restrict__ coef, int taps,
                                float maxv) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dst_w || y >= rows) return;
  const T *row = reinterpret_cast<const T *>(reinterpret_cast<const uint8_t *>(in) + (size_t)y * in_pitch) + left[x];
  const float *c = coef + (size_t)x * taps;
  float acc = 0.0f;
  for (int j = 0; j < taps; ++j) acc = __fadd_rn(acc, __fmul_rn(c[j], (float)row[j]));
  const float r = fminf(fmaxf(floorf(__fadd_rn(acc, 0.5f)), 0.0f), maxv);
  reinterpret_cast<T *>(reinterpret_cast<uint8_t *>(out) + (size_t)y * out_pitch)[x] = (T)r;
}

template <typename T>
__global__ void resize_v_kernel(const T *__restrict__ in, size_t in_pitch, T *__restrict__ out, size_t out_pitch, int w,
                                int dst_h, const int *__restrict__ left, const float *__restrict__ coef, int taps,
                                float maxv) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= dst_h) return;
  const uint8_t *col = reinterpret_cast<const uint8_t *>(in) + (size_t)left[y] * in_pitch;
  const float *c = coef + (size_t)y * taps;
  float acc = 0.0f;
  for (int j = 0; j < taps; ++j)
    acc = __fadd_rn(acc, __fmul_rn(c[j], (float)reinterpret_cast<const T *>(col + (size_t)j * in_pitch)[x]));
  const float r = fminf(fmaxf(floorf(__fadd_rn(acc, 0.5f)), 0.0f), maxv);
  reinterpret_cast<T *>(reinterpret_cast<uint8_t *>(out) + (size_t)y * out_pitch)[x] = (T)r;
}

}  // namespace

// zimg's compute_filter: one row of weights per output sample, folded onto the source samples they fall on.
void build_resize_table(int alg, int src, int dst, std::vector<int> &left, std::vector<float> &coef, int &taps) {
  const double scale = (double)dst / (double)src;
  const double step = std::min(scale, 1.0);
  const double support = (double)kernel_support(alg) / step;
  const int fsize = std::max((int)std::ceil(support) * 2, 1);
  taps = std::min(fsize, src);
  left.assign(dst, 0);
  coef.assign((size_t)dst * taps, 0.0f);
  std::vector<int> idx(fsize);
  std::vector<double> wt(fsize), row(taps);
  for (int i = 0; i < dst; ++i) {
    const double pos = (i + 0.5) / scale;
    const double begin = std::floor(pos - fsize / 2.0 + 0.5) + 0.5;  // round_halfup(pos - fsize / 2) + 0.5
    double total = 0.0;
    for (int j = 0; j < fsize; ++j) total += kernel_value(alg, (begin + j - pos) * step);
    int lo = src;
    for (int j = 0; j < fsize; ++j) {
      const double xpos = begin + j;
      double real = xpos < 0.0 ? -xpos : (xpos >= src ? 2.0 * src - xpos : xpos);  // mirror at the edges
      real = std::min(std::max(real, 0.0), std::nextafter((double)src, 0.0));     // clamp what is still outside
      idx[j] = (int)std::floor(real);
      wt[j] = kernel_value(alg, (xpos - pos) * step) / total;
      lo = std::min(lo, idx[j]);
    }
    const int l = std::max(std::min(lo, src - taps), 0);
    std::fill(row.begin(), row.end(), 0.0);
    for (int j = 0; j < fsize; ++j) row[std::min(idx[j] - l, taps - 1)] += wt[j];  // in order, like the dense matrix row
    left[i] = l;
    for (int k = 0; k < taps; ++k) coef[(size_t)i * taps + k] = (float)row[k];
  }
}

struct SourceFilters::Impl {
  struct Op {
    int kind, a, b, c, d;
    // resize tables on the device: [0] luma h, [1] luma v, [2] chroma h, [3] chroma v
    int *left[4] = {nullptr, nullptr, nullptr, nullptr};
    float *coef[4] = {nullptr, nullptr, nullptr, nullptr};
    int taps[4] = {0, 0, 0, 0};
    int in_w = 0, in_h = 0, out_w = 0, out_h = 0;  // luma sizes around this op
  };
  std::vector<Op> ops;
  int ss_x = 0, ss_y = 0, planes = 3, bytes = 1, bit_depth = 8;
  int src_w = 0, src_h = 0, out_w = 0, out_h = 0;
  uint8_t *scratch[3] = {nullptr, nullptr, nullptr};  // horizontal-pass result, and two ping-pong outputs
  size_t scratch_bytes = 0;
  std::string err;
};

SourceFilters::SourceFilters() : p_(new Impl) {}
SourceFilters::~SourceFilters() {
  for (auto &op : p_->ops)
    for (int k = 0; k < 4; ++k) {
      if (op.left[k]) cudaFree(op.left[k]);
      if (op.coef[k]) cudaFree(op.coef[k]);
    }
  for (uint8_t *s : p_->scratch)
    if (s) cudaFree(s);
  delete p_;
}
const char *SourceFilters::error() const { return p_->err.c_str(); }
int SourceFilters::out_width() const { return p_->out_w; }
int SourceFilters::out_height() const { return p_->out_h; }

bool SourceFilters::configure(const g1s_filter_op *ops, size_t n, int src_w, int src_h, int ss_x, int ss_y, int planes,
                              int bit_depth) {
  Impl &p = *p_;
  p.ss_x = ss_x, p.ss_y = ss_y, p.planes = planes, p.bit_depth = bit_depth, p.bytes = bit_depth > 8 ? 2 : 1;
  p.src_w = src_w, p.src_h = src_h;
  int w = src_w, h = src_h;
  size_t biggest = 0;
  for (size_t i = 0; i < n; ++i) {
    Impl::Op op;
    op.kind = ops[i].kind, op.a = ops[i].a, op.b = ops[i].b, op.c = ops[i].c, op.d = ops[i].d;
    op.in_w = w, op.in_h = h;
    if (op.kind == G1S_FILTER_CROP) {
      if (op.a < 0 || op.b < 0 || op.c < 0 || op.d < 0 || op.a + op.b >= h || op.c + op.d >= w) {
        p.err = "crop removes the whole frame";
        return false;
      }
      if (planes == 3 && (((op.a | op.b) & ((1 << ss_y) - 1)) || ((op.c | op.d) & ((1 << ss_x) - 1)))) {
        p.err = "crop offsets must be multiples of the chroma subsampling";
        return false;
      }
      w -= op.c + op.d, h -= op.a + op.b;
    } else if (op.kind == G1S_FILTER_RESIZE) {
      if (op.a <= 0 || op.b <= 0 || op.c < G1S_RESIZE_HERMITE || op.c > G1S_RESIZE_SPLINE36) {
        p.err = "Both width and height must be provided to resize filter";
        return false;
      }
      const int dims[4][2] = {{w, op.a}, {h, op.b}, {(w + ss_x) >> ss_x, (op.a + ss_x) >> ss_x},
                              {(h + ss_y) >> ss_y, (op.b + ss_y) >> ss_y}};
      for (int k = 0; k < (planes == 3 ? 4 : 2); ++k) {
        std::vector<int> left;
        std::vector<float> coef;
        build_resize_table(op.c, dims[k][0], dims[k][1], left, coef, op.taps[k]);
        if (cudaMalloc(&op.left[k], left.size() * sizeof(int)) != cudaSuccess ||
            cudaMalloc(&op.coef[k], coef.size() * sizeof(float)) != cudaSuccess) {
          p.err = "cudaMalloc failed for the resize tables";
          p.ops.push_back(op);
          return false;
        }
        cudaMemcpy(op.left[k], left.data(), left.size() * sizeof(int), cudaMemcpyHostToDevice);
        cudaMemcpy(op.coef[k], coef.data(), coef.size() * sizeof(float), cudaMemcpyHostToDevice);
      }
      biggest = std::max(biggest, (size_t)std::max(w, op.a) * std::max(h, op.b));
      w = op.a, h = op.b;
    } else {
      p.err = "unknown filter kind";
      return false;
    }
    op.out_w = w, op.out_h = h;
    p.ops.push_back(op);
  }
  p.out_w = w, p.out_h = h;
  if (biggest) {
    // one plane set (Y + 2 chroma <= 3 luma planes) per scratch buffer, 256-byte aligned rows
    p.scratch_bytes = (biggest + 4096) * 3 * p.bytes + (1 << 20);
    for (auto &s : p.scratch)
      if (cudaMalloc(&s, p.scratch_bytes) != cudaSuccess) {
        p.err = "cudaMalloc failed for the resize scratch";
        return false;
      }
  }
  return true;
}

bool SourceFilters::apply(const DevPlanes &src, const DevPlanes &dst, cudaStream_t st) {
  Impl &p = *p_;
  DevPlanes cur = src;
  int w = p.src_w, h = p.src_h, pp = 0;  // current luma size, ping-pong index
  const float maxv = (float)((1 << p.bit_depth) - 1);
  for (size_t i = 0; i < p.ops.size(); ++i) {
    const Impl::Op &op = p.ops[i];
    if (op.kind == G1S_FILTER_CROP) {
      for (int c = 0; c < p.planes; ++c) {
        const int top = c ? op.a >> p.ss_y : op.a, lft = c ? op.c >> p.ss_x : op.c;
        cur.ptr[c] = cur.ptr[c] + (size_t)top * cur.pitch[c] + (size_t)lft * p.bytes;
      }
      w = op.out_w, h = op.out_h;
      continue;
    }
    // resize: horizontal pass into scratch[0], vertical pass into the destination (last op) or a ping-pong buffer
    const bool last = i + 1 == p.ops.size();
    DevPlanes mid, out;
    size_t off_mid = 0, off_out = 0;
    for (int c = 0; c < p.planes; ++c) {
      const int iw = c ? (w + p.ss_x) >> p.ss_x : w, ih = c ? (h + p.ss_y) >> p.ss_y : h;
      const int ow = c ? (op.a + p.ss_x) >> p.ss_x : op.a, oh = c ? (op.b + p.ss_y) >> p.ss_y : op.b;
      (void)iw;
      mid.pitch[c] = ((size_t)ow * p.bytes + 255) / 256 * 256;
      mid.ptr[c] = p.scratch[0] + off_mid;
      off_mid += mid.pitch[c] * ih;
      if (last) {
        out.ptr[c] = dst.ptr[c], out.pitch[c] = dst.pitch[c];
      } else {
        out.pitch[c] = mid.pitch[c];
        out.ptr[c] = p.scratch[1 + pp] + off_out;
        off_out += out.pitch[c] * oh;
      }
      const int kh = c ? 2 : 0, kv = c ? 3 : 1;
      const dim3 gh((ow + 127) / 128, ih), gv((ow + 127) / 128, oh);
      if (p.bytes == 1) {
        resize_h_kernel<uint8_t><<<gh, 128, 0, st>>>(cur.ptr[c], cur.pitch[c], mid.ptr[c], mid.pitch[c], ih, ow, op.left[kh],
                                                    op.coef[kh], op.taps[kh], maxv);
        resize_v_kernel<uint8_t><<<gv, 128, 0, st>>>(mid.ptr[c], mid.pitch[c], out.ptr[c], out.pitch[c], ow, oh, op.left[kv],
                                                    op.coef[kv], op.taps[kv], maxv);
      } else {
        resize_h_kernel<uint16_t><<<gh, 128, 0, st>>>(reinterpret_cast<const uint16_t *>(cur.ptr[c]), cur.pitch[c],
                                                     reinterpret_cast<uint16_t *>(mid.ptr[c]), mid.pitch[c], ih, ow,
                                                     op.left[kh], op.coef[kh], op.taps[kh], maxv);
        resize_v_kernel<uint16_t><<<gv, 128, 0, st>>>(reinterpret_cast<const uint16_t *>(mid.ptr[c]), mid.pitch[c],
                                                     reinterpret_cast<uint16_t *>(out.ptr[c]), out.pitch[c], ow, oh,
                                                     op.left[kv], op.coef[kv], op.taps[kv], maxv);
      }
    }
    cur = out;
    pp ^= 1;
    w = op.a, h = op.b;
  }
  // a chain that ends in a crop (or is only crops) leaves a view: copy it into the destination
  if (p.ops.empty() || p.ops.back().kind == G1S_FILTER_CROP) {
    for (int c = 0; c < p.planes; ++c) {
      const int cw = c ? (w + p.ss_x) >> p.ss_x : w, ch = c ? (h + p.ss_y) >> p.ss_y : h;
      if (cudaMemcpy2DAsync(dst.ptr[c], dst.pitch[c], cur.ptr[c], cur.pitch[c], (size_t)cw * p.bytes, ch,
                            cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        p.err = "cudaMemcpy2DAsync failed in the filter chain";
        return false;
      }
    }
  }
  if (cudaGetLastError() != cudaSuccess) {
    p.err = "a filter kernel failed to launch";
    return false;
  }
  return true;
}

}  // namespace g1s
