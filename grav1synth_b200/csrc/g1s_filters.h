// Device-side source filters (crop / resize) of the `diff` command; see g1s_filters.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "../../include/g1s.h"

namespace g1s {

struct DevPlanes {
  uint8_t *ptr[3] = {nullptr, nullptr, nullptr};
  size_t pitch[3] = {0, 0, 0};
};

// weights of one resize axis: left[i] = first source sample of output sample i, coef[i * taps + j] its weights
void build_resize_table(int alg, int src, int dst, std::vector<int> &left, std::vector<float> &coef, int &taps);

class SourceFilters {
 public:
  SourceFilters();
  ~SourceFilters();
  SourceFilters(const SourceFilters &) = delete;
  // the current device owns the tables and scratch buffers
  bool configure(const g1s_filter_op *ops, size_t n, int src_w, int src_h, int ss_x, int ss_y, int planes, int bit_depth);
  int out_width() const;
  int out_height() const;
  // src: planes of src_w x src_h on the device; dst: planes of out_w x out_h.  Stream ordered, no synchronisation.
  bool apply(const DevPlanes &src, const DevPlanes &dst, cudaStream_t st);
  const char *error() const;

 private:
  struct Impl;
  Impl *p_;
};

}  // namespace g1s
