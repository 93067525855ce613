// Stress test of the host pool and of NoiseModel::fold_run under ThreadSanitizer (built and run by
// tests/test_pool_stress.py; no CUDA involved): back-to-back parallel_for calls of different sizes, then runs of frames
// folded with helper threads against the frame-by-frame fold.
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../grav1synth_b200/csrc/g1s_model.h"
#include "../grav1synth_b200/csrc/g1s_pool.h"

using namespace g1s;

// A frame of "scene" `seed`: the scene's systems plus a small per-frame perturbation, so that frames of one scene agree
// (is_different says no) and frames of different scenes do not.
static LatestFrame make_frame(unsigned seed, double scale, unsigned jitter) {
  LatestFrame lf;
  std::mt19937_64 rng(seed), jr(1000 + jitter);
  std::uniform_real_distribution<double> U0(0, 1), J(-1e-3, 1e-3);
  auto U = [&](std::mt19937_64 &r) { return U0(r) * (1.0 + J(jr)); };
  lf.enough_flat = true, lf.channels = 3, lf.fail_channel = -1, lf.fail_text = nullptr;
  for (int c = 0; c < 3; ++c) {
    ChannelState &s = lf.ch[c];
    const int n = s.eqns.n;
    std::vector<double> M(n * n);
    for (auto &v : M) v = U(rng) - 0.5;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double a = 0;
        for (int k = 0; k < n; ++k) a += M[i * n + k] * M[j * n + k];
        s.eqns.A[i * n + j] = scale * (a + (i == j ? 5.0 : 0));
      }
    for (int i = 0; i < n; ++i) s.eqns.b[i] = scale * U(rng);
    s.num_observations = 1000000;
    auto &e = s.strength.eqns;
    for (int i = 0; i < 20; ++i) {
      e.A[i * 20 + i] = 100 + 10 * U(rng);
      if (i + 1 < 20) e.A[i * 20 + i + 1] = e.A[(i + 1) * 20 + i] = 20 * U(rng);
      e.b[i] = 50 * scale * U(rng);
    }
    s.strength.num_equations = 7000, s.strength.total = 3000 * scale;
    s.solve_ar(c != 0);
    s.strength.solve_bumped();
  }
  return lf;
}

int main() {
  HostPool pool(6);
  // (1) the pool alone: sums of squares through calls of every size, each checked
  {
    std::mt19937 rng(3);
    for (int rep = 0; rep < 20000; ++rep) {
      const int n = 1 + (int)(rng() % 97);
      std::vector<long long> out(n, 0);
      pool.parallel_for(n, [&](int i) { out[i] = (long long)i * i + rep; });
      for (int i = 0; i < n; ++i)
        if (out[i] != (long long)i * i + rep) {
          std::printf("pool: item %d of call %d not run exactly once\n", i, rep);
          return 1;
        }
    }
  }
  // (2) fold_run against fold on a stream with a scene cut
  StreamGeometry g{};
  g.width = 640, g.height = 352, g.ss_x = 1, g.ss_y = 1, g.planes = 3, g.nbw = 20, g.nbh = 11, g.nb = 220;
  std::vector<LatestFrame> stream;
  for (int k = 0; k < 150; ++k) stream.push_back(make_frame(1, 1.0, k % 5));
  for (int k = 0; k < 120; ++k) stream.push_back(make_frame(11, 37.0, k % 7));
  for (int k = 0; k < 60; ++k) stream.push_back(make_frame(1, 1.0, k % 3));
  NoiseModel::ParallelFor par = [&](int n, const std::function<void(int)> &fn) { pool.parallel_for(n, fn); };
  for (int rep = 0; rep < 30; ++rep) {
    DiffSequencer a(24, 1, g), b(24, 1, g);
    for (const LatestFrame &lf : stream) a.consume_latest(lf);
    const int chunk = 40 + 13 * rep;
    for (size_t at = 0; at < stream.size(); at += chunk)
      b.consume_latest_batch(stream.data() + at, (int)std::min<size_t>(chunk, stream.size() - at), par);
    const std::vector<g1s_segment> ta = a.finish(), tb = b.finish();
    if (ta.size() != tb.size() || std::memcmp(ta.data(), tb.data(), ta.size() * sizeof(g1s_segment)) != 0) {
      std::printf("fold_run: tables differ (rep %d, %zu vs %zu segments)\n", rep, ta.size(), tb.size());
      return 1;
    }
    if (rep == 0) std::printf("segments: %zu\n", ta.size());
    if (ta.size() < 2 || ta.size() > 8) {
      std::printf("the stream was meant to have a few scene cuts, not %zu segments\n", ta.size());
      return 1;
    }
  }
  std::printf("ok\n");
  return 0;
}
