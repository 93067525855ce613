// Host thread pool of the engine (fork-join), header-only so that the stress test can build it without CUDA.
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace g1s {

// Minimal fork-join pool for the order independent halves of the host model.  Every parallel_for is its own Job
// object: a worker that is late leaving one call holds THAT call's job (its counter is exhausted, it just leaves) and
// cannot take items of, or count against, the next one -- calls may follow each other back to back with different n.
class HostPool {
 public:
  explicit HostPool(int threads) {
    for (int i = 1; i < threads; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> l(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }
  // runs fn(0..n-1), the caller participates; returns when all are done.  One caller at a time.
  void parallel_for(int n, const std::function<void(int)> &fn) {
    if (n <= 0) return;
    if (workers_.empty() || n == 1) {
      for (int i = 0; i < n; ++i) fn(i);
      return;
    }
    auto job = std::make_shared<Job>();
    job->fn = &fn;
    job->n = n;
    job->pending.store(n);
    {
      std::lock_guard<std::mutex> l(m_);
      job_ = job;
      ++epoch_;
    }
    cv_.notify_all();
    run_items(*job);
    std::unique_lock<std::mutex> l(m_);
    done_cv_.wait(l, [&] { return job->pending.load() == 0; });
    job_.reset();  // late workers keep their own reference; fn is not touched once pending is 0
  }

 private:
  struct Job {
    const std::function<void(int)> *fn = nullptr;
    int n = 0;
    std::atomic<int> next{0}, pending{0};
  };
  void run_items(Job &job) {
    for (;;) {
      const int i = job.next.fetch_add(1);
      if (i >= job.n) return;
      (*job.fn)(i);
      if (job.pending.fetch_sub(1) == 1) {
        std::lock_guard<std::mutex> l(m_);  // pairs with the caller's wait: no lost wake-up
        done_cv_.notify_all();
      }
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return stop_ || epoch_ != seen; });
        if (stop_) return;
        seen = epoch_;
        job = job_;
      }
      if (job) run_items(*job);
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  std::shared_ptr<Job> job_;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

}  // namespace g1s
