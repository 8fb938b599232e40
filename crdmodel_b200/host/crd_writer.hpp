// crd_writer.hpp — the per-subdomain text output of the reference (" %.16e" per value, one line per output time,
// src/FHNmodel_torus.cpp:393-410,438-455), taken off the time loop's critical path: the state snapshot is handed
// to a background thread, which formats it with all host cores (snprintf into per-chunk buffers, the bytes are
// exactly what fprintf would produce) and writes the chunks in order while the GPU integrates the next interval.
#pragma once
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace crd {

class AsyncWriter {
 public:
  AsyncWriter(FILE *f0, FILE *f1, bool all_vars, long npoints, int nthreads = 0)
      : f0_(f0), f1_(f1), all_(all_vars), n_(npoints) {
    nt_ = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt_ < 1) nt_ = 1;
    if (nt_ > 32) nt_ = 32;
    worker_ = std::thread([this] { run(); });
  }
  ~AsyncWriter() { finish(); }

  // copy the interleaved state (2*npoints doubles) and queue it; blocks while two snapshots are already pending
  void submit(const double *state) {
    std::vector<double> snap(state, state + 2 * n_);
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return q_.size() < 2; });
    q_.push_back(std::move(snap));
    cv_.notify_all();
  }

  void finish() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (done_) return;
      done_ = true;
      cv_.notify_all();
    }
    if (worker_.joinable()) worker_.join();
  }

 private:
  void run() {
    for (;;) {
      std::vector<double> snap;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty() || done_; });
        if (q_.empty()) return;
        snap = std::move(q_.front());
        q_.pop_front();
        cv_.notify_all();
      }
      write_var(snap, 0, f0_);
      if (all_) write_var(snap, 1, f1_);
    }
  }

  void write_var(const std::vector<double> &s, int var, FILE *f) {
    const int T = (n_ < 4096) ? 1 : nt_;
    std::vector<std::vector<char>> buf(T);
    std::vector<size_t> len(T, 0);
    auto fmt = [&](int t) {
      const long a = n_ * t / T, b = n_ * (t + 1) / T;
      buf[t].resize((size_t)(b - a) * 25 + 2);
      size_t pos = 0;
      for (long k = a; k < b; ++k) pos += (size_t)snprintf(&buf[t][pos], 26, " %.16e", s[2 * k + var]);
      len[t] = pos;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(fmt, t);
    fmt(0);
    for (auto &x : th) x.join();
    for (int t = 0; t < T; ++t) fwrite(buf[t].data(), 1, len[t], f);
    fputc('\n', f);
  }

  FILE *f0_, *f1_;
  bool all_;
  long n_;
  int nt_;
  std::thread worker_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::vector<double>> q_;
  bool done_ = false;
};

}  // namespace crd
