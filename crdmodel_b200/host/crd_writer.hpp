// crd_writer.hpp — the per-subdomain text output of the reference (" %.16e" per value, one line per output time,
// src/FHNmodel_torus.cpp:393-410,438-455), taken off the time loop's critical path.
//
// The time loop only ENQUEUES an output (crd_snapshot_begin, include/crd_b200.h: a small gather kernel on the integrator's
// stream splits the interleaved state into contiguous per-variable arrays — variable 0 only unless includeAllVars — and an
// asynchronous device-to-host copy into one of a few page-locked buffers follows on a side stream) and goes on integrating.
// This writer's background thread waits for the copy's event, formats straight from the page-locked buffer with all host
// cores (snprintf into per-chunk buffers: the bytes are exactly what fprintf would produce), writes the chunks in order and
// hands the buffer back.  No host copy of the state is made anywhere.
#pragma once
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace crd {

// what the worker gets when it asks for a queued output: contiguous values of variable 0 (and 1, or NULL), and how to give
// the storage back
struct OutputView {
  const double *v0 = nullptr, *v1 = nullptr;
  std::function<void()> release;
};

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

  // queue one output line (per variable): `wait` is called on the worker thread, blocks until the data are on the host and
  // returns where they are.  Never blocks the caller (the storage behind `wait` bounds how many can be in flight).
  void submit(std::function<OutputView()> wait) {
    std::lock_guard<std::mutex> lk(mu_);
    q_.push_back(std::move(wait));
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

  bool failed() const { return failed_; }

 private:
  void run() {
    for (;;) {
      std::function<OutputView()> wait;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty() || done_; });
        if (q_.empty()) return;
        wait = std::move(q_.front());
        q_.pop_front();
      }
      OutputView v = wait();
      if (!v.v0) { failed_ = true; if (v.release) v.release(); continue; }
      write_var(v.v0, f0_);
      if (all_ && v.v1) write_var(v.v1, f1_);
      if (v.release) v.release();
    }
  }

  void write_var(const double *s, FILE *f) {
    const int T = (n_ < 4096) ? 1 : nt_;
    std::vector<std::vector<char>> buf(T);
    std::vector<size_t> len(T, 0);
    auto fmt = [&](int t) {
      const long a = n_ * t / T, b = n_ * (t + 1) / T;
      buf[t].resize((size_t)(b - a) * 25 + 2);
      size_t pos = 0;
      for (long k = a; k < b; ++k) pos += (size_t)snprintf(&buf[t][pos], 26, " %.16e", s[k]);
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
  std::deque<std::function<OutputView()>> q_;
  bool done_ = false;
  volatile bool failed_ = false;
};

}  // namespace crd
