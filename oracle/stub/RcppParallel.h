// TEST INFRASTRUCTURE ONLY -- stand-in for <RcppParallel.h> (TBB parallelFor).
// parallelFor splits [begin,end) into chunks of >= grainSize and runs them on
// a persistent pool of std::thread workers; thread count from oracle_set_threads() (default 1).
#ifndef ORACLE_STUB_RCPPPARALLEL_H
#define ORACLE_STUB_RCPPPARALLEL_H
#include <cstddef>
#include <thread>
#include <vector>
#include <atomic>
#include <algorithm>
#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>

extern "C" int oracle_get_threads(void);

namespace RcppParallel {
// Thin views over the stub Rcpp containers (chimera.cpp:63-68 uses RMatrix<int> / RVector<int>).
template <typename T> class RMatrix {
 public:
  template <typename M> RMatrix(const M &m) : p_(m.d->data()), nr_(m.nr), nc_(m.nc) {}
  const T *begin() const { return p_; }
  T *begin() { return p_; }
  std::size_t nrow() const { return nr_; }
  std::size_t ncol() const { return nc_; }
 private:
  T *p_; std::size_t nr_, nc_;
};
template <typename T> class RVector {
 public:
  template <typename V> RVector(const V &v) : p_(v.d->data()), n_(v.d->size()) {}
  T &operator[](std::size_t i) { return p_[i]; }
  const T &operator[](std::size_t i) const { return p_[i]; }
  std::size_t size() const { return n_; }
 private:
  T *p_; std::size_t n_;
};
struct Worker {
  virtual ~Worker() {}
  virtual void operator()(std::size_t begin, std::size_t end) = 0;
};
// A persistent worker pool, like TBB's: the reference calls parallelFor once per split round (~100 x per dada() pass and
// thousands of times per bimera table), so spawning and joining nt - 1 std::threads per call would handicap the CPU arm.
class Pool {
 public:
  static Pool &get() { static Pool p; return p; }
  void run(int nt, const std::function<void()> &body) {
    std::unique_lock<std::mutex> lk(m_);
    ensure(nt - 1);
    job_ = &body; want_ = nt - 1; left_ = nt - 1; gen_++;
    cv_.notify_all();
    lk.unlock();
    body();                                   // the calling thread works too
    lk.lock();
    done_.wait(lk, [&] { return left_ == 0; });
    job_ = nullptr;
  }
 private:
  Pool() {}
  ~Pool() {
    { std::unique_lock<std::mutex> lk(m_); quit_ = true; gen_++; cv_.notify_all(); }
    for (auto &t : th_) t.join();
  }
  void ensure(int n) {
    while ((int)th_.size() < n) {
      const int id = (int)th_.size();
      th_.emplace_back([this, id] {
        unsigned long seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
          cv_.wait(lk, [&] { return quit_ || (gen_ != seen && id < want_); });
          if (quit_) return;
          seen = gen_;
          const std::function<void()> *j = job_;
          lk.unlock();
          (*j)();
          lk.lock();
          if (--left_ == 0) done_.notify_all();
        }
      });
    }
  }
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> th_;
  const std::function<void()> *job_ = nullptr;
  int want_ = 0, left_ = 0;
  unsigned long gen_ = 0;
  bool quit_ = false;
};
inline void parallelFor(std::size_t begin, std::size_t end, Worker &w, std::size_t grainSize = 1) {
  int nt = oracle_get_threads();
  if (nt <= 1 || end - begin <= grainSize) { w(begin, end); return; }
  std::size_t n = end - begin;
  std::size_t chunk = std::max<std::size_t>(grainSize, (n + (std::size_t)nt * 8 - 1) / ((std::size_t)nt * 8));
  std::atomic<std::size_t> next(begin);
  // An exception of a chunk (Rcpp::stop inside b_compare_parallel, say) is carried to the calling thread once every worker has
  // left the loop, as TBB does: thrown through run() it would unwind `next` and `body` under the workers still using them.
  std::exception_ptr first;
  std::mutex first_mu;
  const std::function<void()> body = [&]() {
    for (;;) {
      std::size_t b = next.fetch_add(chunk);
      if (b >= end) break;
      try { w(b, std::min(end, b + chunk)); }
      catch (...) {
        std::lock_guard<std::mutex> lk(first_mu);
        if (!first) first = std::current_exception();
        next.store(end);                       // no new chunks
      }
    }
  };
  Pool::get().run(nt, body);
  if (first) std::rethrow_exception(first);
}
}  // namespace RcppParallel
#endif
