// TEST INFRASTRUCTURE ONLY -- stand-in for <RcppParallel.h> (TBB parallelFor).
// parallelFor splits [begin,end) into chunks of >= grainSize and runs them on
// std::thread workers; thread count from oracle_set_threads() (default 1).
#ifndef ORACLE_STUB_RCPPPARALLEL_H
#define ORACLE_STUB_RCPPPARALLEL_H
#include <cstddef>
#include <thread>
#include <vector>
#include <atomic>
#include <algorithm>

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
inline void parallelFor(std::size_t begin, std::size_t end, Worker &w, std::size_t grainSize = 1) {
  int nt = oracle_get_threads();
  if (nt <= 1 || end - begin <= grainSize) { w(begin, end); return; }
  std::size_t n = end - begin;
  std::size_t chunk = std::max<std::size_t>(grainSize, (n + (std::size_t)nt * 8 - 1) / ((std::size_t)nt * 8));
  std::atomic<std::size_t> next(begin);
  auto body = [&]() {
    for (;;) {
      std::size_t b = next.fetch_add(chunk);
      if (b >= end) break;
      w(b, std::min(end, b + chunk));
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(body);
  body();
  for (auto &t : th) t.join();
}
}  // namespace RcppParallel
#endif
