// TEST INFRASTRUCTURE ONLY -- minimal stand-in for <Rcpp.h>.
//
// Purpose: let the reference's nine hot-path translation units
// (/root/reference/src/{Rmain,cluster,containers,pval,error,kmers,misc,
// nwalign_endsfree,nwalign_vectorized}.cpp) compile UNMODIFIED without R, so the
// real reference implementation can be executed as the parity oracle
// (oracle/_ref/libdada2ref.so, see oracle/Makefile).  Only the Rcpp surface those
// files touch is provided.  Containers have Rcpp's reference (shared) semantics.
// Rcpp::ppois is routed to oracle/rmath_ppois.c (restatement of R nmath).
#ifndef ORACLE_STUB_RCPP_H
#define ORACLE_STUB_RCPP_H

#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cstdint>
#include <climits>
#include <cmath>
#include <string>
#include <vector>
#include <memory>
#include <stdexcept>
#include <utility>

extern "C" double oracle_ppois(double x, double lambda, int lower_tail, int log_p);

#define NA_INTEGER INT_MIN
static inline double oracle_na_real() {
  // R's NA_real_: quiet NaN with low word 1954
  union { double d; uint64_t u; } v; v.u = 0x7FF00000000007A2ULL; return v.d;
}
#define NA_REAL (oracle_na_real())

static inline void Rprintf(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
}

namespace Rcpp {

struct exception : public std::runtime_error {
  explicit exception(const std::string &m) : std::runtime_error(m) {}
};

inline void stop(const char *fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  throw exception(buf);
}
inline void checkUserInterrupt() {}

struct Nil {};                      // R_NilValue (evaluate.cpp returns it from vector-valued functions on bad input)
#define R_NilValue (Rcpp::Nil())

template <typename T> class Vec {
 public:
  std::shared_ptr<std::vector<T>> d;
  Vec() : d(std::make_shared<std::vector<T>>()) {}
  Vec(Nil) : d(std::make_shared<std::vector<T>>()) {}
  explicit Vec(size_t n) : d(std::make_shared<std::vector<T>>(n, T())) {}
  Vec(size_t n, const T &v) : d(std::make_shared<std::vector<T>>(n, v)) {}
  size_t size() const { return d->size(); }
  T &operator[](size_t i) { return (*d)[i]; }
  const T &operator[](size_t i) const { return (*d)[i]; }
  T &operator()(size_t i) { return (*d)[i]; }
  void push_back(const T &v) { d->push_back(v); }
};
struct Any;
class IntegerVector : public Vec<int> {
 public: using Vec<int>::Vec;
  template <typename... A> static IntegerVector create(const A &...a);       // IntegerVector::create(_["match"]=m, ...) (evaluate.cpp:118)
};
class LogicalVector : public Vec<int> { public: using Vec<int>::Vec; };
class NumericVector : public Vec<double> {
 public: using Vec<double>::Vec;
  static double get_na() { return NA_REAL; }
};
class CharacterVector : public Vec<std::string> {
 public: using Vec<std::string>::Vec;
  CharacterVector() : Vec<std::string>() {}
  CharacterVector(const std::string &one) : Vec<std::string>() { d->push_back(one); }     // `return(ostr);` (evaluate.cpp:173)
};

template <typename T> class Mat {  // column-major like R
 public:
  std::shared_ptr<std::vector<T>> d; int nr, nc;
  Mat() : d(std::make_shared<std::vector<T>>()), nr(0), nc(0) {}
  Mat(int r, int c) : d(std::make_shared<std::vector<T>>((size_t)r * c, T())), nr(r), nc(c) {}
  int nrow() const { return nr; }
  int ncol() const { return nc; }
  T &operator()(size_t r, size_t c) { return (*d)[r + (size_t)nr * c]; }
};
class NumericMatrix : public Mat<double> { public: using Mat<double>::Mat; };
class IntegerMatrix : public Mat<int> { public: using Mat<int>::Mat; };

class List;
// Type-erased named slot.
struct Any {
  std::string name;
  std::shared_ptr<std::vector<int>> iv;
  std::shared_ptr<std::vector<double>> nv;
  std::shared_ptr<std::vector<std::string>> sv;
  int nr = 0, nc = 0;           // for matrices
  std::shared_ptr<List> list;   // nested DataFrame
};

class List {
 public:
  std::vector<Any> items;
  template <typename... A> static List create(const A &...a) {
    List l; int dummy[] = {0, (l.items.push_back(a), 0)...}; (void)dummy; return l;
  }
  const Any *get(const char *nm) const {
    for (auto &a : items) if (a.name == nm) return &a;
    return nullptr;
  }
};
class DataFrame : public List {
 public:
  template <typename... A> static DataFrame create(const A &...a) {
    DataFrame l; int dummy[] = {0, (l.items.push_back(a), 0)...}; (void)dummy; return l;
  }
};

struct NamedPlaceholder {
  std::string name;
  Any operator=(int v) const { Any a; a.name = name; a.iv = std::make_shared<std::vector<int>>(1, v); return a; }
  Any operator=(const IntegerVector &v) const { Any a; a.name = name; a.iv = v.d; return a; }
  Any operator=(const NumericVector &v) const { Any a; a.name = name; a.nv = v.d; return a; }
  Any operator=(const CharacterVector &v) const { Any a; a.name = name; a.sv = v.d; return a; }
  Any operator=(const std::vector<std::string> &v) const {
    Any a; a.name = name; a.sv = std::make_shared<std::vector<std::string>>(v); return a;
  }
  Any operator=(const IntegerMatrix &m) const { Any a; a.name = name; a.iv = m.d; a.nr = m.nr; a.nc = m.nc; return a; }
  Any operator=(const NumericMatrix &m) const { Any a; a.name = name; a.nv = m.d; a.nr = m.nr; a.nc = m.nc; return a; }
  Any operator=(const DataFrame &df) const { Any a; a.name = name; a.list = std::make_shared<List>(df); return a; }
};
struct Placeholder {
  NamedPlaceholder operator[](const char *nm) const { NamedPlaceholder p; p.name = nm; return p; }
};
static const Placeholder _ = Placeholder();

template <typename... A> IntegerVector IntegerVector::create(const A &...a) {
  IntegerVector v; int dummy[] = {0, (v.push_back((*a.iv)[0]), 0)...}; (void)dummy; return v;
}

inline NumericVector ppois(const IntegerVector &x, double lambda, bool lower = true, bool log_p = false) {
  NumericVector r(x.size());
  for (size_t i = 0; i < x.size(); i++) r[i] = oracle_ppois((double)x[i], lambda, lower ? 1 : 0, log_p ? 1 : 0);
  return r;
}
template <typename T> T as(const NumericVector &v) { return (T)v[0]; }

}  // namespace Rcpp
#endif
