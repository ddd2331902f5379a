// oracle/ref_shim/RcppArmadillo.h
//
// TEST INFRASTRUCTURE ONLY.  A minimal stand-in for the two third-party
// header libraries the reference's native file includes (Rcpp and Armadillo,
// neither present in this image), written from scratch and containing ONLY the
// handful of types and functions that /root/reference/src/optimization.cpp
// touches.  With it the reference's own, unmodified source file compiles where
// it lies (see oracle/Makefile -> oracle/_ref/libtopolow_ref.so) and is used
// to pin oracle/topolow_oracle.cpp bit for bit.
//
// Two things here are NOT neutral plumbing and are declared:
//   * std::random_device is redirected to a settable constant (the reference
//     seeds std::mt19937 from it, src/optimization.cpp:153-154), which is the
//     only way to make the reference reproducible;
//   * accu() keeps Armadillo's published two-accumulator summation order for a
//     linear proxy so that the MAE rounds as the real library's would.
// Every other operation is elementwise and has one possible IEEE result.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <limits>
#include <map>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace topolow_ref_shim {
extern unsigned g_seed;
struct fixed_device {
  unsigned operator()() { return g_seed; }
};
}  // namespace topolow_ref_shim
namespace std {
using topolow_seeded_device = ::topolow_ref_shim::fixed_device;
}
// All std headers the reference includes are already included above, so this
// macro only rewrites the one use in the reference's translation unit.
#define random_device topolow_seeded_device

// ----------------------------------------------------------------------------
// Armadillo subset
// ----------------------------------------------------------------------------
namespace arma {

typedef unsigned long long uword;

template <class T>
struct Col {
  std::vector<T> v;
  Col() {}
  explicit Col(size_t n) : v(n) {}
  size_t n_elem() const { return v.size(); }
  T& operator[](size_t i) { return v[i]; }
  const T& operator[](size_t i) const { return v[i]; }
};
typedef Col<double> vec;
typedef Col<uword> uvec;
typedef Col<long long> ivec;

struct mat {
  std::vector<double> m;
  uword n_rows = 0, n_cols = 0;
  mat() {}
  mat(uword r, uword c) : m(r * c), n_rows(r), n_cols(c) {}
  mat(const double* aux, uword r, uword c, bool /*copy_aux_mem*/) : m(aux, aux + r * c), n_rows(r), n_cols(c) {}
  double* colptr(uword c) { return m.data() + c * n_rows; }
  const double* colptr(uword c) const { return m.data() + c * n_rows; }
  bool is_finite() const {
    for (double x : m) if (!std::isfinite(x)) return false;
    return true;
  }
  mat rows(const uvec& idx) const {
    mat out(idx.n_elem(), n_cols);
    for (uword c = 0; c < n_cols; ++c)
      for (size_t r = 0; r < idx.n_elem(); ++r) out.m[c * out.n_rows + r] = m[c * n_rows + idx[r]];
    return out;
  }
};

inline mat operator-(const mat& a, const mat& b) {
  mat o(a.n_rows, a.n_cols);
  for (size_t i = 0; i < o.m.size(); ++i) o.m[i] = a.m[i] - b.m[i];
  return o;
}
inline mat square(const mat& a) {
  mat o(a.n_rows, a.n_cols);
  for (size_t i = 0; i < o.m.size(); ++i) o.m[i] = a.m[i] * a.m[i];
  return o;
}
// sum(X, 1): row sums, accumulated column by column.
inline vec sum(const mat& a, int dim) {
  vec o(a.n_rows);
  if (dim != 1) throw std::logic_error("shim: only sum(X,1)");
  for (uword r = 0; r < a.n_rows; ++r) o[r] = a.n_cols ? a.m[r] : 0.0;
  for (uword c = 1; c < a.n_cols; ++c)
    for (uword r = 0; r < a.n_rows; ++r) o[r] += a.m[c * a.n_rows + r];
  return o;
}
inline vec sqrt(const vec& a) {
  vec o(a.n_elem());
  for (size_t i = 0; i < a.n_elem(); ++i) o[i] = std::sqrt(a[i]);
  return o;
}
inline vec abs(const vec& a) {
  vec o(a.n_elem());
  for (size_t i = 0; i < a.n_elem(); ++i) o[i] = std::fabs(a[i]);
  return o;
}
inline vec operator-(const vec& a, const vec& b) {
  vec o(a.n_elem());
  for (size_t i = 0; i < a.n_elem(); ++i) o[i] = a[i] - b[i];
  return o;
}
inline vec operator%(const vec& a, const vec& b) {
  vec o(a.n_elem());
  for (size_t i = 0; i < a.n_elem(); ++i) o[i] = a[i] * b[i];
  return o;
}
inline uvec operator%(const uvec& a, const uvec& b) {
  uvec o(a.n_elem());
  for (size_t i = 0; i < a.n_elem(); ++i) o[i] = a[i] * b[i];
  return o;
}
inline uvec operator+(const uvec& a, const uvec& b) {
  uvec o(a.n_elem());
  for (size_t i = 0; i < a.n_elem(); ++i) o[i] = a[i] + b[i];
  return o;
}
inline uvec operator==(const ivec& a, int s) {
  uvec o(a.n_elem());
  for (size_t i = 0; i < a.n_elem(); ++i) o[i] = (a[i] == s) ? 1 : 0;
  return o;
}
inline uvec operator<(const vec& a, const vec& b) {
  uvec o(a.n_elem());
  for (size_t i = 0; i < a.n_elem(); ++i) o[i] = (a[i] < b[i]) ? 1 : 0;
  return o;
}
inline uvec operator>(const vec& a, const vec& b) {
  uvec o(a.n_elem());
  for (size_t i = 0; i < a.n_elem(); ++i) o[i] = (a[i] > b[i]) ? 1 : 0;
  return o;
}
// accu over a linear proxy: two running sums (even / odd), added at the end.
template <class T>
inline T accu(const Col<T>& a) {
  T v1 = T(0), v2 = T(0);
  size_t i, j;
  const size_t n = a.n_elem();
  for (i = 0, j = 1; j < n; i += 2, j += 2) { v1 += a[i]; v2 += a[j]; }
  if (i < n) v1 += a[i];
  return v1 + v2;
}

template <class Out>
struct conv_to {
  template <class In>
  static Out from(const std::vector<In>& x) {
    Out o(x.size());
    for (size_t i = 0; i < x.size(); ++i) o[i] = static_cast<decltype(o[0] + 0)>(x[i]);
    return o;
  }
  template <class In>
  static Out from(const Col<In>& x) {
    Out o(x.n_elem());
    for (size_t i = 0; i < x.n_elem(); ++i) o[i] = static_cast<decltype(o[0] + 0)>(x[i]);
    return o;
  }
};

}  // namespace arma

// ----------------------------------------------------------------------------
// Rcpp subset
// ----------------------------------------------------------------------------
namespace Rcpp {

template <class T>
struct VectorView {
  const T* p = nullptr;
  size_t n = 0;
  VectorView() {}
  VectorView(const T* p_, size_t n_) : p(p_), n(n_) {}
  const T& operator[](size_t i) const { return p[i]; }
  const T* begin() const { return p; }
  const T* end() const { return p + n; }
  size_t size() const { return n; }
};
template <class T>
struct MatrixView {
  T* p = nullptr;
  int r = 0, c = 0;
  MatrixView() {}
  MatrixView(T* p_, int r_, int c_) : p(p_), r(r_), c(c_) {}
  int nrow() const { return r; }
  int ncol() const { return c; }
  T* begin() { return p; }
  const T* begin() const { return p; }
};
typedef VectorView<int> IntegerVector;
typedef VectorView<double> NumericVector;
typedef MatrixView<double> NumericMatrix;
typedef MatrixView<int> IntegerMatrix;

template <class Out, class T>
inline Out as(const VectorView<T>& v) {
  return Out(v.begin(), v.end());
}

struct stop_error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
inline void stop(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw stop_error(buf);
}
inline void checkUserInterrupt() {}

struct NullStream {
  template <class T>
  NullStream& operator<<(const T&) { return *this; }
};
static NullStream Rcout;

struct Value {
  arma::mat m;
  double d = 0.0;
  Value() {}
  Value(const arma::mat& x) : m(x) {}
  Value(double x) : d(x) {}
  Value(int x) : d(x) {}
  Value(bool x) : d(x ? 1.0 : 0.0) {}
};
struct NamedValue {
  std::string name;
  Value value;
};
struct Named {
  std::string name;
  explicit Named(const char* n) : name(n) {}
  template <class T>
  NamedValue operator=(const T& x) const { return NamedValue{name, Value(x)}; }
};
struct List {
  std::map<std::string, Value> items;
  template <class... A>
  static List create(const A&... a) {
    List l;
    const NamedValue all[] = {a...};
    for (const auto& nv : all) l.items[nv.name] = nv.value;
    return l;
  }
};

}  // namespace Rcpp
