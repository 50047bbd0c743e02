// armadillo_shim.h — TEST INFRASTRUCTURE ONLY (part of oracle/shim, see oracle/Makefile).
//
// A minimal, eager stand-in for the subset of Armadillo that the reference's translation units
// (/root/reference/src/{AugTree,IntermediateNode,InputNode,clusterFunArmadillo}.cpp) use, so that
// those files compile UNMODIFIED in an image that has neither R nor RcppArmadillo.  Arithmetic follows
// Armadillo's published evaluation order for the shapes that occur on this path:
//   mat*vec (4x4)      gemv_emul_tinysq: y[i] = A(i,0)*x0 + A(i,1)*x1 + A(i,2)*x2 + A(i,3)*x3
//   dot / sum / mean   two interleaved accumulators (op_dot::direct_dot_arma, arrayops::accumulate)
//   vec / scalar       true division per element;  exp(fvec) = expf;  vec % fvec promotes to double
// The Armadillo version is not pinned by the reference (DESCRIPTION:15), so this is a restatement of
// library behaviour, not the library: differences are at the last-ulp level and far below the 1e-9
// parity tolerance.
#pragma once
#include <cmath>
#include <cstddef>
#include <iostream>
#include <string>
#include <vector>
#include <sys/types.h>

namespace arma {
using std::cerr;
using std::cout;
using std::endl;
typedef unsigned long long uword;

namespace fill {
struct fill_zeros {};
struct fill_ones {};
static const fill_zeros zeros = fill_zeros();
static const fill_ones ones = fill_ones();
}  // namespace fill

template <class T>
class Mat {
 public:
  typedef T elem_type;
  typedef const T* const_iterator;
  typedef T* iterator;
  uword n_rows, n_cols, n_elem;
  std::vector<T> mem;
  Mat() : n_rows(0), n_cols(0), n_elem(0) {}
  Mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c) {}
  Mat(uword r, uword c, const fill::fill_zeros&) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c, T(0)) {}
  Mat(uword r, uword c, const fill::fill_ones&) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c, T(1)) {}
  T& at(uword i) { return mem[i]; }
  const T& at(uword i) const { return mem[i]; }
  T& at(uword r, uword c) { return mem[r + c * n_rows]; }
  const T& at(uword r, uword c) const { return mem[r + c * n_rows]; }
  T& operator()(uword i) { return mem[i]; }
  const T& operator()(uword i) const { return mem[i]; }
  T& operator()(uword r, uword c) { return mem[r + c * n_rows]; }
  const T& operator()(uword r, uword c) const { return mem[r + c * n_rows]; }
  T& operator[](uword i) { return mem[i]; }
  const T& operator[](uword i) const { return mem[i]; }
  iterator begin() { return mem.data(); }
  iterator end() { return mem.data() + n_elem; }
  const_iterator begin() const { return mem.data(); }
  const_iterator end() const { return mem.data() + n_elem; }
  const T* memptr() const { return mem.data(); }
  T* memptr() { return mem.data(); }
  uword size() const { return n_elem; }
  void print(std::ostream& o) const {
    for (uword r = 0; r < n_rows; r++) { for (uword c = 0; c < n_cols; c++) o << at(r, c) << " "; o << "\n"; }
  }
  void print(const std::string& h = "") const { if (!h.empty()) std::cout << h << "\n"; print(std::cout); }
};

template <class T>
class Col : public Mat<T> {
 public:
  Col() : Mat<T>() { this->n_cols = 1; }
  explicit Col(uword n) : Mat<T>(n, 1) {}
  Col(uword n, const fill::fill_zeros& f) : Mat<T>(n, 1, f) {}
  Col(uword n, const fill::fill_ones& f) : Mat<T>(n, 1, f) {}
  Col<T> rows(uword a, uword b) const {   // inclusive range, like Mat::rows
    Col<T> out(b - a + 1);
    for (uword i = a; i <= b; i++) out.mem[i - a] = this->mem[i];
    return out;
  }
  Col<T>& operator-=(const T val) { for (uword i = 0; i < this->n_elem; i++) this->mem[i] -= val; return *this; }
  Col<T>& operator+=(const T val) { for (uword i = 0; i < this->n_elem; i++) this->mem[i] += val; return *this; }
};

typedef Mat<double> mat;
typedef Col<double> vec;
typedef Col<float> fvec;
typedef Mat<uword> umat;
typedef Col<uword> uvec;
struct cx_mat { void print(std::ostream&) const {} };   // only named by explicit instantiations of print_matrix

// ---- element-wise / scalar operators
template <class T>
Mat<T> operator-(const Mat<T>& X, const typename Mat<T>::elem_type k) {
  Mat<T> out(X.n_rows, X.n_cols);
  for (uword i = 0; i < X.n_elem; i++) out.mem[i] = X.mem[i] - k;
  return out;
}
template <class T>
Col<T> operator+(const Col<T>& a, const Col<T>& b) {
  Col<T> out(a.n_elem);
  for (uword i = 0; i < a.n_elem; i++) out.mem[i] = a.mem[i] + b.mem[i];
  return out;
}
template <class T>
Col<T> operator%(const Col<T>& a, const Col<T>& b) {   // Schur product
  Col<T> out(a.n_elem);
  for (uword i = 0; i < a.n_elem; i++) out.mem[i] = a.mem[i] * b.mem[i];
  return out;
}
inline Col<double> operator%(const Col<double>& a, const Col<float>& b) {   // glue_mixed_schur: upgrade both
  Col<double> out(a.n_elem);
  for (uword i = 0; i < a.n_elem; i++) out.mem[i] = a.mem[i] * double(b.mem[i]);
  return out;
}
template <class T>
Col<T> operator/(const Col<T>& a, const typename Mat<T>::elem_type k) {   // eop_scalar_div_post
  Col<T> out(a.n_elem);
  for (uword i = 0; i < a.n_elem; i++) out.mem[i] = a.mem[i] / k;
  return out;
}

// ---- mat * vec: gemv_emul_tinysq for square n <= 4, plain row sums otherwise
inline Col<double> operator*(const Mat<double>& A, const Col<double>& x) {
  Col<double> y(A.n_rows);
  const double* Am = A.memptr();
  const uword n = A.n_rows;
  if (n == 4 && A.n_cols == 4) {
    const double x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3];
    for (uword i = 0; i < 4; i++) y[i] = Am[i] * x0 + Am[i + 4] * x1 + Am[i + 8] * x2 + Am[i + 12] * x3;
  } else {
    for (uword i = 0; i < n; i++) {
      double acc = 0.0;
      for (uword j = 0; j < A.n_cols; j++) acc += Am[i + j * n] * x[j];
      y[i] = acc;
    }
  }
  return y;
}

// ---- reductions
template <class T>
T max(const Col<T>& X) {
  T m = X.mem[0];
  for (uword i = 1; i < X.n_elem; i++) if (X.mem[i] > m) m = X.mem[i];
  return m;
}
template <class T>
T accu_paired(const T* X, uword n) {   // arrayops::accumulate
  T a1 = T(0), a2 = T(0);
  uword i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { a1 += X[i]; a2 += X[j]; }
  if (i < n) a1 += X[i];
  return a1 + a2;
}
template <class T> T sum(const Col<T>& X) { return accu_paired(X.memptr(), X.n_elem); }
template <class T> T mean(const Col<T>& X) { return accu_paired(X.memptr(), X.n_elem) / T(X.n_elem); }
template <class T>
T dot(const Col<T>& A, const Col<T>& B) {   // op_dot::direct_dot_arma
  T v1 = T(0), v2 = T(0);
  uword i, j;
  const uword n = A.n_elem;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { v1 += A[i] * B[i]; v2 += A[j] * B[j]; }
  if (i < n) v1 += A[i] * B[i];
  return v1 + v2;
}
inline Col<float> exp(const Col<float>& X) {   // eop_exp on float = expf
  Col<float> out(X.n_elem);
  for (uword i = 0; i < X.n_elem; i++) out.mem[i] = std::exp(X.mem[i]);
  return out;
}
inline Col<double> exp(const Col<double>& X) {
  Col<double> out(X.n_elem);
  for (uword i = 0; i < X.n_elem; i++) out.mem[i] = std::exp(X.mem[i]);
  return out;
}

// ---- conv_to
template <class Out> struct conv_to;
template <class T>
struct conv_to<Col<T> > {
  template <class U> static Col<T> from(const Col<U>& in) {
    Col<T> out(in.n_elem);
    for (uword i = 0; i < in.n_elem; i++) out.mem[i] = T(in.mem[i]);
    return out;
  }
};
template <class T>
struct conv_to<std::vector<T> > {
  template <class U> static std::vector<T> from(const Col<U>& in) {
    std::vector<T> out(in.n_elem);
    for (uword i = 0; i < in.n_elem; i++) out[i] = T(in.mem[i]);
    return out;
  }
};
}  // namespace arma
