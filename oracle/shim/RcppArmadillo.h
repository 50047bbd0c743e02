// RcppArmadillo.h stand-in — TEST INFRASTRUCTURE ONLY (oracle/shim, see oracle/Makefile).
//
// Just enough of the Rcpp surface (SEXP, List, Named, XPtr, NumericVector, IntegerVector,
// IntegerMatrix, CharacterMatrix, as<>, wrap, exception) for the reference's
// clusterFunArmadillo.cpp / AugTree.cpp to compile UNMODIFIED without R.  An "R object" is a small
// tagged record; oracle/ref_harness.cpp builds them from plain arrays and reads results back.
#pragma once
#include "armadillo_shim.h"
#include <exception>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

struct SEXPREC {
  enum Kind { NIL, REAL, INT, STR, LIST, EXTPTR, RAW } kind;
  std::vector<double> real;
  std::vector<int> integer;
  std::vector<std::string> str;
  std::vector<SEXPREC*> list;
  std::vector<std::string> names;
  std::vector<unsigned char> raw;
  std::map<std::string, SEXPREC*> attrs;
  int nrow, ncol;
  void* ptr;
  explicit SEXPREC(Kind k) : kind(k), nrow(-1), ncol(-1), ptr(0) {}
};
typedef SEXPREC* SEXP;

namespace shim {
inline std::vector<SEXP>& arena() { static std::vector<SEXP> a; return a; }
inline SEXP alloc(SEXPREC::Kind k) { SEXP s = new SEXPREC(k); arena().push_back(s); return s; }
inline void release_from(size_t mark) {
  std::vector<SEXP>& a = arena();
  for (size_t i = mark; i < a.size(); i++) delete a[i];
  a.resize(mark);
}
}  // namespace shim

namespace Rcpp {

class exception : public std::exception {
  std::string msg_;
 public:
  explicit exception(const char* m) : msg_(m) {}
  virtual ~exception() throw() {}
  virtual const char* what() const throw() { return msg_.c_str(); }
};

class NumericVector;
class IntegerVector;
class List;
class RawVector;
inline SEXP wrap(SEXP s);
inline SEXP wrap(const NumericVector& v);
inline SEXP wrap(const IntegerVector& v);
inline SEXP wrap(const List& l);
inline SEXP wrap(const RawVector& v);

// obj.attr("name"): readable (missing -> NULL object) and assignable
extern SEXPREC g_shim_nil;
class AttrProxy {
  SEXP obj_;
  std::string name_;
 public:
  AttrProxy(SEXP o, const std::string& n) : obj_(o), name_(n) {}
  operator SEXP() const { std::map<std::string, SEXP>::const_iterator it = obj_->attrs.find(name_); return it == obj_->attrs.end() ? &g_shim_nil : it->second; }
  template <class T> AttrProxy& operator=(const T& v) { obj_->attrs[name_] = wrap(v); return *this; }
};

class NumericVector {
 public:
  SEXP s;
  NumericVector(SEXP x) : s(x) {}
  NumericVector(int n) : s(shim::alloc(SEXPREC::REAL)) { s->real.assign(n, 0.0); }
  template <class It> NumericVector(It b, It e) : s(shim::alloc(SEXPREC::REAL)) { for (; b != e; ++b) s->real.push_back((double)*b); }
  template <class A, class B, class C, class D> static NumericVector create(A a, B b, C c, D d) {
    NumericVector v(4); v.s->real[0] = (double)a; v.s->real[1] = (double)b; v.s->real[2] = (double)c; v.s->real[3] = (double)d; return v;
  }
  AttrProxy attr(const std::string& n) { return AttrProxy(s, n); }
  operator SEXP() const { return s; }
  int size() const { return (int)s->real.size(); }
  double& operator[](int i) { return s->real[i]; }
};
class IntegerVector {
 public:
  SEXP s;
  IntegerVector(SEXP x) : s(x) {
    if (x->kind == SEXPREC::REAL && x->integer.empty()) for (size_t i = 0; i < x->real.size(); i++) x->integer.push_back((int)x->real[i]);
  }
  template <class A, class B> static IntegerVector create(A a, B b) {
    SEXP x = shim::alloc(SEXPREC::INT); x->integer.push_back((int)a); x->integer.push_back((int)b); return IntegerVector(x);
  }
  operator SEXP() const { return s; }
  int size() const { return (int)s->integer.size(); }
  int& operator[](int i) { return s->integer[i]; }
};
class IntegerMatrix {
 public:
  SEXP s;
  IntegerMatrix(SEXP x) : s(x) {}
  operator SEXP() const { return s; }
  int nrow() const { return s->nrow; }
  int ncol() const { return s->ncol; }
};
class CharacterMatrix {
 public:
  SEXP s;
  CharacterMatrix(SEXP x) : s(x) {}
  operator SEXP() const { return s; }
  int nrow() const { return s->nrow; }
  int ncol() const { return s->ncol; }
  std::string operator()(int i, int j) const { return s->str[(size_t)i + (size_t)j * s->nrow]; }
};

template <class T>
class XPtr {
  T* p_;
 public:
  XPtr(T* p, bool /*set_delete_finalizer*/) : p_(p) {}
  XPtr(SEXP s) : p_(static_cast<T*>(s->ptr)) {}
  T* operator->() const { return p_; }
  T* get() const { return p_; }
  operator SEXP() const { SEXP s = shim::alloc(SEXPREC::EXTPTR); s->ptr = p_; return s; }
};

// ---- wrap
inline SEXP wrap(SEXP s) { return s; }
inline SEXP wrap(int x) { SEXP s = shim::alloc(SEXPREC::INT); s->integer.push_back(x); return s; }
inline SEXP wrap(double x) { SEXP s = shim::alloc(SEXPREC::REAL); s->real.push_back(x); return s; }
template <class T> SEXP wrap(const XPtr<T>& p) { return static_cast<SEXP>(p); }
template <class T>
SEXP wrap(const arma::Mat<T>& m) {   // arma matrices of uword come back to R as numeric with dim
  SEXP s = shim::alloc(SEXPREC::REAL);
  s->real.assign(m.mem.begin(), m.mem.end());
  s->nrow = (int)m.n_rows; s->ncol = (int)m.n_cols;
  return s;
}
template <class T> SEXP wrap(const arma::Col<T>& m) { return wrap(static_cast<const arma::Mat<T>&>(m)); }
template <class T>
SEXP wrap(const std::vector<T>& v) {
  SEXP s = shim::alloc(SEXPREC::LIST);
  for (size_t i = 0; i < v.size(); i++) s->list.push_back(wrap(v[i]));
  return s;
}

struct NamedValue { std::string name; SEXP val; };
class Named {
  std::string n_;
 public:
  explicit Named(const std::string& n) : n_(n) {}
  template <class T> NamedValue operator=(const T& v) const { NamedValue nv; nv.name = n_; nv.val = wrap(v); return nv; }
};

class ListProxy {    // element of a list: readable as an object, assignable
  SEXP l_;
  int i_;
 public:
  ListProxy(SEXP l, int i) : l_(l), i_(i) {}
  operator SEXP() const { return l_->list[i_]; }
  SEXP operator->() const { return l_->list[i_]; }
  template <class T> ListProxy& operator=(const T& v) { l_->list[i_] = wrap(v); return *this; }
};

class List {
 public:
  SEXP s;
  List() : s(shim::alloc(SEXPREC::LIST)) {}
  List(SEXP x) : s(x) {}
  List(int n) : s(shim::alloc(SEXPREC::LIST)) { s->list.assign(n, &g_shim_nil); }
  operator SEXP() const { return s; }
  int size() const { return (int)s->list.size(); }
  ListProxy operator[](int i) const { return ListProxy(s, i); }
  AttrProxy attr(const std::string& n) { return AttrProxy(s, n); }
  void push(const NamedValue& nv) { s->list.push_back(nv.val); s->names.push_back(nv.name); }
  static List create(const NamedValue& a) { List l; l.push(a); return l; }
  static List create(const NamedValue& a, const NamedValue& b) { List l; l.push(a); l.push(b); return l; }
  static List create(const NamedValue& a, const NamedValue& b, const NamedValue& c) { List l; l.push(a); l.push(b); l.push(c); return l; }
};

class RObject {
  SEXP s_;
 public:
  RObject(SEXP s) : s_(s) {}
  bool isNULL() const { return s_ == 0 || s_->kind == SEXPREC::NIL; }
  operator SEXP() const { return s_; }
};

class RawVector {
 public:
  SEXP s;
  explicit RawVector(size_t n) : s(shim::alloc(SEXPREC::RAW)) { s->raw.assign(n, 0); }
  RawVector(SEXP x) : s(x) {}
  unsigned char& operator[](size_t i) { return s->raw[i]; }
  long size() const { return (long)s->raw.size(); }
  AttrProxy attr(const std::string& n) { return AttrProxy(s, n); }
  operator SEXP() const { return s; }
};
}  // namespace Rcpp
// R's type tags, as far as the drop-in shim asks for them
#define RAWSXP 24
inline int TYPEOF(SEXP s) { return s == 0 || s->kind == SEXPREC::NIL ? 0 : s->kind == SEXPREC::RAW ? RAWSXP : -1; }
namespace Rcpp {

inline SEXP wrap(const NumericVector& v) { return v.s; }
inline SEXP wrap(const IntegerVector& v) { return v.s; }
inline SEXP wrap(const List& l) { return l.s; }
inline SEXP wrap(const RawVector& v) { return v.s; }

// ---- as<>
template <class T> struct Exporter;
template <> struct Exporter<arma::vec> {
  static arma::vec get(SEXP s) { arma::vec v(s->real.size()); for (size_t i = 0; i < s->real.size(); i++) v[i] = s->real[i]; return v; }
};
template <> struct Exporter<arma::uvec> {
  static arma::uvec get(SEXP s) {
    if (s->kind == SEXPREC::INT) { arma::uvec v(s->integer.size()); for (size_t i = 0; i < s->integer.size(); i++) v[i] = (arma::uword)s->integer[i]; return v; }
    arma::uvec v(s->real.size()); for (size_t i = 0; i < s->real.size(); i++) v[i] = (arma::uword)s->real[i]; return v;
  }
};
template <> struct Exporter<arma::umat> {
  static arma::umat get(SEXP s) {
    arma::umat m(s->nrow, s->ncol);
    if (s->kind == SEXPREC::INT) for (size_t i = 0; i < s->integer.size(); i++) m[i] = (arma::uword)s->integer[i];
    else for (size_t i = 0; i < s->real.size(); i++) m[i] = (arma::uword)s->real[i];
    return m;
  }
};
template <> struct Exporter<arma::mat> {
  static arma::mat get(SEXP s) { arma::mat m(s->nrow, s->ncol); for (size_t i = 0; i < s->real.size(); i++) m[i] = s->real[i]; return m; }
};
template <> struct Exporter<std::string> { static std::string get(SEXP s) { return s->str[0]; } };
template <class T> struct Exporter<std::vector<T> > {
  static std::vector<T> get(SEXP s) {
    std::vector<T> out;
    if (s->kind == SEXPREC::STR) {   // character vector -> vector<string>
      for (size_t i = 0; i < s->str.size(); i++) { SEXPREC tmp(SEXPREC::STR); tmp.str.push_back(s->str[i]); out.push_back(Exporter<T>::get(&tmp)); }
      return out;
    }
    out.reserve(s->list.size());
    for (size_t i = 0; i < s->list.size(); i++) out.push_back(Exporter<T>::get(s->list[i]));
    return out;
  }
};
template <class T> T as(SEXP s) { return Exporter<T>::get(s); }

}  // namespace Rcpp
