// Rcpp.h stand-in — TEST INFRASTRUCTURE ONLY (oracle/shim).
// What rpkg/src/clusterFunArmadillo.cpp (this repository's drop-in replacement of the reference's entry-point file)
// includes; the classes live in RcppArmadillo.h's stand-in, this adds stop() and R_NilValue.
#pragma once
#include "RcppArmadillo.h"
#include <cstdarg>
#include <cstdio>

#define R_NilValue (&::Rcpp::g_shim_nil)

namespace Rcpp {
inline void stop(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw exception(buf);
}
}  // namespace Rcpp

// ---- what the SEXP-level glue (rpkg/src/RcppExports.cpp) needs: argument holders, BEGIN_RCPP / END_RCPP, registration
#include <type_traits>
namespace shim {
inline std::string& last_error() { static std::string e; return e; }
template <class T> typename std::enable_if<std::is_arithmetic<T>::value, T>::type as(SEXP s) {
  return !s->integer.empty() && s->real.empty() ? (T)s->integer[0] : (T)s->real[0];
}
template <class T> typename std::enable_if<!std::is_arithmetic<T>::value, T>::type as(SEXP s) { return T(s); }
}  // namespace shim
namespace Rcpp {
struct RNGScope {};
namespace traits {
template <class T> struct InHolder { T v; InHolder(SEXP s) : v(shim::as<T>(s)) {} operator T() { return v; } };
template <class T> struct InHolder<T&> { T v; InHolder(SEXP s) : v(shim::as<T>(s)) {} operator T&() { return v; } };
template <class T> struct input_parameter { typedef InHolder<T> type; };
}  // namespace traits
}  // namespace Rcpp
#define RcppExport extern "C"
#define BEGIN_RCPP try {
#define END_RCPP } catch (std::exception& e_) { shim::last_error() = e_.what(); return R_NilValue; }
#ifndef FALSE
#define FALSE 0
#endif
typedef void* (*DL_FUNC)();
struct R_CallMethodDef { const char* name; DL_FUNC fun; int numArgs; };
struct DllInfo { const R_CallMethodDef* calls; };
inline int R_registerRoutines(DllInfo* d, const void*, const R_CallMethodDef* calls, const void*, const void*) { d->calls = calls; return 1; }
inline int R_useDynamicSymbols(DllInfo*, int) { return 0; }

