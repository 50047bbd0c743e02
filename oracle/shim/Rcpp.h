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
