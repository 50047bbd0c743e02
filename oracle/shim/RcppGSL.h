// RcppGSL.h stand-in — TEST INFRASTRUCTURE ONLY (oracle/shim)
#pragma once
#include <gsl/gsl_rng.h>
