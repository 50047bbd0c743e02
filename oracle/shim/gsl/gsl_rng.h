// gsl/gsl_rng.h stand-in — TEST INFRASTRUCTURE ONLY (oracle/shim).
// GSL's "taus" generator (rng/taus.c: Tausworthe, L'Ecuyer 1996; default seed 0 -> 1; LCG(69069)
// seeding; six warm-up draws) and gsl_rng_uniform_int (rng/rng.c).  Known answer from the GSL manual
// ("GSL_RNG_TYPE=taus GSL_RNG_SEED=123 ... first value = 2720986350") is checked in tests/.
#pragma once
#include <stdint.h>
#include <stdlib.h>
typedef struct { int id; } gsl_rng_type;
static const gsl_rng_type gsl_rng_taus_type = {1};
static const gsl_rng_type* const gsl_rng_taus = &gsl_rng_taus_type;
typedef struct { const gsl_rng_type* type; uint32_t s1, s2, s3; } gsl_rng;
static inline unsigned long gsl_rng_get(gsl_rng* r) {
#define SHIM_TAUSWORTHE(s, a, b, c, d) ((((s) & (c)) << (d)) ^ ((((s) << (a)) ^ (s)) >> (b)))
  r->s1 = SHIM_TAUSWORTHE(r->s1, 13, 19, 4294967294u, 12);
  r->s2 = SHIM_TAUSWORTHE(r->s2, 2, 25, 4294967288u, 4);
  r->s3 = SHIM_TAUSWORTHE(r->s3, 3, 11, 4294967280u, 17);
#undef SHIM_TAUSWORTHE
  return r->s1 ^ r->s2 ^ r->s3;
}
static inline void gsl_rng_set(gsl_rng* r, unsigned long s) {
  if (s == 0) s = 1;
  r->s1 = (uint32_t)(69069u * (uint32_t)s);
  r->s2 = (uint32_t)(69069u * r->s1);
  r->s3 = (uint32_t)(69069u * r->s2);
  for (int i = 0; i < 6; i++) gsl_rng_get(r);
}
static inline gsl_rng* gsl_rng_alloc(const gsl_rng_type* t) {
  gsl_rng* r = (gsl_rng*)malloc(sizeof(gsl_rng));
  r->type = t;
  gsl_rng_set(r, 0);   // gsl_rng_default_seed
  return r;
}
static inline void gsl_rng_free(gsl_rng* r) { free(r); }
static inline unsigned long gsl_rng_uniform_int(gsl_rng* r, unsigned long n) {
  const unsigned long range = 0xffffffffUL;
  if (n > range || n == 0) return 0;   // GSL_ERROR_VAL(..., GSL_EINVAL, 0)
  unsigned long scale = range / n, k;
  do { k = gsl_rng_get(r) / scale; } while (k >= n);
  return k;
}
