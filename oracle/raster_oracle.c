/*
 * oracle/raster_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load the library built from this file.  See raster_oracle.inc for the
 * algorithm, its citations and the "parity unpinned" statement.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp, generic x86-64)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* variants of the restated rasterizer (see raster_oracle.inc); process-wide, set by tests only */
static int acfm_variant_zmax_keps = 0;
static int acfm_variant_tie_cuda = 0;

#define REAL float
#define SFX(x) x##_f32
#include "raster_oracle.inc"
#undef REAL
#undef SFX
#undef K_EPS

#define REAL double
#define SFX(x) x##_f64
#include "raster_oracle.inc"
#undef REAL
#undef SFX
#undef K_EPS

void acfm_oracle_set_variant(double k_eps, int zmax_keps, int tie_cuda) {
  k_eps_f32 = (float)k_eps;
  k_eps_f64 = k_eps;
  acfm_variant_zmax_keps = zmax_keps;
  acfm_variant_tie_cuda = tie_cuda;
}

int acfm_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
