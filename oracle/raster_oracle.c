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

int acfm_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
