// api.cu — error plumbing and version of the C ABI (include/acfm_b200.h).
#include <stdarg.h>

#include "common.cuh"

#include <atomic>
#include <math.h>

namespace {
thread_local char g_err[512] = "";
std::atomic<float> g_raster_eps{ACFM_K_EPS_DEFAULT};
std::atomic<int> g_bwd_headroom_bits{30};
}

float acfm_raster_epsilon() { return g_raster_eps.load(std::memory_order_relaxed); }

extern "C" int acfm_set_raster_epsilon(float eps) {
  ACFM_REQUIRE(eps >= 0.0f && isfinite(eps), ACFM_ERR_BAD_ARG, "acfm_set_raster_epsilon: eps must be finite and >= 0");
  g_raster_eps.store(eps, std::memory_order_relaxed);
  return ACFM_OK;
}

extern "C" float acfm_get_raster_epsilon(void) { return acfm_raster_epsilon(); }

int acfm_raster_bwd_headroom_bits() { return g_bwd_headroom_bits.load(std::memory_order_relaxed); }

extern "C" int acfm_set_raster_bwd_headroom_bits(int bits) {
  ACFM_REQUIRE(bits >= 16 && bits <= 30, ACFM_ERR_BAD_ARG, "acfm_set_raster_bwd_headroom_bits: bits must be in 16..30");
  g_bwd_headroom_bits.store(bits, std::memory_order_relaxed);
  return ACFM_OK;
}

void acfm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int acfm_version(void) { return 100; }  // 0.1.0

extern "C" const char* acfm_last_error_string(void) { return g_err; }
