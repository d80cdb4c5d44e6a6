// api.cu — error plumbing and version of the C ABI (include/acfm_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
}

void acfm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int acfm_version(void) { return 100; }  // 0.1.0

extern "C" const char* acfm_last_error_string(void) { return g_err; }
