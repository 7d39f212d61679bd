// Library-level C ABI: version string and the per-thread error message behind the int status codes.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/indm_b200.h"
#include "common.cuh"

static thread_local char g_err[1024] = "";

void indm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* indm_last_error(void) { return g_err; }
extern "C" const char* indm_version(void) { return "indm_b200 0.1 (sm_100a)"; }
