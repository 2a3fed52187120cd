// common.cpp — error reporting and misc host helpers for libkws_b200.so.
#include "common.h"

#include <string.h>

namespace kws {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int device_sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

}  // namespace kws

extern "C" const char* kws_last_error(void) { return kws::g_error; }
extern "C" int kws_abi_version(void) { return 1; }
