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

// Page-locked host staging buffers for the host-buffer entry (pipeline.py).  write_combined = 1 asks for
// cudaHostAllocWriteCombined: the pages are not snooped during the DMA, which is what lets a host->device copy run at
// the PCIe line rate on this platform (measured on the B200 box: 55 GB/s against 17-35 GB/s from cacheable pinned
// pages); the host should only WRITE such a buffer (audio decode / capture output), CPU reads from it are slow.
extern "C" int kws_host_alloc(void** out, size_t bytes, int write_combined) {
  KWS_REQUIRE(out != nullptr && bytes > 0, "kws_host_alloc: NULL out or zero size");
  *out = nullptr;
  KWS_CUDA_CHECK(cudaHostAlloc(out, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  return KWS_OK;
}
extern "C" int kws_host_free(void* p) {
  if (p) KWS_CUDA_CHECK(cudaFreeHost(p));
  return KWS_OK;
}
