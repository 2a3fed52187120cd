// common.h — shared host-side helpers for libkws_b200.so (error reporting, CUDA checks).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/kws_b200.h"

namespace kws {

void set_error(const char* fmt, ...);

#define KWS_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t err__ = (expr);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      ::kws::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
      return KWS_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define KWS_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::kws::set_error(__VA_ARGS__);  \
      return KWS_ERR_ARG;             \
    }                                 \
  } while (0)

inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int device_sm_count();

}  // namespace kws
