// common.h — error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/ance_b200.h"

namespace ance {

void set_error(const char* fmt, ...);  // defined in capi.cu

#define ANCE_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::ance::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return ANCE_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

#define ANCE_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::ance::set_error(__VA_ARGS__);    \
      return ANCE_ERR_INVALID;           \
    }                                    \
  } while (0)

}  // namespace ance
