// common.cuh -- error plumbing and small device helpers shared by every translation unit.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/mscs.h"

namespace mscs {

// thread-local last-error string (the only mutable state in the library)
void set_error(const char* fmt, ...);
const char* get_error();

#define MSCS_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) { ::mscs::set_error(__VA_ARGS__); return -1; }     \
  } while (0)

#define MSCS_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::mscs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                        \
      return (int)_e;                                                                     \
    }                                                                                     \
  } while (0)

#define MSCS_LAUNCH_CHECK()                                                               \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ::mscs::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),       \
                        __FILE__, __LINE__);                                              \
      return (int)_e;                                                                     \
    }                                                                                     \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mscs
