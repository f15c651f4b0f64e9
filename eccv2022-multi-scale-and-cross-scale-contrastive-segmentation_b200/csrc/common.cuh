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

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// A step is a chain of ~19 launches, most of them a few microseconds long: with plain stream order every boundary
// costs the launch latency plus the drain of the previous grid.  Kernels launched through launch_k() carry the
// programmatic-stream-serialization attribute (unless MSCS_PDL=0): the next grid is scheduled as soon as every CTA of
// the previous one has executed pdl_trigger() (first statement of every kernel) and SM resources allow, runs its
// prologue, and blocks in pdl_wait() until the previous grid has COMPLETED and its writes are visible.  Every thread of
// every kernel executes pdl_wait() before its first global-memory access, so a grid completes only after all of its
// predecessors did (ordering stays transitive).  Launched without the attribute both instructions are no-ops.
bool pdl_enabled();        // api.cu: MSCS_PDL environment switch, read once

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... P, typename... A>
static inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mscs
