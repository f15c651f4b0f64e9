// tmem_probe.cu -- does tcgen05.ld traffic from epilogue warps slow down a concurrent tcgen05.mma
// stream?  One CTA per SM: warp 0 issues 128x256x16 (or 128x128x16) SS MMAs into accumulator 0
// (columns 0..N-1) while `nld` warps (4..4+nld-1) keep reading a DIFFERENT TMEM region
// (columns 256..) with tcgen05.ld.32x32b.x32.  Reports cycles per MMA with and without the readers,
// and the readers' achieved bytes/clk.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200/csrc/ptx.cuh"
using namespace mscs;

template <int N>
__global__ void __launch_bounds__(640, 1) probe(long long* out, int iters, int nld, int ld_cols, int mufu) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                 // [4][128][64]
  uint8_t* smB = smem + 4 * 16384;     // [4][N][64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smB + 4 * N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(slot + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); *stop = 0; }
  if (warp == 0) ptx::tmem_alloc(slot, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = *slot;
  const uint32_t idesc = ptx::umma_idesc_bf16(128, N, 0, 0);
  const uint32_t a_addr = ptx::smem_u32(smA), b_addr = ptx::smem_u32(smB);
  if (iters == 0) {
    // readers only: fixed number of loads per warp, no MMA traffic
    if (warp >= 4 && warp < 4 + nld) {
      const int quad = warp & 3;
      float acc = 0.f;
      long long t0 = clock64();
      for (int r = 0; r < 512; ++r)
        for (int c = 0; c < ld_cols; c += 32) {
          uint32_t v[32];
          ptx::tmem_ld32(tm + ((uint32_t)(quad * 32) << 16) + 256 + c, v);
          ptx::tmem_ld_wait(v);
          acc += __uint_as_float(v[0]) + __uint_as_float(v[13]) + __uint_as_float(v[31]);
        }
      long long t1 = clock64();
      if (blockIdx.x == 0 && lane == 0) { out[1 + (warp - 4)] = 512 * (ld_cols / 32); out[20] = t1 - t0; }
      if (acc == 123.456f) out[30] = 1;
    }
  } else if (warp == 0) {
    long long t0 = clock64();
    if (ptx::elect_one()) {
      for (int i = 0; i < iters; ++i) {
        const int kb = (i >> 2) & 3, k = i & 3;
        const uint64_t ad = ptx::umma_desc_sw128(a_addr + kb * 16384 + k * 32, 16, 1024);
        const uint64_t bd = ptx::umma_desc_sw128(b_addr + kb * N * 128 + k * 32, 16, 1024);
        ptx::umma_ss(tm, ad, bd, idesc, (i & 15) != 0);
      }
      ptx::umma_commit(bar);
    }
    __syncwarp();
    ptx::mbar_wait(bar, 0);
    long long t1 = clock64();
    if (lane == 0) { *stop = 1; if (blockIdx.x == 0) out[0] = t1 - t0; }
  } else if (warp >= 4 && warp < 4 + nld) {
    const int quad = warp & 3;
    long long n = 0;
    float acc = 0.f;
    long long t0 = clock64();
    while (!*stop) {
      for (int c = 0; c < ld_cols; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(tm + ((uint32_t)(quad * 32) << 16) + 256 + c, v);
        ptx::tmem_ld_wait(v);
        if (mufu == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += ptx::ex2(__uint_as_float(v[j]) * 1e-30f);
        } else if (mufu >= 5) {      // packed f32x2: 5 -> no polynomial, 6 -> 1/4, 7 -> 1/2
          uint64_t a01 = 0, a23 = 0;
          const uint64_t sc2 = ptx::pack2(1e-30f, 1e-30f);
          const int mask = mufu == 5 ? 0 : mufu == 6 ? 0x8 : 0xA;      // per element PAIR (j/2 & 3)
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const uint64_t x = ptx::pack2u(v[j], v[j + 1]);
            uint64_t e;
            if ((mask >> ((j >> 1) & 3)) & 1) {
              e = ptx::ex2_poly2(x, sc2);
            } else {
              float lo, hi;
              ptx::unpack2(ptx::mul2(x, sc2), lo, hi);
              e = ptx::pack2(ptx::ex2(lo), ptx::ex2(hi));
            }
            if ((j >> 1) & 1) a23 = ptx::add2(a23, e); else a01 = ptx::add2(a01, e);
          }
          float s0, s1;
          ptx::unpack2(ptx::add2(a01, a23), s0, s1);
          acc += s0 + s1;
        } else if (mufu >= 2) {      // mix: mufu == 2 -> 1/4 polynomial, 3 -> 3/8, 4 -> 1/2
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          const int mask = mufu == 2 ? 0x88 : mufu == 3 ? 0x92 : 0xAA;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(v[j]);
            const float e = ((mask >> (j & 7)) & 1) ? ptx::ex2_poly(x, 1e-30f) : ptx::ex2(x * 1e-30f);
            if ((j & 3) == 0) a0 += e; else if ((j & 3) == 1) a1 += e; else if ((j & 3) == 2) a2 += e; else a3 += e;
          }
          acc += (a0 + a1) + (a2 + a3);
        } else {
          acc += __uint_as_float(v[0]) + __uint_as_float(v[13]) + __uint_as_float(v[31]);
        }
        ++n;
      }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[1 + (warp - 4)] = n; out[20] = t1 - t0; }
    if (acc == 123.456f) out[30] = 1;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

template <int N>
void run(int iters, int nld, int ld_cols, int mufu) {
  long long* out; cudaMalloc(&out, 64 * 8); cudaMemset(out, 0, 64 * 8);
  size_t smem = 1024 + 4 * 16384 + 4 * N * 128 + 64;
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<N><<<148, 640, smem>>>(out, iters, nld, ld_cols, mufu);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[32]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  long long nl = 0; for (int i = 0; i < nld; ++i) nl += h[1 + i];
  printf("N=%d readers=%2d mufu=%d: %.1f cycles/MMA (ideal %d); tcgen05.ld.x32: %lld loads, %.1f B/clk/SM = %.2f elements/clk/SM\n", N, nld, mufu,
         iters ? (double)h[0] / iters : 0.0, N / 2, nl, h[20] ? (double)nl * 4096 / h[20] : 0.0, h[20] ? (double)nl * 1024 / h[20] : 0.0);
  cudaFree(out);
}

int main() {

  for (int mufu : {1, 2, 5, 6, 7})
    for (int nld : {8, 16}) { run<256>(8192, nld, 256, mufu); }
  return 0;
}
