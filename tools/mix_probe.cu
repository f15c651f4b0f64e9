// mix_probe.cu -- does tcgen05.mma kind::f16 accept DIFFERENT 16-bit formats for A and B (bf16 x fp16 -> fp32)?
// Checks D = A * B^T (M = 128, N = 64, K = 64) exactly on small integers for four format pairs, A from smem (SS) and
// A from tensor memory (TS).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mix_probe tools/mix_probe.cu
#include <cstdio>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "../eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200/csrc/ptx.cuh"
using namespace mscs;

__host__ __device__ inline float aval(int r, int k) { return (float)(((r * 7 + k * 3) % 11) - 5); }
__host__ __device__ inline float bval(int n, int k) { return (float)(((n * 5 + k * 2) % 7) - 3) * 0.5f; }

__device__ uint16_t enc(float v, int fmt) {      // fmt: 0 = fp16, 1 = bf16
  if (fmt) { __nv_bfloat16 b = __float2bfloat16(v); return *reinterpret_cast<uint16_t*>(&b); }
  __half h = __float2half(v); return *reinterpret_cast<uint16_t*>(&h);
}
__device__ void fill_tile(uint8_t* base, int rows, bool isA, int fmt) {
  for (int e = threadIdx.x; e < rows * 64; e += blockDim.x) {
    int r = e / 64, k = e % 64;
    int chunk = (k / 8) ^ (r & 7);
    *reinterpret_cast<uint16_t*>(base + r * 128 + chunk * 16 + (k % 8) * 2) = enc(isA ? aval(r, k) : bval(r, k), fmt);
  }
}
__host__ __device__ constexpr uint32_t idesc_fmt(int M, int N, int afmt, int bfmt) {
  return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int N = 64;
__global__ void __launch_bounds__(128, 1) probe(float* out, int afmt, int bfmt, int ts) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem; uint8_t* smB = smem + 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smB + N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fill_tile(smA, 128, true, afmt);
  fill_tile(smB, N, false, bfmt);
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc(slot, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = *slot, tD = tm, tA = tm + 256;
  {
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < 32; c0 += 8) {
      uint32_t v[8];
      for (int c = 0; c < 8; ++c)
        v[c] = (uint32_t)enc(aval(r, 2 * (c0 + c)), afmt) | ((uint32_t)enc(aval(r, 2 * (c0 + c) + 1), afmt) << 16);
      ptx::tmem_st8(tm + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t idesc = idesc_fmt(128, N, afmt, bfmt);
  const uint32_t a_addr = ptx::smem_u32(smA), b_addr = ptx::smem_u32(smB);
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) {
      const uint64_t bd = ptx::umma_desc_sw128(b_addr + k * 32, 16, 1024);
      if (!ts) ptx::umma_ss(tD, ptx::umma_desc_sw128(a_addr + k * 32, 16, 1024), bd, idesc, k != 0);
      else ptx::umma_ts(tD, tA + k * 8, bd, idesc, k != 0);
    }
    ptx::umma_commit(bar);
  }
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld32(tD + ((uint32_t)(warp * 32) << 16) + c0, v);
    ptx::tmem_ld_wait(v);
    for (int c = 0; c < 32; ++c) out[(size_t)(warp * 32 + lane) * N + c0 + c] = __uint_as_float(v[c]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

int main() {
  float* o; cudaMalloc(&o, 128 * N * 4);
  size_t smem = 1024 + 16384 + N * 128 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* nm[2] = {"fp16", "bf16"};
  for (int ts = 0; ts < 2; ++ts)
    for (int af = 0; af < 2; ++af)
      for (int bf = 0; bf < 2; ++bf) {
        cudaMemset(o, 0, 128 * N * 4);
        probe<<<1, 128, smem>>>(o, af, bf, ts);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("A %s x B %s (%s): CUDA error %s\n", nm[af], nm[bf], ts ? "TS" : "SS", cudaGetErrorString(e)); return 1; }
        std::vector<float> h(128 * N);
        cudaMemcpy(h.data(), o, 128 * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < 128; ++r)
          for (int n = 0; n < N; ++n) {
            float ref = 0;
            for (int k = 0; k < 64; ++k) ref += aval(r, k) * bval(n, k);
            if (h[r * N + n] != ref) ++bad;
          }
        printf("A %s x B %s (%s): %d/%d wrong\n", nm[af], nm[bf], ts ? "TS" : "SS", bad, 128 * N);
      }
  return 0;
}
