// umma_probe.cu -- micro-benchmark / semantics probe for tcgen05.mma operand modes on sm_100a.
//   * checks D = A * B^T numerically for SS (A in smem) and TS (A in TMEM, written with tcgen05.st)
//   * measures cycles per MMA for SS/TS at N = 128 and N = 256 (M = 128, K = 16, bf16 -> fp32)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200/csrc/ptx.cuh"
using namespace mscs;

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// values: small integers exactly representable in bf16
__host__ __device__ inline float aval(int r, int k) { return (float)(((r * 7 + k * 3) % 11) - 5); }
__host__ __device__ inline float bval(int n, int k) { return (float)(((n * 5 + k * 2) % 7) - 3); }

// smem tile [rows][64 bf16] with the 128-byte swizzle (16-byte chunk index XOR row%8)
__device__ void fill_tile(uint8_t* base, int rows, int kofs, bool isA) {
  for (int e = threadIdx.x; e < rows * 64; e += blockDim.x) {
    int r = e / 64, k = e % 64;
    float v = isA ? aval(r, kofs + k) : bval(r, kofs + k);
    int chunk = (k / 8) ^ (r & 7);
    *reinterpret_cast<__nv_bfloat16*>(base + r * 128 + chunk * 16 + (k % 8) * 2) = __float2bfloat16(v);
  }
}

template <int N>
__global__ void __launch_bounds__(128, 1) probe(float* out_ss, float* out_ts, long long* cyc, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                 // [128][64]
  uint8_t* smB = smem + 16384;         // [N][64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smB + N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fill_tile(smA, 128, 0, true);
  fill_tile(smB, N, 0, false);
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc(slot, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = *slot;
  const uint32_t tD = tm, tA = tm + 256;      // D: N columns at 0; A (TMEM copy): 32 columns at 256
  // A into TMEM: lane = row, 64 bf16 = 32 packed columns (element 2c in the low half of column c)
  {
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < 32; c0 += 8) {
      uint32_t v[8];
      for (int c = 0; c < 8; ++c) {
        __nv_bfloat162 p = __floats2bfloat162_rn(aval(r, 2 * (c0 + c)), aval(r, 2 * (c0 + c) + 1));
        v[c] = *reinterpret_cast<uint32_t*>(&p);
      }
      tmem_st8(tm + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
    }
    tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t idesc = ptx::umma_idesc_bf16(128, N, 0, 0);
  const uint32_t a_addr = ptx::smem_u32(smA), b_addr = ptx::smem_u32(smB);
  uint32_t phase = 0;
  auto readback = [&](float* out) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld32(tD + ((uint32_t)(warp * 32) << 16) + c0, v);
      ptx::tmem_ld_wait(v);
      if (blockIdx.x == 0)
        for (int c = 0; c < 32; ++c) out[(size_t)(warp * 32 + lane) * N + c0 + c] = __uint_as_float(v[c]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
  };
  // ---- independent accumulators: does alternating D tiles hide the accumulate dependency? ----
  for (int nacc = 2; nacc <= 4 && nacc * N <= 256; nacc += 2) {
    long long t0 = 0;
    if (threadIdx.x == 0) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int k = (i / nacc) & 3;
        const uint64_t bd = ptx::umma_desc_sw128(b_addr + k * 32, 16, 1024);
        ptx::umma_ss(tD + (i % nacc) * N, ptx::umma_desc_sw128(a_addr + k * 32, 16, 1024), bd, idesc, 1);
      }
      ptx::umma_commit(bar);
    }
    ptx::mbar_wait(bar, phase); phase ^= 1;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[2 + nacc / 2 - 1] = clock64() - t0;
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
  }
  for (int mode = 0; mode < 2; ++mode) {      // 0 = SS, 1 = TS
    // ---- correctness: K = 64 (4 MMAs) ----
    if (threadIdx.x == 0) {
      for (int k = 0; k < 4; ++k) {
        const uint64_t bd = ptx::umma_desc_sw128(b_addr + k * 32, 16, 1024);
        if (mode == 0) ptx::umma_ss(tD, ptx::umma_desc_sw128(a_addr + k * 32, 16, 1024), bd, idesc, k != 0);
        else umma_ts(tD, tA + k * 8, bd, idesc, k != 0);
      }
      ptx::umma_commit(bar);
    }
    ptx::mbar_wait(bar, phase); phase ^= 1;
    ptx::tc_fence_after();
    readback(mode == 0 ? out_ss : out_ts);
    // ---- timing ----
    long long t0 = 0;
    if (threadIdx.x == 0) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int k = i & 3;
        const uint64_t bd = ptx::umma_desc_sw128(b_addr + k * 32, 16, 1024);
        if (mode == 0) ptx::umma_ss(tD, ptx::umma_desc_sw128(a_addr + k * 32, 16, 1024), bd, idesc, 1);
        else umma_ts(tD, tA + k * 8, bd, idesc, 1);
      }
      ptx::umma_commit(bar);
    }
    ptx::mbar_wait(bar, phase); phase ^= 1;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[mode] = clock64() - t0;
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
  }
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

template <int N>
int run(int iters) {
  float *o1, *o2; long long* cyc;
  cudaMalloc(&o1, 128 * N * 4); cudaMalloc(&o2, 128 * N * 4); cudaMalloc(&cyc, 64); cudaMemset(cyc, 0, 64);
  cudaMemset(o1, 0, 128 * N * 4); cudaMemset(o2, 0, 128 * N * 4);
  size_t smem = 1024 + 16384 + N * 128 + 64;
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<N><<<148, 128, smem>>>(o1, o2, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
  std::vector<float> h1(128 * N), h2(128 * N); long long hc[4];
  cudaMemcpy(h1.data(), o1, 128 * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h2.data(), o2, 128 * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, cyc, 32, cudaMemcpyDeviceToHost);
  int bad1 = 0, bad2 = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < 64; ++k) ref += aval(r, k) * bval(n, k);
      if (h1[r * N + n] != ref) ++bad1;
      if (h2[r * N + n] != ref) ++bad2;
    }
  printf("N=%d  SS: %d/%d wrong, %.1f cycles/MMA   TS: %d/%d wrong, %.1f cycles/MMA  (ideal %d)\n", N, bad1, 128 * N,
         (double)hc[0] / iters, bad2, 128 * N, (double)hc[1] / iters, N / 2);
  printf("      SS alternating 2 accumulators: %.1f cycles/MMA; 4 accumulators: %.1f\n", (double)hc[2] / iters, (double)hc[3] / iters);
  if (bad2) printf("   TS sample: got %.1f %.1f %.1f want-first %.1f\n", h2[0], h2[1], h2[N], h1[0]);
  return 0;
}

int main() {
  run<128>(4096);
  run<256>(4096);
  return 0;
}
