// mma_probe.cu -- cycles per tcgen05.mma (M = 128, K = 16, bf16) for the operand modes the similarity kernels
// use: A from smem (SS) or TMEM (TS), B K-major or MN-major, N = 64 / 128 / 256.  Tight warp-uniform issue loop
// (elect.sync), chains of 16 MMAs into one accumulator like the kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200/csrc/ptx.cuh"
using namespace mscs;

// mode: 0 SS K-major B, 1 TS K-major B, 2 TS MN-major B, 3 SS MN-major B
template <int N, int mode, int alt>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                 // [4][128][64]  64 KB
  uint8_t* smB = smem + 4 * 16384;     // [4][256][64]  128 KB (K-major: N rows; MN-major: [4 ch blocks][128 K rows][64])
  uint64_t* bar = reinterpret_cast<uint64_t*>(smB + 4 * 256 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc(slot, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = *slot;
  {   // A region in TMEM (columns 384..511) := small values
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = 0x3c003c00u;
    for (int c = 0; c < 128; c += 32) ptx::tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + 384 + c, v);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const bool mn = mode >= 2;
  const uint32_t idesc = ptx::umma_idesc_bf16(128, N, 0, mn ? 1 : 0);
  const uint32_t a_addr = ptx::smem_u32(smA), b_addr = ptx::smem_u32(smB);
  if (warp == 0) {
    long long t0 = clock64();
    if (ptx::elect_one()) {
      // tight issue: descriptors stepped with one 32-bit add, 16 MMAs unrolled (chains of 16 into one accumulator)
      const uint64_t ad0 = ptx::umma_desc_sw128(a_addr, 16, 1024);
      const uint64_t bd0 = mn ? ptx::umma_desc_sw128(b_addr, 16384, 1024) : ptx::umma_desc_sw128(b_addr, 16, 1024);
      const uint32_t a_lo = (uint32_t)ad0, a_hi = (uint32_t)(ad0 >> 32), b_lo = (uint32_t)bd0, b_hi = (uint32_t)(bd0 >> 32);
      for (int i = 0; i < iters; i += 16) {
        const uint32_t d = tm + ((alt && (i & 16)) ? 128 : 0);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const uint32_t boff = mn ? (uint32_t)(kk * 16 * 128) >> 4 : (uint32_t)((kk >> 2) * 256 * 128 + (kk & 3) * 32) >> 4;
          const uint32_t aoff = (uint32_t)((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
          if (mode == 0 || mode == 3) ptx::umma_ss2(d, a_lo + aoff, a_hi, b_lo + boff, b_hi, idesc, kk != 0);
          else ptx::umma_ts2(d, tm + 384 + kk * 8, b_lo + boff, b_hi, idesc, kk != 0);
        }
      }
      ptx::umma_commit(bar);
    }
    __syncwarp();
    ptx::mbar_wait(bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

template <int N, int mode, int alt>
void run(long long* out, size_t smem) {
  const char* names[4] = {"SS, B K-major ", "TS, B K-major ", "TS, B MN-major", "SS, B MN-major"};
  cudaFuncSetAttribute(probe<N, mode, alt>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemset(out, 0, 64);
  probe<N, mode, alt><<<148, 128, smem>>>(out, 8192);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("%s N=%3d %s: %.1f cycles/MMA (ideal %d)\n", names[mode], N, alt ? "2 accumulators" : "1 accumulator ", (double)h / 8192, N / 2);
}
template <int mode>
void run_mode(long long* out, size_t smem) {
  run<64, mode, 0>(out, smem); run<64, mode, 1>(out, smem);
  run<128, mode, 0>(out, smem); run<128, mode, 1>(out, smem);
  run<256, mode, 0>(out, smem);
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  size_t smem = 1024 + 4 * 16384 + 4 * 256 * 128 + 64;
  run_mode<0>(out, smem); run_mode<1>(out, smem); run_mode<2>(out, smem); run_mode<3>(out, smem);
  return 0;
}
