// gather.cu -- K2 (gather sampled embeddings + L2 normalise) and the scatter half of K4
// (normalisation backward + dense gradient).
//   gather  : features[b,:,idx] (V2.py:123) + F.normalize(p=2, dim=1, eps=1e-12) (V2.py:138)
//   scatter : autograd backward of both: dx = (dF - f (f.dF)) / max(||x||, eps), written into
//             a dense zero (n,C,h,w) gradient at the sampled pixels.
// HBM-bound byte movers: NCHW means one anchor is C scalars `plane` floats apart (one 32 B
// sector per useful 4 B), so one warp owns one anchor and keeps 8 independent loads in flight.
#include "sim_tc.cuh"
#include <stdlib.h>

namespace mscs {

constexpr int kMaxC = 256;

__global__ void __launch_bounds__(256)
k_gather_normalize(const float* __restrict__ feat, int C, int C_pad, int plane, const int* __restrict__ pix,
                   int N, int N_pad, __nv_bfloat16* __restrict__ anc_bf16, float* __restrict__ anc_f32,
                   float* __restrict__ inv_norm) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N_pad) return;
  __nv_bfloat16* orow = anc_bf16 + (size_t)row * C_pad;
  if (row >= N) {                       // zero padding rows (TMA tiles read them)
    for (int c = lane; c < C_pad; c += 32) orow[c] = __float2bfloat16(0.f);
    return;
  }
  const int gp = pix[row];
  if (gp < 0) {                         // pooled mode: the row is owned by another rank (all-reduced later)
    for (int c = lane; c < C_pad; c += 32) orow[c] = __float2bfloat16(0.f);
    return;
  }
  const int b = gp / plane, p = gp - b * plane;
  const float* src = feat + ((size_t)b * C) * plane + p;
  float v[kMaxC / 32];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxC / 32; ++j) {
    int c = lane + 32 * j;
    v[j] = (c < C) ? __ldg(src + (size_t)c * plane) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < kMaxC / 32; ++j) ss = fmaf(v[j], v[j], ss);
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0) inv_norm[row] = inv;
#pragma unroll
  for (int j = 0; j < kMaxC / 32; ++j) {
    int c = lane + 32 * j;
    float f = v[j] * inv;
    if (c < C) anc_f32[(size_t)row * C + c] = f;
    if (c < C_pad) orow[c] = __float2bfloat16(c < C ? f : 0.f);
  }
}

__global__ void __launch_bounds__(256)
k_scatter_grad(const float* __restrict__ dF, int ldF, const float* __restrict__ anc_f32,
               const float* __restrict__ inv_norm, const int* __restrict__ pix, int N, int C, int plane,
               float* __restrict__ dfeat) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const float* g = dF + (size_t)row * ldF;
  const float* f = anc_f32 + (size_t)row * C;
  float gv[kMaxC / 32], fv[kMaxC / 32];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxC / 32; ++j) {
    int c = lane + 32 * j;
    gv[j] = (c < C) ? g[c] : 0.f;
    fv[j] = (c < C) ? f[c] : 0.f;
    dot = fmaf(gv[j], fv[j], dot);
  }
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
  const bool clamped = inv >= 1e12f;        // ||x|| <= eps: F.normalize divides by the constant eps
  const int gp = pix[row];
  if (gp < 0) return;
  const int b = gp / plane, p = gp - b * plane;
  float* dst = dfeat + ((size_t)b * C) * plane + p;
#pragma unroll
  for (int j = 0; j < kMaxC / 32; ++j) {
    int c = lane + 32 * j;
    if (c < C) dst[(size_t)c * plane] = (clamped ? gv[j] : (gv[j] - fv[j] * dot)) * inv;
  }
}

// Gather in ADDRESS order: one warp per 8-pixel octet of the slot map (18% of them hold a sampled
// pixel at cfg-2).  Neighbouring warps then touch neighbouring 32-byte sectors of the same channel
// planes, so the 256 strided sector reads of an anchor hit open DRAM pages instead of random ones.
__device__ __forceinline__ void
gather_sectors_body(int block, int nblocks, const float* __restrict__ feat, int C, int C_pad, int plane,
                    const int* __restrict__ slot, int n_octets, __nv_bfloat16* __restrict__ anc_bf16,
                    float* __restrict__ anc_f32, float* __restrict__ inv_norm, const int* __restrict__ n_rows_dev) {
  const int lane = threadIdx.x & 31;
  if (n_rows_dev != nullptr && block == nblocks - 1) {
    // device-driven call: the row count is only known on the device; this (extra) block zeroes the padding
    // rows [N, N_pad) of the operand matrix that the TMA tiles read
    const int N = *n_rows_dev, N_pad = (N + 255) / 256 * 256;
    uint32_t* z = reinterpret_cast<uint32_t*>(anc_bf16 + (size_t)N * C_pad);
    for (int i = threadIdx.x; i < (N_pad - N) * (C_pad / 2); i += blockDim.x) z[i] = 0u;
    return;
  }
  const int oct = block * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (oct >= n_octets) return;
  const int gp = oct * 8;
  const int s_mine = (lane < 8) ? slot[gp + lane] : -1;
  unsigned act = __ballot_sync(0xffffffffu, s_mine >= 0);
  if (act == 0) return;
  const int b = gp / plane, p = gp - b * plane;
  const float* src0 = feat + ((size_t)b * C) * plane + p;
  while (act) {
    const int j = __ffs(act) - 1;
    act &= act - 1;
    const int row = __shfl_sync(0xffffffffu, s_mine, j);
    const float* src = src0 + j;
    float v[kMaxC / 32];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxC / 32; ++q) {
      const int c = lane + 32 * q;
      v[q] = (c < C) ? __ldg(src + (size_t)c * plane) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < kMaxC / 32; ++q) ss = fmaf(v[q], v[q], ss);
    if (anc_bf16 == nullptr) {      // raw mode (projector-tail path): the sampled rows as they are, no normalisation
#pragma unroll
      for (int q = 0; q < kMaxC / 32; ++q) {
        const int c = lane + 32 * q;
        if (c < C) anc_f32[(size_t)row * C + c] = v[q];
      }
      continue;
    }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) inv_norm[row] = inv;
    __nv_bfloat16* orow = anc_bf16 + (size_t)row * C_pad;
#pragma unroll
    for (int q = 0; q < kMaxC / 32; ++q) {
      const int c = lane + 32 * q;
      const float f = v[q] * inv;
      if (c < C) anc_f32[(size_t)row * C + c] = f;
      if (c < C_pad) orow[c] = __float2bfloat16(c < C ? f : 0.f);
    }
  }
}

__global__ void __launch_bounds__(256)
k_gather_sectors(const float* __restrict__ feat, int C, int C_pad, int plane, const int* __restrict__ slot,
                 int n_octets, __nv_bfloat16* __restrict__ anc_bf16, float* __restrict__ anc_f32,
                 float* __restrict__ inv_norm, const int* __restrict__ n_rows_dev) {
  pdl_trigger();
  gather_sectors_body(blockIdx.x, gridDim.x, feat, C, C_pad, plane, slot, n_octets, anc_bf16, anc_f32, inv_norm,
                      n_rows_dev);
}

// all scales of a call in ONE launch (the per-scale launches of the small scales do not fill the GPU and each
// costs a launch gap): block -> scale through the block prefix
struct GatherBatch {
  const float* feat[MSCS_MAX_SCALES]; const int* slot[MSCS_MAX_SCALES]; const int* n_rows_dev[MSCS_MAX_SCALES];
  __nv_bfloat16* bf16[MSCS_MAX_SCALES]; float* f32[MSCS_MAX_SCALES]; float* inv[MSCS_MAX_SCALES];
  int C[MSCS_MAX_SCALES], plane[MSCS_MAX_SCALES], n_oct[MSCS_MAX_SCALES], block0[MSCS_MAX_SCALES + 1];
  int count;
};
__global__ void __launch_bounds__(256) k_gather_sectors_batch(const __grid_constant__ GatherBatch g) {
  pdl_trigger();      // the next kernel of the forward (k_row_ranges) is launched programmatically, see common.cuh
  int s = 0;
  while (s + 1 < g.count && (int)blockIdx.x >= g.block0[s + 1]) ++s;
  gather_sectors_body(blockIdx.x - g.block0[s], g.block0[s + 1] - g.block0[s], g.feat[s], g.C[s],
                      (g.C[s] + 63) / 64 * 64, g.plane[s], g.slot[s], g.n_oct[s], g.bf16[s], g.f32[s], g.inv[s],
                      g.n_rows_dev[s]);
}

// ---------------------------------------------------------------------------------------
// Gather through TMA (round 2; opt-in, MSCS_GATHER=tma -- measured slower than the octet kernel: 0.137 against 0.094 ms
// at cfg-2, the TMA unit handles the 256 32-byte rows of a box at ~0.2 rows per clock and SM).  The octet kernel above issues, per sampled pixel, 8 load instructions whose 32 lanes
// hit 32 different channel planes: 256 L1 wavefronts per pixel -- the LSU, not DRAM, bounds it (93 us at cfg-2 with the
// stream to itself, 2.9 TB/s of sector traffic).  Here the strided walk over the channel planes is ONE bulk-tensor
// copy per 8-pixel octet: the feature map is described as a 2D tensor [n*C rows][plane columns] and a box of
// {8 pixels, C rows} (C x 32 bytes, exactly the sectors the octet kernel touches) lands in shared memory as
// tile[c][8 px]; the lanes then read 32-byte rows of it (two 128-bit loads per channel) and hold all eight pixels of
// their channels in registers.  Block = 4 warps; a block first compacts the octets of its 1024-pixel chunk of the slot
// map that hold a sampled pixel, then every warp streams its share through two tile slots (the copy of the next octet
// is in flight while the current one is normalised and stored).
// ---------------------------------------------------------------------------------------
constexpr int kTmaOctets = 128;      // octets per block chunk (1024 pixels)
struct GatherTma {
  alignas(64) CUtensorMap map[MSCS_MAX_SCALES];
  const int* slot[MSCS_MAX_SCALES]; const int* n_rows_dev[MSCS_MAX_SCALES];
  __nv_bfloat16* bf16[MSCS_MAX_SCALES]; float* f32[MSCS_MAX_SCALES]; float* inv[MSCS_MAX_SCALES];
  int C[MSCS_MAX_SCALES], plane[MSCS_MAX_SCALES], n_oct[MSCS_MAX_SCALES], block0[MSCS_MAX_SCALES + 1];
  int count;
};
__global__ void __launch_bounds__(128) k_gather_tma_batch(const __grid_constant__ GatherTma g) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  float* tiles = reinterpret_cast<float*>(smem);                       // [4 warps][2 slots][kMaxC][8]
  int* l_oct = reinterpret_cast<int*>(smem + 4 * 2 * kMaxC * 32);     // compacted octets of the chunk
  uint64_t* bars = reinterpret_cast<uint64_t*>(l_oct + kTmaOctets);   // [4 warps][2 slots]
  int* cnt = reinterpret_cast<int*>(bars + 8);
  int s = 0;
  while (s + 1 < g.count && (int)blockIdx.x >= g.block0[s + 1]) ++s;
  const int lb = (int)blockIdx.x - g.block0[s], nb = g.block0[s + 1] - g.block0[s];
  const int C = g.C[s], C_pad = (C + 63) / 64 * 64, plane = g.plane[s];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lb == nb - 1) {      // padding rows [N, N_pad) of the operand matrix (the TMA tiles of K3 read them)
    const int N = *g.n_rows_dev[s], N_pad = (N + 255) / 256 * 256;
    uint32_t* z = reinterpret_cast<uint32_t*>(g.bf16[s] + (size_t)N * C_pad);
    for (int i = threadIdx.x; i < (N_pad - N) * (C_pad / 2); i += blockDim.x) z[i] = 0u;
    return;
  }
  if (threadIdx.x == 0) {
    *cnt = 0;
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  {      // one octet per thread: keep those with a sampled pixel
    const int oct = lb * kTmaOctets + (int)threadIdx.x;
    if (oct < g.n_oct[s]) {
      const int4 a = *reinterpret_cast<const int4*>(g.slot[s] + (size_t)oct * 8);
      const int4 b = *reinterpret_cast<const int4*>(g.slot[s] + (size_t)oct * 8 + 4);
      if ((a.x & a.y & a.z & a.w & b.x & b.y & b.z & b.w) >= 0) l_oct[atomicAdd(cnt, 1)] = oct;      // any entry >= 0
    }
  }
  __syncthreads();
  const int total = *cnt;
  float* my_tiles = tiles + (size_t)warp * 2 * kMaxC * 8;
  uint64_t* my_bars = bars + warp * 2;
  const uint32_t tx = (uint32_t)C * 32u;
  auto issue = [&](int k, int slot_i) {      // lane 0: bulk copy of octet l_oct[k] into tile slot slot_i
    const int oct = l_oct[k];
    const int gp = oct * 8, b = gp / plane, p = gp - b * plane;
    ptx::mbar_expect_tx(&my_bars[slot_i], tx);
    ptx::tma_load_2d(my_tiles + (size_t)slot_i * kMaxC * 8, &g.map[s], &my_bars[slot_i], p, b * C);
  };
  int k = warp;
  if (k < total && lane == 0) issue(k, 0);
  uint32_t n_done = 0;
  for (; k < total; k += 4, ++n_done) {
    const int cur = n_done & 1;
    if (k + 4 < total && lane == 0) issue(k + 4, cur ^ 1);      // the other slot was consumed two iterations ago
    ptx::mbar_wait(&my_bars[cur], (n_done >> 1) & 1, 301);
    const int oct = l_oct[k];
    const int row_mine = lane < 8 ? g.slot[s][(size_t)oct * 8 + lane] : -1;
    const float* tile = my_tiles + (size_t)cur * kMaxC * 8;
    float v[kMaxC / 32][8];
#pragma unroll
    for (int q = 0; q < kMaxC / 32; ++q) {
      const int c = lane + 32 * q;
      if (c < C) {
        const float4 lo = *reinterpret_cast<const float4*>(tile + c * 8);
        const float4 hi = *reinterpret_cast<const float4*>(tile + c * 8 + 4);
        v[q][0] = lo.x; v[q][1] = lo.y; v[q][2] = lo.z; v[q][3] = lo.w;
        v[q][4] = hi.x; v[q][5] = hi.y; v[q][6] = hi.z; v[q][7] = hi.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[q][j] = 0.f;
      }
    }
    __syncwarp();      // every lane has read its rows: the slot may be refilled by the issue of the next iteration
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int row = __shfl_sync(0xffffffffu, row_mine, j);
      if (row < 0) continue;      // warp-uniform
      float ss = 0.f;
#pragma unroll
      for (int q = 0; q < kMaxC / 32; ++q) ss = fmaf(v[q][j], v[q][j], ss);
      ss = warp_sum(ss);
      const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
      if (lane == 0) g.inv[s][row] = inv;
      __nv_bfloat16* orow = g.bf16[s] + (size_t)row * C_pad;
#pragma unroll
      for (int q = 0; q < kMaxC / 32; ++q) {
        const int c = lane + 32 * q;
        const float f = v[q][j] * inv;
        if (c < C) g.f32[s][(size_t)row * C + c] = f;
        if (c < C_pad) orow[c] = __float2bfloat16(c < C ? f : 0.f);
      }
    }
  }
}

// slot map: slot[image*plane + pixel] = sorted anchor row sampled there, or -1
__global__ void k_slot_map(const int* __restrict__ pix, int N, int* __restrict__ slot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N && pix[i] >= 0) slot[pix[i]] = i;      // negative: row owned by another rank (pooled mode)
}

// Sector writer: the dense gradient is already zero; one warp per 8-pixel octet (= one 32-byte
// sector per channel plane) rewrites every sector that holds at least one sampled pixel with
// full-sector stores (values + explicit zeros), so no partial-sector read-modify-write reaches HBM.
// pooled mode: gradient row i was computed by rank i / rows_per_rank and is READ FROM THAT RANK'S slab over NVLink
// (n == 0: every row is local, `dF` is used)
struct PullRows { const float* p[MSCS_MAX_RANKS]; int n, rows_per_rank; };

__device__ __forceinline__ void
scatter_sectors_body(const float* __restrict__ dF, int ldF, const float* __restrict__ anc_f32,
                  const float* __restrict__ inv_norm, const int* __restrict__ slot, int n_octets, int C,
                  int plane, float* __restrict__ dfeat, int block_base, const PullRows* pull = nullptr) {
  const int lane = threadIdx.x & 31;
  const int oct = ((int)blockIdx.x - block_base) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (oct >= n_octets) return;
  const int gp = oct * 8;
  const int s_mine = (lane < 8) ? slot[gp + lane] : -1;
  if (__ballot_sync(0xffffffffu, s_mine >= 0) == 0) return;
  const int b = gp / plane, p = gp - b * plane;
  float dx[8][kMaxC / 32];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int row = __shfl_sync(0xffffffffu, s_mine, j);
    if (row >= 0) {                                   // warp-uniform
      const float* g = (pull ? pull->p[row / pull->rows_per_rank] : dF) + (size_t)row * ldF;
      if (anc_f32 == nullptr) {      // raw mode (projector-tail path): the rows ARE the pixel gradients
#pragma unroll
        for (int q = 0; q < kMaxC / 32; ++q) {
          const int c = lane + 32 * q;
          dx[j][q] = (c < C) ? g[c] : 0.f;
        }
        continue;
      }
      const float* f = anc_f32 + (size_t)row * C;
      float gv[kMaxC / 32], fv[kMaxC / 32], dot = 0.f;
#pragma unroll
      for (int q = 0; q < kMaxC / 32; ++q) {
        const int c = lane + 32 * q;
        gv[q] = (c < C) ? g[c] : 0.f;
        fv[q] = (c < C) ? f[c] : 0.f;
        dot = fmaf(gv[q], fv[q], dot);
      }
      dot = warp_sum(dot);
      const float inv = inv_norm[row];
      const bool clamped = inv >= 1e12f;
#pragma unroll
      for (int q = 0; q < kMaxC / 32; ++q) dx[j][q] = (clamped ? gv[q] : (gv[q] - fv[q] * dot)) * inv;
    } else {
#pragma unroll
      for (int q = 0; q < kMaxC / 32; ++q) dx[j][q] = 0.f;
    }
  }
#pragma unroll
  for (int q = 0; q < kMaxC / 32; ++q) {
    const int c = lane + 32 * q;
    if (c < C) {
      float4* dst = reinterpret_cast<float4*>(dfeat + ((size_t)b * C + c) * plane + p);
      dst[0] = make_float4(dx[0][q], dx[1][q], dx[2][q], dx[3][q]);
      dst[1] = make_float4(dx[4][q], dx[5][q], dx[6][q], dx[7][q]);
    }
  }
}

__global__ void __launch_bounds__(256)
k_scatter_sectors(const float* __restrict__ dF, int ldF, const float* __restrict__ anc_f32,
                  const float* __restrict__ inv_norm, const int* __restrict__ slot, int n_octets, int C,
                  int plane, float* __restrict__ dfeat, int block_base) {
  scatter_sectors_body(dF, ldF, anc_f32, inv_norm, slot, n_octets, C, plane, dfeat, block_base);
}
__global__ void __launch_bounds__(256)
k_scatter_sectors_pull(const __grid_constant__ PullRows pull, int ldF, const float* __restrict__ anc_f32,
                       const float* __restrict__ inv_norm, const int* __restrict__ slot, int n_octets, int C,
                       int plane, float* __restrict__ dfeat) {
  scatter_sectors_body(nullptr, ldF, anc_f32, inv_norm, slot, n_octets, C, plane, dfeat, 0, &pull);
}


// ---------------------------------------------------------------------------------------
// Dense gradient in ONE streaming pass -- OPT-IN alternative (MSCS_DENSE=1) to "zero-fill ahead of time + rewrite
// the sampled sectors".  The 535 MB zero fill is not really hidden (memset / copy kernels take the SMs away from
// whatever they overlap: ~70 us of step time wherever it was scheduled), so here every byte of the dense gradient
// is written exactly once, zeros and values alike.  Measured: the writer reaches only 3.3 TB/s (a plain fill: 7),
// 0.170 ms against 0.107 + ~0.065 for the default path -- a wash, so it is not the default.  (Variants tried: 8
// channel planes per thread, 4 float4 per thread with the loads hoisted, a shared-memory mask tile: all slower.)
//   k_dx_rows:      dF row -> dx row in place   (normalisation backward, N x C, tiny)
//   k_dense_write:  out[b][c][p..p+3] = slot[b][p+i] >= 0 ? dx[slot][c] : 0   (one float4 per thread)
// ---------------------------------------------------------------------------------------
struct DenseBatch {
  float* dF[MSCS_MAX_SCALES]; const float* f32[MSCS_MAX_SCALES]; const float* inv[MSCS_MAX_SCALES];
  const int* slot[MSCS_MAX_SCALES]; float* dfeat[MSCS_MAX_SCALES]; const int* n_rows_dev[MSCS_MAX_SCALES];
  int ldF[MSCS_MAX_SCALES], C[MSCS_MAX_SCALES], plane[MSCS_MAX_SCALES], n[MSCS_MAX_SCALES], rows[MSCS_MAX_SCALES];
  long long v4_0[MSCS_MAX_SCALES + 1];      // block prefix of the dense writer
  int rowblk0[MSCS_MAX_SCALES + 1];         // block prefix of the row kernel (8 rows per block)
  uint32_t* mask[MSCS_MAX_SCALES];          // 1 bit per pixel: sampled or not (built by the row kernel's extra blocks)
  int maskblk0[MSCS_MAX_SCALES + 1];        // block prefix of the mask blocks (8192 pixels per block), after the row blocks
  int count;
};

__global__ void __launch_bounds__(256) k_dx_rows(const __grid_constant__ DenseBatch g) {
  int s = 0;
  if ((int)blockIdx.x >= g.rowblk0[g.count]) {
    // ---- mask blocks: one 32-bit word per 32 pixels of the slot map (the dense writer then reads 1 bit per pixel
    // instead of a 4-byte slot entry per pixel and channel) ----
    const int mb = (int)blockIdx.x - g.rowblk0[g.count];
    while (s + 1 < g.count && mb >= g.maskblk0[s + 1]) ++s;
    const unsigned npix = (unsigned)g.n[s] * (unsigned)g.plane[s];
    const unsigned w = (unsigned)(mb - g.maskblk0[s]) * 256u + threadIdx.x;      // word index
    if (w * 32u >= npix) return;
    uint32_t bits = 0;
    const int* sl = g.slot[s] + (size_t)w * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (w * 32u + i * 4u < npix) {          // npix is a multiple of 4
        const int4 q = __ldg(reinterpret_cast<const int4*>(sl) + i);
        bits |= (uint32_t)(q.x >= 0) << (4 * i) | (uint32_t)(q.y >= 0) << (4 * i + 1) |
                (uint32_t)(q.z >= 0) << (4 * i + 2) | (uint32_t)(q.w >= 0) << (4 * i + 3);
      }
    }
    g.mask[s][w] = bits;
    return;
  }
  while (s + 1 < g.count && (int)blockIdx.x >= g.rowblk0[s + 1]) ++s;
  const int lane = threadIdx.x & 31;
  const int row = ((int)blockIdx.x - g.rowblk0[s]) * 8 + (threadIdx.x >> 5);
  if (row >= g.rows[s]) return;
  const int C = g.C[s];
  float* gr = g.dF[s] + (size_t)row * g.ldF[s];
  const float* f = g.f32[s] + (size_t)row * C;
  float gv[kMaxC / 32], fv[kMaxC / 32], dot = 0.f;
#pragma unroll
  for (int q = 0; q < kMaxC / 32; ++q) {
    const int c = lane + 32 * q;
    gv[q] = (c < C) ? gr[c] : 0.f;
    fv[q] = (c < C) ? f[c] : 0.f;
    dot = fmaf(gv[q], fv[q], dot);
  }
  dot = warp_sum(dot);
  const float inv = g.inv[s][row];
  const bool clamped = inv >= 1e12f;        // ||x|| <= eps: F.normalize divides by the constant eps
#pragma unroll
  for (int q = 0; q < kMaxC / 32; ++q) {
    const int c = lane + 32 * q;
    if (c < C) gr[c] = (clamped ? gv[q] : (gv[q] - fv[q] * dot)) * inv;
  }
}

// Purely sequential writer (the order a memset would use: interleaving several channel planes per block measured
// slower): block = 2048 consecutive floats of the flattened [b][c][pixel] gradient of ONE scale, thread = two float4.
// One mask bit per pixel says whether anything was sampled there (97.5 % of the float4 at cfg-2 are plain zero
// stores); only then the slot entries and the dx values are fetched.  32-bit index arithmetic; v4_0 holds BLOCK
// prefixes per scale.
__global__ void __launch_bounds__(256) k_dense_write(const __grid_constant__ DenseBatch g) {
  int s = 0;
  while (s + 1 < g.count && (long long)blockIdx.x >= g.v4_0[s + 1]) ++s;
  const unsigned blk = blockIdx.x - (unsigned)g.v4_0[s];
  const unsigned plane = (unsigned)g.plane[s], C = (unsigned)g.C[s];
  const unsigned total = (unsigned)g.n[s] * C * plane;                          // floats of this scale (< 2^32)
  const uint32_t* __restrict__ mask = g.mask[s];
  float* __restrict__ out = g.dfeat[s];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const unsigned f = blk * 2048u + (u * 256u + threadIdx.x) * 4u;
    if (f >= total) return;
    const unsigned bc = f / plane, p = f - bc * plane;
    const unsigned b = bc / C, c = bc - b * C;
    const unsigned lin = b * plane + p;
    const uint32_t m = (__ldg(mask + (lin >> 5)) >> (lin & 31u)) & 0xFu;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m) {
      const int4 sl = *reinterpret_cast<const int4*>(g.slot[s] + lin);
      const float* __restrict__ dx = g.dF[s];
      const int ld = g.ldF[s];
      if (sl.x >= 0) v.x = dx[(size_t)sl.x * ld + c];
      if (sl.y >= 0) v.y = dx[(size_t)sl.y * ld + c];
      if (sl.z >= 0) v.z = dx[(size_t)sl.z * ld + c];
      if (sl.w >= 0) v.w = dx[(size_t)sl.w * ld + c];
    }
    __stcs(reinterpret_cast<float4*>(out + f), v);
  }
}


// ---------------------------------------------------------------------------------------
// Channels-last feature maps (memory order [n][h][w][C], e.g. a projector run in torch.channels_last): an anchor is
// ONE contiguous row of C floats instead of C sectors `plane` floats apart, so gather and scatter become plain row
// copies (33 MB instead of 354 MB of DRAM reads at cfg-2).  Warp per sorted anchor row; lane l handles the channels
// l + 32 q -- the same lane/channel assignment as the NCHW kernels, so the results are bit-identical.
// ---------------------------------------------------------------------------------------
struct RowsBatch {
  const float* feat[MSCS_MAX_SCALES]; const int* pix[MSCS_MAX_SCALES]; const int* n_rows_dev[MSCS_MAX_SCALES];
  __nv_bfloat16* bf16[MSCS_MAX_SCALES]; float* f32[MSCS_MAX_SCALES]; float* inv[MSCS_MAX_SCALES];
  const float* dF[MSCS_MAX_SCALES]; float* dfeat[MSCS_MAX_SCALES];
  int C[MSCS_MAX_SCALES], ldF[MSCS_MAX_SCALES], rows[MSCS_MAX_SCALES], block0[MSCS_MAX_SCALES + 1];
  int count;
};

__global__ void __launch_bounds__(256) k_gather_rows_nhwc(const __grid_constant__ RowsBatch g) {
  pdl_trigger();
  int s = 0;
  while (s + 1 < g.count && (int)blockIdx.x >= g.block0[s + 1]) ++s;
  const int lane = threadIdx.x & 31;
  const int row = ((int)blockIdx.x - g.block0[s]) * 8 + (threadIdx.x >> 5);
  const int N = g.n_rows_dev[s] ? *g.n_rows_dev[s] : g.rows[s], N_pad = (N + 255) / 256 * 256;
  if (row >= N_pad) return;
  const int C = g.C[s], C_pad = (C + 63) / 64 * 64;
  __nv_bfloat16* orow = g.bf16[s] + (size_t)row * C_pad;
  const int gp = row < N ? g.pix[s][row] : -1;
  if (gp < 0) {                         // padding row, or (pooled mode) a row owned by another rank
    for (int c = lane; c < C_pad; c += 32) orow[c] = __float2bfloat16(0.f);
    return;
  }
  const float* src = g.feat[s] + (size_t)gp * C;
  float v[kMaxC / 32];
  float ss = 0.f;
#pragma unroll
  for (int q = 0; q < kMaxC / 32; ++q) {
    const int c = lane + 32 * q;
    v[q] = (c < C) ? __ldg(src + c) : 0.f;
  }
#pragma unroll
  for (int q = 0; q < kMaxC / 32; ++q) ss = fmaf(v[q], v[q], ss);
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0) g.inv[s][row] = inv;
#pragma unroll
  for (int q = 0; q < kMaxC / 32; ++q) {
    const int c = lane + 32 * q;
    const float f = v[q] * inv;
    if (c < C) g.f32[s][(size_t)row * C + c] = f;
    if (c < C_pad) orow[c] = __float2bfloat16(c < C ? f : 0.f);
  }
}

// normalisation backward + row store into the (pre-zeroed) channels-last dense gradient
__global__ void __launch_bounds__(256) k_scatter_rows_nhwc(const __grid_constant__ RowsBatch g) {
  int s = 0;
  while (s + 1 < g.count && (int)blockIdx.x >= g.block0[s + 1]) ++s;
  const int lane = threadIdx.x & 31;
  const int row = ((int)blockIdx.x - g.block0[s]) * 8 + (threadIdx.x >> 5);
  if (row >= g.rows[s]) return;
  const int gp = g.pix[s][row];
  if (gp < 0) return;
  const int C = g.C[s];
  const float* gr = g.dF[s] + (size_t)row * g.ldF[s];
  const float* f = g.f32[s] + (size_t)row * C;
  float gv[kMaxC / 32], fv[kMaxC / 32], dot = 0.f;
#pragma unroll
  for (int q = 0; q < kMaxC / 32; ++q) {
    const int c = lane + 32 * q;
    gv[q] = (c < C) ? gr[c] : 0.f;
    fv[q] = (c < C) ? f[c] : 0.f;
    dot = fmaf(gv[q], fv[q], dot);
  }
  dot = warp_sum(dot);
  const float inv = g.inv[s][row];
  const bool clamped = inv >= 1e12f;        // ||x|| <= eps: F.normalize divides by the constant eps
  float* dst = g.dfeat[s] + (size_t)gp * C;
#pragma unroll
  for (int q = 0; q < kMaxC / 32; ++q) {
    const int c = lane + 32 * q;
    if (c < C) dst[c] = (clamped ? gv[q] : (gv[q] - fv[q] * dot)) * inv;
  }
}

// One-pass dense writer, second form (round 2): a thread OWNS 4 consecutive pixels and walks the channel planes.
// Its four slot entries are fetched once (one coalesced int4), so the sampled / not-sampled decision is loop
// invariant: 97 % of the threads run a pure store loop, the others add up to four independent (predicated) loads of
// dx[row][c] per channel.  A block writes 2 KB contiguous per channel plane (512 pixels), i.e. whole DRAM pages in
// linear order per plane -- the order of the memset it replaces, without its second pass over the sampled sectors.
struct DenseStream {
  const float* dx[MSCS_MAX_SCALES]; const int* slot[MSCS_MAX_SCALES]; float* dfeat[MSCS_MAX_SCALES];
  int ld[MSCS_MAX_SCALES], C[MSCS_MAX_SCALES], plane[MSCS_MAX_SCALES], npix[MSCS_MAX_SCALES], blk0[MSCS_MAX_SCALES + 1];
  int count;
};
__global__ void __launch_bounds__(128) k_dense_stream(const __grid_constant__ DenseStream g) {
  int s = 0;
  while (s + 1 < g.count && (int)blockIdx.x >= g.blk0[s + 1]) ++s;
  const int lin = (((int)blockIdx.x - g.blk0[s]) * 128 + (int)threadIdx.x) * 4;
  if (lin >= g.npix[s]) return;
  const int plane = g.plane[s], C = g.C[s], ld = g.ld[s];
  const int4 sl = *reinterpret_cast<const int4*>(g.slot[s] + lin);
  const int b = lin / plane, p = lin - b * plane;
  float* __restrict__ out = g.dfeat[s] + ((size_t)b * C) * plane + p;
  const float* __restrict__ dx = g.dx[s];
  const float* r0 = dx + (size_t)max(sl.x, 0) * ld;
  const float* r1 = dx + (size_t)max(sl.y, 0) * ld;
  const float* r2 = dx + (size_t)max(sl.z, 0) * ld;
  const float* r3 = dx + (size_t)max(sl.w, 0) * ld;
  const bool h0 = sl.x >= 0, h1 = sl.y >= 0, h2 = sl.z >= 0, h3 = sl.w >= 0;
#pragma unroll 8
  for (int c = 0; c < C; ++c) {
    float4 v;
    v.x = h0 ? __ldg(r0 + c) : 0.f;
    v.y = h1 ? __ldg(r1 + c) : 0.f;
    v.z = h2 ? __ldg(r2 + c) : 0.f;
    v.w = h3 ? __ldg(r3 + c) : 0.f;
    __stcs(reinterpret_cast<float4*>(out + (size_t)c * plane), v);
  }
}

}  // namespace mscs

using namespace mscs;

extern "C" int mscs_gather_normalize_sectors(const float* feat, int n, int C, int plane, const int32_t* slot, int N,
                                             void* anc_bf16, float* anc_f32, float* inv_norm, void* stream_) {
  MSCS_CHECK_ARG(feat && slot && anc_bf16 && anc_f32 && inv_norm, "null pointer argument");
  MSCS_CHECK_ARG(C >= 1 && C <= kMaxC, "C=%d unsupported (1..%d)", C, kMaxC);
  MSCS_CHECK_ARG(n >= 1 && plane >= 8 && plane % 8 == 0 && N >= 1, "bad sizes (plane must be a multiple of 8)");
  cudaStream_t st = (cudaStream_t)stream_;
  const int C_pad = (C + 63) / 64 * 64, N_pad = (N + 255) / 256 * 256;
  if (N_pad > N)      // zero padding rows of the operand matrix (TMA tiles read them)
    MSCS_CUDA(cudaMemsetAsync((__nv_bfloat16*)anc_bf16 + (size_t)N * C_pad, 0,
                              sizeof(__nv_bfloat16) * (size_t)(N_pad - N) * C_pad, st));
  const int n_oct = n * (plane / 8);
  k_gather_sectors<<<ceil_div(n_oct, 8), 256, 0, st>>>(feat, C, C_pad, plane, slot, n_oct, (__nv_bfloat16*)anc_bf16,
                                                       anc_f32, inv_norm, nullptr);
  MSCS_LAUNCH_CHECK();
  return 0;
}

// Projector-tail path: the sampled pixels' c_in-vectors as they are (no normalisation), rows in
// sorted-anchor order, and the matching scatter of row gradients into the pre-zeroed dense (n, c_in, h, w) gradient.
extern "C" int mscs_gather_rows_raw(const float* feat, int n, int C, int plane, const int32_t* slot, float* rows,
                                    void* stream_) {
  MSCS_CHECK_ARG(feat && slot && rows, "null pointer argument");
  MSCS_CHECK_ARG(C >= 1 && C <= kMaxC, "C=%d unsupported (1..%d)", C, kMaxC);
  MSCS_CHECK_ARG(n >= 1 && plane >= 8 && plane % 8 == 0, "bad sizes (plane must be a multiple of 8)");
  const int n_oct = n * (plane / 8);
  k_gather_sectors<<<ceil_div(n_oct, 8), 256, 0, (cudaStream_t)stream_>>>(feat, C, (C + 63) / 64 * 64, plane, slot, n_oct,
                                                                          nullptr, rows, nullptr, nullptr);
  MSCS_LAUNCH_CHECK();
  return 0;
}
extern "C" int mscs_scatter_rows_raw(const float* drows, int ld, const int32_t* slot, int n, int C, int plane,
                                     float* dfeat, void* stream_) {
  MSCS_CHECK_ARG(drows && slot && dfeat, "null pointer argument");
  MSCS_CHECK_ARG(C >= 1 && C <= kMaxC && ld >= C, "C=%d / ld=%d unsupported", C, ld);
  MSCS_CHECK_ARG(plane % 8 == 0, "plane %d is not a multiple of 8 pixels", plane);
  const int n_oct = n * (plane / 8);
  k_scatter_sectors<<<ceil_div(n_oct, 8), 256, 0, (cudaStream_t)stream_>>>(drows, ld, nullptr, nullptr, slot, n_oct, C,
                                                                           plane, dfeat, 0);
  MSCS_LAUNCH_CHECK();
  return 0;
}

// Same, with the number of sampled rows read from device memory (`n_rows_dev`, e.g. &plan_dev[s].N): no host
// knowledge of the sampling result is needed to enqueue it.
extern "C" int mscs_gather_normalize_sectors_async(const float* feat, int n, int C, int plane, const int32_t* slot,
                                                   const int32_t* n_rows_dev, void* anc_bf16, float* anc_f32,
                                                   float* inv_norm, void* stream_) {
  MSCS_CHECK_ARG(feat && slot && n_rows_dev && anc_bf16 && anc_f32 && inv_norm, "null pointer argument");
  MSCS_CHECK_ARG(C >= 1 && C <= kMaxC, "C=%d unsupported (1..%d)", C, kMaxC);
  MSCS_CHECK_ARG(n >= 1 && plane >= 8 && plane % 8 == 0, "bad sizes (plane must be a multiple of 8)");
  cudaStream_t st = (cudaStream_t)stream_;
  const int C_pad = (C + 63) / 64 * 64;
  const int n_oct = n * (plane / 8);
  k_gather_sectors<<<ceil_div(n_oct, 8) + 1, 256, 0, st>>>(feat, C, C_pad, plane, slot, n_oct, (__nv_bfloat16*)anc_bf16,
                                                           anc_f32, inv_norm, n_rows_dev);
  MSCS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mscs_slot_map(const int32_t* pix, int N, int n_pixels, int32_t* slot, void* stream_) {
  MSCS_CHECK_ARG(pix && slot && N >= 1 && n_pixels >= 1, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream_;
  MSCS_CUDA(cudaMemsetAsync(slot, 0xff, sizeof(int) * (size_t)n_pixels, st));
  k_slot_map<<<ceil_div(N, 256), 256, 0, st>>>(pix, N, slot);
  MSCS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mscs_scatter_sectors(const float* dF, int ldF, const float* anc_f32, const float* inv_norm,
                                    const int32_t* slot, int n, int C, int plane, float* dfeat, void* stream_) {
  MSCS_CHECK_ARG(dF && anc_f32 && inv_norm && slot && dfeat, "null pointer argument");
  MSCS_CHECK_ARG(C >= 1 && C <= kMaxC && ldF >= C, "C=%d / ldF=%d unsupported", C, ldF);
  MSCS_CHECK_ARG(plane % 8 == 0, "plane %d is not a multiple of 8 pixels: use mscs_scatter_grad", plane);
  const int n_oct = n * (plane / 8);
  k_scatter_sectors<<<ceil_div(n_oct, 8), 256, 0, (cudaStream_t)stream_>>>(dF, ldF, anc_f32, inv_norm, slot, n_oct,
                                                                           C, plane, dfeat, 0);
  MSCS_LAUNCH_CHECK();
  return 0;
}

// pooled mode: the same scatter with every gradient row read from the rank that computed it (row / rows_per_rank)
extern "C" int mscs_scatter_sectors_pull(void* const* slabs, int world, size_t dF_byte_off, int rows_per_rank, int ldF,
                                         const float* anc_f32, const float* inv_norm, const int32_t* slot, int n, int C,
                                         int plane, float* dfeat, void* stream_) {
  MSCS_CHECK_ARG(slabs && anc_f32 && inv_norm && slot && dfeat, "null pointer argument");
  MSCS_CHECK_ARG(world >= 1 && world <= MSCS_MAX_RANKS && rows_per_rank >= 1, "bad world %d / rows per rank %d", world,
                 rows_per_rank);
  MSCS_CHECK_ARG(C >= 1 && C <= kMaxC && ldF >= C, "C=%d / ldF=%d unsupported", C, ldF);
  MSCS_CHECK_ARG(plane % 8 == 0, "plane %d is not a multiple of 8 pixels", plane);
  PullRows pull{};
  pull.n = world; pull.rows_per_rank = rows_per_rank;
  for (int r = 0; r < world; ++r) {
    MSCS_CHECK_ARG(slabs[r] != nullptr, "peer %d: null slab pointer", r);
    pull.p[r] = reinterpret_cast<const float*>((const char*)slabs[r] + dF_byte_off);
  }
  const int n_oct = n * (plane / 8);
  k_scatter_sectors_pull<<<ceil_div(n_oct, 8), 256, 0, (cudaStream_t)stream_>>>(pull, ldF, anc_f32, inv_norm, slot,
                                                                                n_oct, C, plane, dfeat);
  MSCS_LAUNCH_CHECK();
  return 0;
}

// One launch for every scale of a call (see mscs.h)
extern "C" int mscs_gather_normalize_sectors_batch(const mscs_gather_item* items, int count, void* stream_) {
  MSCS_CHECK_ARG(items && count >= 1 && count <= MSCS_MAX_SCALES, "bad item count %d", count);
  GatherBatch g{};
  g.count = count;
  int blocks = 0;
  for (int s = 0; s < count; ++s) {
    const mscs_gather_item& it = items[s];
    MSCS_CHECK_ARG(it.feat && it.slot && it.n_rows_dev && it.anc_bf16 && it.anc_f32 && it.inv_norm,
                   "item %d: null pointer argument", s);
    MSCS_CHECK_ARG(it.C >= 1 && it.C <= kMaxC, "item %d: C=%d unsupported (1..%d)", s, it.C, kMaxC);
    MSCS_CHECK_ARG(it.n >= 1 && it.plane >= 8 && it.plane % 8 == 0, "item %d: plane must be a multiple of 8", s);
    g.feat[s] = it.feat; g.slot[s] = it.slot; g.n_rows_dev[s] = it.n_rows_dev;
    g.bf16[s] = (__nv_bfloat16*)it.anc_bf16; g.f32[s] = it.anc_f32; g.inv[s] = it.inv_norm;
    g.C[s] = it.C; g.plane[s] = it.plane; g.n_oct[s] = it.n * (it.plane / 8);
  }
  for (int s = 0; s < count; ++s) {
    g.block0[s] = blocks;
    blocks += ceil_div(g.n_oct[s], 8) + 1;      // + the block that zeroes the padding rows
  }
  g.block0[count] = blocks;
  k_gather_sectors_batch<<<blocks, 256, 0, (cudaStream_t)stream_>>>(g);
  MSCS_LAUNCH_CHECK();
  return 0;
}

// the same gather through bulk-tensor copies (k_gather_tma_batch)
extern "C" int mscs_gather_normalize_tma_batch(const mscs_gather_item* items, int count, void* stream_) {
  MSCS_CHECK_ARG(items && count >= 1 && count <= MSCS_MAX_SCALES, "bad item count %d", count);
  GatherTma g{};
  g.count = count;
  int blocks = 0;
  for (int s = 0; s < count; ++s) {
    const mscs_gather_item& it = items[s];
    MSCS_CHECK_ARG(it.feat && it.slot && it.n_rows_dev && it.anc_bf16 && it.anc_f32 && it.inv_norm,
                   "item %d: null pointer argument", s);
    MSCS_CHECK_ARG(it.C >= 1 && it.C <= kMaxC, "item %d: C=%d unsupported (1..%d)", s, it.C, kMaxC);
    MSCS_CHECK_ARG(it.n >= 1 && it.plane >= 8 && it.plane % 8 == 0, "item %d: plane must be a multiple of 8", s);
    MSCS_CHECK_ARG((uintptr_t)it.feat % 16 == 0, "item %d: feature map must be 16-byte aligned", s);
    if (make_tensor_map_2d(&g.map[s], (int)CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, it.feat, (unsigned long long)it.plane,
                           (unsigned long long)it.n * it.C, (unsigned long long)it.plane * 4, 8, (unsigned)it.C))
      return -1;
    g.slot[s] = it.slot; g.n_rows_dev[s] = it.n_rows_dev;
    g.bf16[s] = (__nv_bfloat16*)it.anc_bf16; g.f32[s] = it.anc_f32; g.inv[s] = it.inv_norm;
    g.C[s] = it.C; g.plane[s] = it.plane; g.n_oct[s] = it.n * (it.plane / 8);
    g.block0[s] = blocks;
    blocks += ceil_div(g.n_oct[s], kTmaOctets) + 1;      // + the block that zeroes the padding rows
  }
  g.block0[count] = blocks;
  const size_t smem = 128 + (size_t)4 * 2 * kMaxC * 32 + sizeof(int) * kTmaOctets + 8 * sizeof(uint64_t) + 16;
  static bool attr_done = false;
  if (!attr_done) {
    MSCS_CUDA(cudaFuncSetAttribute(k_gather_tma_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  k_gather_tma_batch<<<blocks, 128, smem, (cudaStream_t)stream_>>>(g);
  MSCS_LAUNCH_CHECK();
  return 0;
}

struct ScatterBatch {
  const float* dF[MSCS_MAX_SCALES]; const float* f32[MSCS_MAX_SCALES]; const float* inv[MSCS_MAX_SCALES];
  const int* slot[MSCS_MAX_SCALES]; float* dfeat[MSCS_MAX_SCALES];
  int ldF[MSCS_MAX_SCALES], C[MSCS_MAX_SCALES], plane[MSCS_MAX_SCALES], n_oct[MSCS_MAX_SCALES], block0[MSCS_MAX_SCALES + 1];
  int count;
};
__global__ void __launch_bounds__(256) k_scatter_sectors_batch(const __grid_constant__ ScatterBatch g) {
  int s = 0;
  while (s + 1 < g.count && (int)blockIdx.x >= g.block0[s + 1]) ++s;
  scatter_sectors_body(g.dF[s], g.ldF[s], g.f32[s], g.inv[s], g.slot[s], g.n_oct[s], g.C[s], g.plane[s], g.dfeat[s],
                       g.block0[s]);
}

extern "C" int mscs_scatter_sectors_batch(const mscs_scatter_item* items, int count, void* stream_) {
  MSCS_CHECK_ARG(items && count >= 1 && count <= MSCS_MAX_SCALES, "bad item count %d", count);
  ScatterBatch g{};
  g.count = count;
  int blocks = 0;
  for (int s = 0; s < count; ++s) {
    const mscs_scatter_item& it = items[s];
    MSCS_CHECK_ARG(it.dF && it.anc_f32 && it.inv_norm && it.slot && it.dfeat, "item %d: null pointer argument", s);
    MSCS_CHECK_ARG(it.C >= 1 && it.C <= kMaxC && it.ldF >= it.C, "item %d: C=%d / ldF=%d unsupported", s, it.C, it.ldF);
    MSCS_CHECK_ARG(it.plane % 8 == 0, "item %d: plane %d is not a multiple of 8 pixels", s, it.plane);
    g.dF[s] = it.dF; g.f32[s] = it.anc_f32; g.inv[s] = it.inv_norm; g.slot[s] = it.slot; g.dfeat[s] = it.dfeat;
    g.ldF[s] = it.ldF; g.C[s] = it.C; g.plane[s] = it.plane; g.n_oct[s] = it.n * (it.plane / 8);
    g.block0[s] = blocks;
    blocks += ceil_div(g.n_oct[s], 8);
  }
  g.block0[count] = blocks;
  k_scatter_sectors_batch<<<blocks, 256, 0, (cudaStream_t)stream_>>>(g);
  MSCS_LAUNCH_CHECK();
  return 0;
}


extern "C" int mscs_gather_normalize(const float* feat, int n, int C, int plane, const int32_t* pix, int N,
                                     void* anc_bf16, float* anc_f32, float* inv_norm, void* stream_) {
  MSCS_CHECK_ARG(feat && pix && anc_bf16 && anc_f32 && inv_norm, "null pointer argument");
  MSCS_CHECK_ARG(C >= 1 && C <= kMaxC, "C=%d unsupported (1..%d)", C, kMaxC);
  MSCS_CHECK_ARG(n >= 1 && plane >= 1 && N >= 1, "bad sizes");
  const int C_pad = (C + 63) / 64 * 64, N_pad = (N + 255) / 256 * 256;
  k_gather_normalize<<<ceil_div(N_pad, 8), 256, 0, (cudaStream_t)stream_>>>(
      feat, C, C_pad, plane, pix, N, N_pad, (__nv_bfloat16*)anc_bf16, anc_f32, inv_norm);
  MSCS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mscs_scatter_grad(const float* dF, int ldF, const float* anc_f32, const float* inv_norm,
                                 const int32_t* pix, int N, int n, int C, int plane, float* dfeat,
                                 int zero_fill, void* stream_) {
  MSCS_CHECK_ARG(dF && anc_f32 && inv_norm && pix && dfeat, "null pointer argument");
  MSCS_CHECK_ARG(C >= 1 && C <= kMaxC && ldF >= C, "C=%d / ldF=%d unsupported", C, ldF);
  cudaStream_t st = (cudaStream_t)stream_;
  if (zero_fill) MSCS_CUDA(cudaMemsetAsync(dfeat, 0, sizeof(float) * (size_t)n * C * plane, st));
  k_scatter_grad<<<ceil_div(N, 8), 256, 0, st>>>(dF, ldF, anc_f32, inv_norm, pix, N, C, plane, dfeat);
  MSCS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mscs_scatter_dense_batch(const mscs_scatter_item* items, const int32_t* rows, int count,
                                        uint32_t* mask_scratch, void* stream_) {
  MSCS_CHECK_ARG(items && rows && mask_scratch && count >= 1 && count <= MSCS_MAX_SCALES, "bad arguments");
  DenseBatch g{};
  g.count = count;
  long long v4 = 0;
  int rb = 0, mblk = 0;
  size_t mask_words = 0;
  for (int s = 0; s < count; ++s) {
    const mscs_scatter_item& it = items[s];
    MSCS_CHECK_ARG(it.dF && it.anc_f32 && it.inv_norm && it.slot && it.dfeat, "item %d: null pointer argument", s);
    MSCS_CHECK_ARG(it.C >= 1 && it.C <= kMaxC && it.ldF >= it.C, "item %d: C=%d / ldF=%d unsupported", s, it.C, it.ldF);
    MSCS_CHECK_ARG(it.plane % 4 == 0 && rows[s] >= 0, "item %d: plane %d is not a multiple of 4 pixels", s, it.plane);
    g.dF[s] = const_cast<float*>(it.dF); g.f32[s] = it.anc_f32; g.inv[s] = it.inv_norm; g.slot[s] = it.slot;
    g.dfeat[s] = it.dfeat; g.ldF[s] = it.ldF; g.C[s] = it.C; g.plane[s] = it.plane; g.n[s] = it.n; g.rows[s] = rows[s];
    MSCS_CHECK_ARG((long long)it.n * it.C * it.plane < (1ll << 32), "item %d: more than 2^32 gradient elements", s);
    g.v4_0[s] = v4;      // block prefix: 2048 floats per block
    v4 += ((long long)it.n * it.C * it.plane + 2047) / 2048;
    g.mask[s] = mask_scratch + mask_words;
    g.maskblk0[s] = mblk;
    { const long long words = ((long long)it.n * it.plane + 31) / 32;
      mask_words += (size_t)words; mblk += (int)((words + 255) / 256); }
    g.rowblk0[s] = rb; rb += ceil_div(rows[s], 8);
  }
  g.v4_0[count] = v4; g.rowblk0[count] = rb; g.maskblk0[count] = mblk;
  cudaStream_t st = (cudaStream_t)stream_;
  static const int writer = [] { const char* e = getenv("MSCS_DENSE_WRITER"); return e ? atoi(e) : 1; }();
  if (writer == 0) {      // first form (round 1): linear float4 walk with a pixel mask
    k_dx_rows<<<rb + mblk, 256, 0, st>>>(g);
    MSCS_LAUNCH_CHECK();
    k_dense_write<<<(unsigned)v4, 256, 0, st>>>(g);
    MSCS_LAUNCH_CHECK();
    return 0;
  }
  g.maskblk0[count] = 0;      // no mask blocks: k_dx_rows only turns the gradient rows into dx rows
  k_dx_rows<<<rb > 0 ? rb : 1, 256, 0, st>>>(g);
  MSCS_LAUNCH_CHECK();
  DenseStream d{};
  d.count = count;
  int blk = 0;
  for (int s = 0; s < count; ++s) {
    d.dx[s] = g.dF[s]; d.slot[s] = g.slot[s]; d.dfeat[s] = g.dfeat[s]; d.ld[s] = g.ldF[s]; d.C[s] = g.C[s];
    d.plane[s] = g.plane[s]; d.npix[s] = g.n[s] * g.plane[s]; d.blk0[s] = blk;
    blk += ceil_div(d.npix[s], 512);
  }
  d.blk0[count] = blk;
  k_dense_stream<<<blk, 128, 0, st>>>(d);
  MSCS_LAUNCH_CHECK();
  return 0;
}

// ---- channels-last (NHWC) row gather / scatter, every scale in one launch (see mscs.h) ----
static int rows_batch(const mscs_rows_item* items, int count, bool gather, RowsBatch* g, int* blocks_out) {
  MSCS_CHECK_ARG(items && count >= 1 && count <= MSCS_MAX_SCALES, "bad item count %d", count);
  g->count = count;
  int blocks = 0;
  for (int s = 0; s < count; ++s) {
    const mscs_rows_item& it = items[s];
    MSCS_CHECK_ARG(it.pix && it.anc_f32 && it.inv_norm, "item %d: null pointer argument", s);
    MSCS_CHECK_ARG(it.C >= 1 && it.C <= kMaxC && it.rows >= 0, "item %d: C=%d unsupported (1..%d)", s, it.C, kMaxC);
    if (gather) MSCS_CHECK_ARG(it.feat && it.anc_bf16, "item %d: null pointer argument", s);
    else MSCS_CHECK_ARG(it.dF && it.dfeat && it.ldF >= it.C, "item %d: bad gradient arguments", s);
    g->feat[s] = it.feat; g->pix[s] = it.pix; g->n_rows_dev[s] = it.n_rows_dev;
    g->bf16[s] = (__nv_bfloat16*)it.anc_bf16; g->f32[s] = it.anc_f32; g->inv[s] = it.inv_norm;
    g->dF[s] = it.dF; g->dfeat[s] = it.dfeat; g->C[s] = it.C; g->ldF[s] = it.ldF; g->rows[s] = it.rows;
    g->block0[s] = blocks;
    blocks += ceil_div(gather ? (it.rows + 255) / 256 * 256 : it.rows, 8);
  }
  g->block0[count] = blocks;
  *blocks_out = blocks;
  return 0;
}

extern "C" int mscs_gather_rows_nhwc_batch(const mscs_rows_item* items, int count, void* stream_) {
  RowsBatch g{};
  int blocks = 0;
  if (int rc = rows_batch(items, count, true, &g, &blocks)) return rc;
  if (blocks == 0) return 0;
  k_gather_rows_nhwc<<<blocks, 256, 0, (cudaStream_t)stream_>>>(g);
  MSCS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mscs_scatter_rows_nhwc_batch(const mscs_rows_item* items, int count, void* stream_) {
  RowsBatch g{};
  int blocks = 0;
  if (int rc = rows_batch(items, count, false, &g, &blocks)) return rc;
  if (blocks == 0) return 0;
  k_scatter_rows_nhwc<<<blocks, 256, 0, (cudaStream_t)stream_>>>(g);
  MSCS_LAUNCH_CHECK();
  return 0;
}
