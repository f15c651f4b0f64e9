// sim_fwd.cu -- K3: fused similarity + loss forward on tcgen05 / TMEM, TMA-fed.
//
// Replaces, for every term of a call at once (V2.py:127-192, _ms.py:84-161):
//   normalised anchors x keys^T / tau, the positive / negative masks, exp, the row sums and the
//   per-pair log-probabilities -- without ever materialising the N1 x N2 logits.
// Two sweeps of the same kernel (the positive terms need the complete negative sums):
//   MODE 0  all column tiles:        neg_i  = sum_{y_j != y_i} exp(l_ij)
//   MODE 1  class-diagonal tiles:    pos_i  = sum_{j in P_i} [l_ij - log(exp(l_ij) + neg_i)],
//                                    S_i    = sum_{j in P_i} 1/(exp(l_ij) + neg_i)
//
// CTA = 256 anchor rows (two 128-row UMMA halves, resident in smem) x a run of 128-key tiles
// streamed through a TMA ring; accumulators double-buffered in TMEM (2 x 2 x 128 columns).
// Warp roles: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4..11 epilogue (one thread per
// anchor row: row sums stay in a register for the whole run, one atomicAdd per row at the end).
#include "sim_tc.cuh"

namespace mscs {

constexpr int kFwdRows = 256;
constexpr int kFwdThreads = 384;
constexpr int kFwdStages = 5;

struct FwdTerm {
  const int* a_cls; const int* k_seg;
  float* neg; float* pos; float* ssum;
  int N1, N2, self_mask, a_map, k_map;
  float scale_log2;
};
struct FwdArgs {
  alignas(64) CUtensorMap maps[MSCS_MAX_SCALES];
  FwdTerm t[MSCS_MAX_TERMS];
  WorkTable work;
};

__host__ __device__ constexpr size_t fwd_smem_bytes(int KB) {
  return 1024 /*alignment slack*/ + (size_t)(2 * KB + kFwdStages) * kBlkBytes + 256 /*barriers*/;
}

template <int KB, int MODE>
__global__ void __launch_bounds__(kFwdThreads, 1) k_sim_fwd(const __grid_constant__ FwdArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                                   // [2 halves][KB][128 rows][128 B]
  uint8_t* smB = smem + (size_t)2 * KB * kBlkBytes;      // [stages][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + (size_t)kFwdStages * kBlkBytes);
  uint64_t* b_full = bars;                    // [stages]
  uint64_t* b_empty = bars + kFwdStages;      // [stages]
  uint64_t* a_full = bars + 2 * kFwdStages;
  uint64_t* a_empty = a_full + 1;
  uint64_t* acc_full = a_full + 2;            // [2]
  uint64_t* acc_empty = a_full + 4;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kFwdStages; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    ptx::mbar_init(a_full, 1); ptx::mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 8); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      Walker wk(args.work);
      Segment sg;
      int stage = 0; uint32_t phase = 0, a_phase = 0;
      while (wk.next(sg)) {
        const FwdTerm& t = args.t[sg.owner];
        ptx::mbar_wait(a_empty, a_phase ^ 1, 101);
        ptx::mbar_expect_tx(a_full, 2 * KB * kBlkBytes);
        for (int h = 0; h < 2; ++h)
          for (int kb = 0; kb < KB; ++kb)
            ptx::tma_load_2d(smA + (size_t)(h * KB + kb) * kBlkBytes, &args.maps[t.a_map], a_full, kb * kKBlk,
                             sg.rb * kFwdRows + h * 128);
        a_phase ^= 1;
        for (int ct = sg.c_begin; ct < sg.c_end; ++ct)
          for (int kb = 0; kb < KB; ++kb) {
            ptx::mbar_wait(&b_empty[stage], phase ^ 1, 102);
            ptx::mbar_expect_tx(&b_full[stage], kBlkBytes);
            ptx::tma_load_2d(smB + (size_t)stage * kBlkBytes, &args.maps[t.k_map], &b_full[stage], kb * kKBlk,
                             ct * kTileN);
            if (++stage == kFwdStages) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, kTileN, 0, 0);
      Walker wk(args.work);
      Segment sg;
      int stage = 0; uint32_t phase = 0, a_phase = 0, it = 0;
      const uint32_t a_addr = ptx::smem_u32(smA), b_addr = ptx::smem_u32(smB);
      while (wk.next(sg)) {
        ptx::mbar_wait(a_full, a_phase, 111); a_phase ^= 1;
        ptx::tc_fence_after();
        for (int ct = sg.c_begin; ct < sg.c_end; ++ct, ++it) {
          const uint32_t buf = it & 1;
          ptx::mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1, 112);
          ptx::tc_fence_after();
          for (int kb = 0; kb < KB; ++kb) {
            ptx::mbar_wait(&b_full[stage], phase, 113);
            ptx::tc_fence_after();
            // alternate the two row halves: consecutive MMAs then accumulate into different TMEM tiles,
            // which hides the accumulate-to-accumulate dependency (measured: 114 -> 80 cycles per
            // 128x128x16 MMA, tools/umma_probe.cu)
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint64_t ad = ptx::umma_desc_sw128(a_addr + (h * KB + kb) * kBlkBytes + k * 32, 16, 1024);
                const uint64_t bd = ptx::umma_desc_sw128(b_addr + stage * kBlkBytes + k * 32, 16, 1024);
                ptx::umma_ss(tmem_base + (buf * 2 + h) * 128, ad, bd, idesc, (kb | k) != 0);
              }
            ptx::umma_commit(&b_empty[stage]);
            if (++stage == kFwdStages) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit(&acc_full[buf]);
        }
        ptx::umma_commit(a_empty);
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: one thread per anchor row =================
    const int half = (warp - 4) >> 2, quad = warp & 3;
    Walker wk(args.work);
    Segment sg;
    uint32_t it = 0;
    while (wk.next(sg)) {
      const FwdTerm& t = args.t[sg.owner];
      const int row = sg.rb * kFwdRows + half * 128 + quad * 32 + lane;
      const bool valid = row < t.N1;
      int p0 = 0, p1 = 0;
      if (valid) { const int y = t.a_cls[row]; p0 = t.k_seg[y]; p1 = t.k_seg[y + 1]; }
      const unsigned plen = (unsigned)(p1 - p0);
      int wmin = valid ? p0 : 0x7fffffff, wmax = valid ? p1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        wmin = min(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
        wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
      }
      const float scale = t.scale_log2;
      const float negi = (MODE == 1 && valid) ? t.neg[row] : 1.f;
      const int self_col = t.self_mask ? row : -1;
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;   // MODE 0: 4 partial sums; MODE 1: acc0 = pos (log2 units), acc1 = S
      for (int ct = sg.c_begin; ct < sg.c_end; ++ct, ++it) {
        const uint32_t buf = it & 1;
        ptx::mbar_wait(&acc_full[buf], (it >> 1) & 1, 121);
        ptx::tc_fence_after();
        const int cb = ct * kTileN;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (buf * 2 + half) * 128;
        const bool touches = !(cb + kTileN <= wmin || cb >= wmax);
        if (MODE == 0) {
          const bool fast = !touches && (cb + kTileN <= t.N2);
          uint32_t va[32], vb[32];
          ptx::tmem_ld32(taddr, va);
          ptx::tmem_ld_wait(va);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t (&cur)[32] = (ch & 1) ? vb : va;
            uint32_t (&nxt)[32] = (ch & 1) ? va : vb;
            if (ch < 3) ptx::tmem_ld32(taddr + (ch + 1) * 32, nxt);
            if (fast) {
#pragma unroll
              for (int c = 0; c < 32; c += 4) {
                acc0 += ptx::ex2(__uint_as_float(cur[c]) * scale);
                acc1 += ptx::ex2(__uint_as_float(cur[c + 1]) * scale);
                acc2 += ptx::ex2(__uint_as_float(cur[c + 2]) * scale);
                acc3 += ptx::ex2(__uint_as_float(cur[c + 3]) * scale);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const int col = cb + ch * 32 + c;
                const float e = ptx::ex2(__uint_as_float(cur[c]) * scale);
                const bool isneg = ((unsigned)(col - p0) >= plen) && (col < t.N2);
                acc0 += isneg ? e : 0.f;
              }
            }
            if (ch < 3) ptx::tmem_ld_wait(nxt);
          }
        } else {
          if (touches) {
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
              const int c0 = cb + ch * 32;
              if (c0 + 32 <= wmin || c0 >= wmax) continue;      // warp-uniform
              uint32_t v[32];
              ptx::tmem_ld32(taddr + ch * 32, v);
              ptx::tmem_ld_wait(v);
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const int col = c0 + c;
                const bool ispos = ((unsigned)(col - p0) < plen) && (col != self_col);
                const float x = __uint_as_float(v[c]) * scale;      // logit in log2 units
                const float den = ptx::ex2(x) + negi;
                acc0 += ispos ? (x - ptx::lg2(den)) : 0.f;
                acc1 += ispos ? ptx::rcp(den) : 0.f;
              }
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
      }
      if (valid) {
        if (MODE == 0) atomicAdd(&t.neg[row], (acc0 + acc1) + (acc2 + acc3));
        else { atomicAdd(&t.pos[row], acc0 * kLn2); atomicAdd(&t.ssum[row], acc1); }
      }
    }
  }
  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------
// work table builder: one item per (term, 256-row block).  MODE 0: every column tile;
// MODE 1: the column tiles that overlap the class segments of the block's rows.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_build_work(const __grid_constant__ BuildArgs a) {
  __shared__ int carry;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) { carry = 0; a.prefix[0] = 0; }
  __syncthreads();
  for (int base = 0; base < a.nitems; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int len = 0;
    if (i < a.nitems) {
      int ti = 0;
      for (int q = 1; q < a.num_terms; ++q) if (i >= a.t[q].item_base) ti = q;
      const BuildTerm& t = a.t[ti];
      const int rb = i - t.item_base;
      int ct0 = 0, ct1 = (t.N2 + kTileN - 1) / kTileN;
      if (a.mode == 1) {
        const int r0 = rb * a.rows_per_item, r1 = min(t.N1, r0 + a.rows_per_item) - 1;
        const int p0 = t.k_seg[t.a_cls[r0]], p1 = t.k_seg[t.a_cls[r1] + 1];
        if (p1 > p0) { ct0 = p0 / kTileN; ct1 = (p1 + kTileN - 1) / kTileN; } else { ct0 = ct1 = 0; }
      }
      a.items[i] = WorkItem{ti, rb, ct0, ct1};
      len = ct1 - ct0;
    }
    int s = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, s, o); if ((threadIdx.x & 31) >= o) s += v; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = wsum[threadIdx.x], ws = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, ws, o); if ((int)threadIdx.x >= o) ws += v; }
      wsum[threadIdx.x] = ws - w;
    }
    __syncthreads();
    const int incl = carry + wsum[threadIdx.x >> 5] + s;
    if (i < a.nitems) a.prefix[i + 1] = incl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = incl;
    __syncthreads();
  }
}

// ---- tensor map factory (driver entry point resolved at run time: no link-time libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map(CUtensorMap* out, const void* base, int rows, int c_pad) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    MSCS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
    MSCS_CHECK_ARG(p != nullptr && qr == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = (EncodeTiledFn)p;
  }
  cuuint64_t dims[2] = {(cuuint64_t)c_pad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)c_pad * 2};
  cuuint32_t box[2] = {(cuuint32_t)kKBlk, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MSCS_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

int launch_build_work(const BuildArgs& b, cudaStream_t st) {
  k_build_work<<<1, 1024, 0, st>>>(b);
  MSCS_LAUNCH_CHECK();
  return 0;
}

int sm_count() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
  return n > 0 ? n : 148;
}

template <int KB>
static int launch_fwd(const FwdArgs& args, int mode, cudaStream_t st) {
  const size_t smem = fwd_smem_bytes(KB);
  if (int rc = ensure_trap_buffer()) return rc;
  if (mode == 0) {
    MSCS_CUDA(cudaFuncSetAttribute(k_sim_fwd<KB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sim_fwd<KB, 0><<<sm_count(), kFwdThreads, smem, st>>>(args);
  } else {
    MSCS_CUDA(cudaFuncSetAttribute(k_sim_fwd<KB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sim_fwd<KB, 1><<<sm_count(), kFwdThreads, smem, st>>>(args);
  }
  MSCS_LAUNCH_CHECK();
  return 0;
}

}  // namespace mscs

using namespace mscs;

extern "C" size_t mscs_sim_workspace_bytes(const mscs_sim_job* job) {
  if (!job) return 0;
  size_t items = 0;
  for (int t = 0; t < job->num_terms; ++t) {
    items += 2 * (size_t)ceil_div(job->terms[t].N1, kFwdRows);                 // forward: two sweeps
    items += (size_t)ceil_div(job->terms[t].N1, 128) + ceil_div(job->terms[t].N2, 128);   // backward passes
  }
  return 2 * 4096 + items * (sizeof(WorkItem) + sizeof(int)) + 16 * 64;
}

extern "C" int mscs_sim_forward(const mscs_sim_job* job, void* stream_) {
  int rc = validate_job(job);
  if (rc) return rc;
  MSCS_CHECK_ARG(job->work, "work buffer is null");
  cudaStream_t st = (cudaStream_t)stream_;
  FwdArgs args{};
  // one tensor map per distinct operand matrix
  const void* bases[MSCS_MAX_SCALES]; int nmaps = 0;
  auto map_of = [&](const void* base, int rows) -> int {
    for (int i = 0; i < nmaps; ++i) if (bases[i] == base) return i;
    if (nmaps == MSCS_MAX_SCALES) return -1;
    if (make_tensor_map(&args.maps[nmaps], base, (rows + 255) / 256 * 256, job->C_pad)) return -2;
    bases[nmaps] = base;
    return nmaps++;
  };
  BuildArgs b{};
  int nitems = 0;
  for (int t = 0; t < job->num_terms; ++t) {
    const mscs_term& m = job->terms[t];
    const int am = map_of(m.a_bf16, m.N1), km = map_of(m.k_bf16, m.N2);
    if (am == -2 || km == -2) return -1;
    MSCS_CHECK_ARG(am >= 0 && km >= 0, "too many distinct operand matrices");
    args.t[t] = FwdTerm{m.a_cls, m.k_seg, m.neg_sum, m.pos_sum, m.s_sum, m.N1, m.N2, m.self_mask, am, km,
                        kLog2e / m.temperature};
    b.t[t] = BuildTerm{m.a_cls, m.k_seg, m.N1, m.N2, nitems};
    nitems += ceil_div(m.N1, kFwdRows);
  }
  b.num_terms = job->num_terms; b.nitems = nitems; b.rows_per_item = kFwdRows;
  char* w = (char*)job->work + 4096;      // [0,4096): finalise accumulators
  for (int mode = 0; mode < 2; ++mode) {
    b.mode = mode;
    b.items = (WorkItem*)w; w += align_up(sizeof(WorkItem) * (size_t)nitems, 64);
    b.prefix = (int*)w;     w += align_up(sizeof(int) * (size_t)(nitems + 1), 64);
    rc = launch_build_work(b, st);
    if (rc) return rc;
    args.work = WorkTable{b.items, b.prefix, nitems};
    switch (job->C_pad / 64) {
      case 1: rc = launch_fwd<1>(args, mode, st); break;
      case 2: rc = launch_fwd<2>(args, mode, st); break;
      case 3: rc = launch_fwd<3>(args, mode, st); break;
      default: rc = launch_fwd<4>(args, mode, st); break;
    }
    if (rc) return rc;
  }
  return launch_finalize(job, st);
}
