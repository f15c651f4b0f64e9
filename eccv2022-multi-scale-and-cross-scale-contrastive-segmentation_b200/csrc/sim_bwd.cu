// sim_bwd.cu -- K4 (tensor half): fused backward of the similarity / loss, recomputing logit
// tiles instead of storing them.  Replaces the autograd backward of V2.py:138-188 and
// _ms.py:95-156 up to d loss / d (unit rows); the normalisation backward + dense scatter is
// gather.cu.
//
// One pass computes  dX_rows += scale * W(rows, cols) * Y_cols  with
//   W_ij = E_ij (cS_i + cS_j)                                   (negative pair)
//   W_ij = -( cPN_i / (E_ij + neg_i) + cPN_j / (E_ij + neg_j) ) (positive pair, 0 on the diagonal)
// where cS = S/(div N1), cPN = neg/(div N1) come from the forward (k_finalize).  A single-scale
// term is ONE pass with both the row and column coefficients (W = G + G^T, so no second product);
// a cross-scale term is a pass with row coefficients only (dA) and, unless the key side is
// detached, a pass with rows = keys and column coefficients only (dK).
//
// CTA = 128 rows (X tile resident in smem) x a run of 128-column tiles.  Per tile:
//   MMA 1  S = X Y^T              (both K-major)        -> TMEM, double buffered
//   epilogue: W = f(exp2(S)) as packed bf16 written back INTO the S columns of TMEM (in place)
//   MMA 2  dX += W Y              (A operand = W from TMEM; Y tile re-used from smem as an
//                                  MN-major B operand) -> TMEM
// dX stays in TMEM for the whole run and is flushed with fp32 reductions at the end.
#include "sim_tc.cuh"
#include <stdlib.h>

namespace mscs {

constexpr int kBwdEpiWarps = 16;     // 4 per SM sub-partition: thread = (row, 32-column quarter of the tile).  The
                                     // S -> W conversion sits between the two MMAs of a tile (S buffer cycle =
                                     // S MMAs + conversion + dX MMAs), so its LATENCY bounds the tile rate
constexpr int kBwdThreads = 128 + 32 * kBwdEpiWarps;

struct BwdDev {
  const int* row_cls; const int* col_seg;
  const float* row_cs; const float* row_cpn; const float* row_neg;
  const float* col_cs; const float* col_cpn; const float* col_neg;
  float* dF; int ld;
  int n_rows, n_cols, self_mask, x_map, y_map;
  float scale_log2, out_scale;
};
struct BwdArgs {
  alignas(64) CUtensorMap maps[MSCS_MAX_SCALES];
  BwdDev p[MSCS_MAX_PASSES];
  WorkTable work;
  const float* grad_out;
  int flags;      // MSCS_DEBUG_FLAGS experiments: 64 = un-split dX MMAs (N = C_pad, one release per Y stage)
};

__host__ __device__ constexpr size_t bwd_smem_bytes(int KB) {
  return 1024 + (size_t)(3 * KB) * kBlkBytes + (size_t)kBwdEpiWarps * 2 * 3 * 32 * sizeof(float) + 256;   // 24 barriers + TMEM slot < 256 B
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int KB>
__global__ void __launch_bounds__(kBwdThreads, 1) k_sim_bwd(const __grid_constant__ BwdArgs args) {
  constexpr int CP = KB * 64;                       // padded channel count = N of the second MMA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                                         // [KB][128][128 B]   X tile
  uint8_t* smB = smA + (size_t)KB * kBlkBytes;                 // [2][KB][128][128 B] Y tiles
  float* cstat = reinterpret_cast<float*>(smB + (size_t)2 * KB * kBlkBytes);   // per epilogue warp: [2 buffers][3][32] column coefficients
  uint64_t* bars = reinterpret_cast<uint64_t*>(cstat + kBwdEpiWarps * 2 * 3 * 32);
  uint64_t* a_full = bars;        uint64_t* a_empty = bars + 1;
  uint64_t* b_empty = bars + 20;                                    // [2 stages][2 channel halves]: the dX product
                                                                    // runs channel half by channel half, so the first
                                                                    // K-blocks of a Y stage are refilled while the
                                                                    // second half of the product still reads the rest
  uint64_t* s_full = bars + 4;                                      // [2]
  uint64_t* w_full = bars + 6;                                      // [2 S buffers]: one phase per two tiles, so a
                                                                    // warp that runs a tile ahead of the MMA thread
                                                                    // cannot overrun it
  uint64_t* df_full = bars + 10;  uint64_t* df_empty = bars + 11;
  uint64_t* b_full = bars + 12;                                     // [2 stages][4 K-blocks]: the S MMAs start on
                                                                    // the first 16 KB of a tile, not the whole 64 KB
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  constexpr bool kCanSplit = (KB % 2 == 0);
  const bool split = kCanSplit && !(args.flags & 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::mbar_init(a_full, 1); ptx::mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&b_empty[2 * i], 1); ptx::mbar_init(&b_empty[2 * i + 1], 1);
      for (int kb = 0; kb < 4; ++kb) ptx::mbar_init(&b_full[i * 4 + kb], 1);
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&w_full[i], kBwdEpiWarps);
    }
    ptx::mbar_init(df_full, 1); ptx::mbar_init(df_empty, kBwdEpiWarps);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#ifdef MSCS_WAIT_PROFILE     // effective SM clock of this launch: slot 31 accumulates (ns, cycles) of CTA 0
  const unsigned long long prof_t0 = ptx::globaltimer_ns();
  const long long prof_c0 = clock64();
#endif
  const uint32_t tmem_dF = tmem_base + 256;

  // Roles 0 and 1 run on the whole warp with warp-uniform control flow; one lane issues (see sim_fwd.cu)
  if (warp == 0) {
    // ================= TMA producer =================
    Walker wk(args.work);
    Segment sg;
    uint32_t a_phase = 0, it = 0;
    while (wk.next(sg)) {
      const BwdDev& p = args.p[sg.owner];
      ptx::mbar_wait(a_empty, a_phase ^ 1, 201);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(a_full, KB * kBlkBytes);
        for (int kb = 0; kb < KB; ++kb)
          ptx::tma_load_2d(smA + (size_t)kb * kBlkBytes, &args.maps[p.x_map], a_full, kb * kKBlk, sg.rb * 128);
      }
      a_phase ^= 1;
      for (int ct = sg.c_begin; ct < sg.c_end; ++ct, ++it) {
        const uint32_t st = it & 1;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          // un-split: both halves are released together by the same commit
          ptx::mbar_wait(&b_empty[2 * st + q], ((it >> 1) & 1) ^ 1, 202 + q);
          const int kb0 = q == 0 ? 0 : (KB + 1) / 2, kb1 = q == 0 ? (KB + 1) / 2 : KB;
          if (ptx::elect_one()) {
            for (int kb = kb0; kb < kb1; ++kb) {
              ptx::mbar_expect_tx(&b_full[st * 4 + kb], kBlkBytes);
              ptx::tma_load_2d(smB + (size_t)(st * KB + kb) * kBlkBytes, &args.maps[p.y_map], &b_full[st * 4 + kb],
                               kb * kKBlk, ct * kTileN);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, kTileN, 0, 0);   // S  = X Y^T
    constexpr uint32_t idesc_d = ptx::umma_idesc_bf16(128, CP, 0, 1);       // dX += W Y (B MN-major)
    constexpr uint32_t idesc_dh = ptx::umma_idesc_bf16(128, kCanSplit ? CP / 2 : CP, 0, 1);   // one channel half
    const uint32_t a_addr = ptx::smem_u32(smA), b_addr = ptx::smem_u32(smB);
    Walker wk(args.work);
    Segment sg;
    uint32_t a_phase = 0, it = 0, seg = 0;
    // S MMAs of tile `cur` for K-blocks [kb0, kb1)
    auto issue_s = [&](uint32_t cur, int kb0, int kb1) {
      const uint32_t st = cur & 1, ph = (cur >> 1) & 1;
      // S buffer `st` also holds W of tile cur-2: its consumer (the dX MMAs of tile cur-2) was issued
      // before this point and tcgen05.mma executes in issue order, so no extra barrier is needed
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&b_full[st * 4 + kb], ph, 211);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = ptx::umma_desc_sw128(a_addr + kb * kBlkBytes + k * 32, 16, 1024);
            const uint64_t bd = ptx::umma_desc_sw128(b_addr + (st * KB + kb) * kBlkBytes + k * 32, 16, 1024);
            ptx::umma_ss(tmem_base + st * 128, ad, bd, idesc_s, (kb | k) != 0);
          }
          if (kb == KB - 1) ptx::umma_commit(&s_full[st]);
        }
        __syncwarp();
      }
    };
    while (wk.next(sg)) {
      ptx::mbar_wait(a_full, a_phase, 212); a_phase ^= 1;
      ptx::mbar_wait(df_empty, (seg & 1) ^ 1, 213);
      ptx::tc_fence_after();
      const int ntiles = sg.c_end - sg.c_begin;
      issue_s(it, 0, KB);
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t cur = it + j, st = cur & 1;
        // Issue order S(cur+1) | dX(cur): the whole S chain of the next tile runs while the epilogue
        // turns S(cur) into W(cur).  (Splitting S(cur+1) around dX(cur) to release the Y stage earlier
        // measured slower: 0.585 vs 0.513 ms at cfg-2 -- the dX MMAs then wait for W.)
        if (j + 1 < ntiles) issue_s(cur + 1, 0, KB);
        // W of tile column quarter cq lives in columns [32cq, 32cq+16) of S buffer st: the K = 16 slice ks
        // (tile columns 16ks .. 16ks+15, two bf16 per TMEM column) is at column 32 (ks / 2) + 8 (ks % 2).
        // Y rows 16ks .. 16ks+15 are the matching K slice of B; LBO = next 64-channel block, SBO = next 8 rows.
        ptx::mbar_wait(&w_full[st], (cur >> 1) & 1, 214);
        ptx::tc_fence_after();
        if (!split) {
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t a_tm = tmem_base + st * 128 + (ks >> 1) * 32 + (ks & 1) * 8;
              const uint64_t bd = ptx::umma_desc_sw128(b_addr + st * KB * kBlkBytes + ks * 16 * 128, kBlkBytes, 1024);
              ptx::umma_ts(tmem_dF, a_tm, bd, idesc_d, (j | ks) != 0);
            }
            ptx::umma_commit(&b_empty[2 * st]); ptx::umma_commit(&b_empty[2 * st + 1]);
          }
          __syncwarp();
        } else {
          // channel half q of dX needs only the K-blocks of that half: release them as soon as it is done
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (ptx::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint32_t a_tm = tmem_base + st * 128 + (ks >> 1) * 32 + (ks & 1) * 8;
                const uint64_t bd = ptx::umma_desc_sw128(
                    b_addr + (st * KB + q * (KB / 2)) * kBlkBytes + ks * 16 * 128, kBlkBytes, 1024);
                ptx::umma_ts(tmem_dF + q * (CP / 2), a_tm, bd, idesc_dh, (j | ks) != 0);
              }
              ptx::umma_commit(&b_empty[2 * st + q]);
            }
            __syncwarp();
          }
        }
      }
      if (ptx::elect_one()) { ptx::umma_commit(df_full); ptx::umma_commit(a_empty); }
      __syncwarp();
      it += ntiles; ++seg;
    }
  } else if (warp >= 4) {
    // ================= epilogue: thread = (row, 32-column quarter) =================
    const int cq = (warp - 4) >> 2, quad = warp & 3;
    const int r_loc = quad * 32 + lane;                     // row inside the tile = TMEM lane
    float* my_cs = cstat + (warp - 4) * (2 * 3 * 32);       // this warp's [2 buffers][cs, cpn, neg][32 columns]
    const float gout = *args.grad_out;
    Walker wk(args.work);
    Segment sg;
    uint32_t it = 0, seg = 0;
    while (wk.next(sg)) {
      const BwdDev& p = args.p[sg.owner];
      const int row = sg.rb * 128 + r_loc;
      const bool valid = row < p.n_rows;
      int p0 = 0, p1 = 0;
      if (valid) { const int y = p.row_cls[row]; p0 = p.col_seg[y]; p1 = p.col_seg[y + 1]; }
      const unsigned plen = (unsigned)(p1 - p0);
      int wmin = valid ? p0 : 0x7fffffff, wmax = valid ? p1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        wmin = min(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
        wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
      }
      const float rcs = (valid && p.row_cs) ? p.row_cs[row] : 0.f;
      const float rcpn = (valid && p.row_cpn) ? p.row_cpn[row] : 0.f;
      const float rneg = (valid && p.row_neg) ? p.row_neg[row] : 1.f;
      const int self_col = p.self_mask ? row : -1;
      const float scale = p.scale_log2;
      const uint64_t scale2 = ptx::pack2(scale, scale), rcs2 = ptx::pack2(rcs, rcs);
      // column coefficients: lane l fetches those of column l of this warp's quarter one tile ahead (parked in
      // registers), then the warp stages them in its own smem slot -- no cross-warp synchronisation
      float pf_s = 0.f, pf_pn = 0.f, pf_neg = 1.f;
      auto prefetch_cols = [&](int ct_) {
        pf_s = 0.f; pf_pn = 0.f; pf_neg = 1.f;
        if (ct_ < sg.c_end) {
          const int c = ct_ * kTileN + cq * 32 + lane;
          if (c < p.n_cols) {
            if (p.col_cs) pf_s = p.col_cs[c];
            if (p.col_cpn) pf_pn = p.col_cpn[c];
            if (p.col_neg) pf_neg = p.col_neg[c];
          }
        }
      };
      auto publish_cols = [&](uint32_t b_) {
        float* cs = my_cs + b_ * 96;
        cs[lane] = pf_s; cs[32 + lane] = pf_pn; cs[64 + lane] = pf_neg;
        __syncwarp();
      };
      prefetch_cols(sg.c_begin);
      publish_cols(it & 1);
      for (int ct = sg.c_begin; ct < sg.c_end; ++ct, ++it) {
        const uint32_t buf = it & 1;
        const int cb = ct * kTileN + cq * 32;               // first global column of this thread's quarter
        const float* cs_s = my_cs + buf * 96; const float* cs_pn = cs_s + 32; const float* cs_neg = cs_s + 64;
        prefetch_cols(ct + 1);
        ptx::mbar_wait(&s_full[buf], (it >> 1) & 1, 221);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 128 + cq * 32;
        const bool touches = !(cb + 32 <= wmin || cb >= wmax);
        uint32_t v[32];
        ptx::tmem_ld32(taddr, v);
        ptx::tmem_ld_wait(v);
        uint32_t packed[16];
        if (!touches) {
          // W = exp2(S scale) (cS_row + cS_col), two columns per packed fp32x2 instruction
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 cc = *reinterpret_cast<const float4*>(cs_s + c);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float x0, x1;
              ptx::unpack2(ptx::mul2(ptx::pack2u(v[c + 2 * q], v[c + 2 * q + 1]), scale2), x0, x1);
              const uint64_t e2 = ptx::pack2(ptx::ex2(x0), ptx::ex2(x1));
              const uint64_t c2 = ptx::add2(rcs2, q == 0 ? ptx::pack2(cc.x, cc.y) : ptx::pack2(cc.z, cc.w));
              float w0, w1;
              ptx::unpack2(ptx::mul2(e2, c2), w0, w1);
              packed[(c >> 1) + q] = pack_bf16(w0, w1);
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float w[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int j = c + q;
              const int col = cb + j;
              const float e = ptx::ex2(__uint_as_float(v[j]) * scale);
              const bool ispos = (unsigned)(col - p0) < plen;
              const float wn = e * (rcs + cs_s[j]);
              const float wp = -(rcpn * ptx::rcp(e + rneg) + cs_pn[j] * ptx::rcp(e + cs_neg[j]));
              w[q] = ispos ? (col == self_col ? 0.f : wp) : wn;
            }
            packed[c >> 1] = pack_bf16(w[0], w[1]);
          }
        }
        // 32 logits -> 16 packed columns, written over S columns this thread has already read
        ptx::tmem_st16(taddr, packed);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&w_full[buf]);
        publish_cols(buf ^ 1);       // (the __syncwarp above: every lane is done reading slot buf^1's predecessor)
      }
      // ---- flush dX: this warp drains channel quarter cq of its 32 rows ----
      ptx::mbar_wait(df_full, seg & 1, 222);
      ptx::tc_fence_after();
      const float sc = p.out_scale * gout;
      float* drow = p.dF + (size_t)row * p.ld;
#pragma unroll 1
      for (int c0 = cq * (CP / 4); c0 < (cq + 1) * (CP / 4); c0 += 16) {
        uint32_t u[16];
        ptx::tmem_ld16(tmem_dF + ((uint32_t)(quad * 32) << 16) + c0, u);
        ptx::tmem_ld_wait16(u);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 16; c += 4)
            red_add_v4(drow + c0 + c, __uint_as_float(u[c]) * sc, __uint_as_float(u[c + 1]) * sc,
                       __uint_as_float(u[c + 2]) * sc, __uint_as_float(u[c + 3]) * sc);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(df_empty);
      ++seg;
    }
  }
  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
#ifdef MSCS_WAIT_PROFILE
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&ptx::g_wait_ns[31], ptx::globaltimer_ns() - prof_t0);
    atomicAdd(&ptx::g_wait_cnt[31], (unsigned long long)(clock64() - prof_c0));
  }
#endif
}

}  // namespace mscs

using namespace mscs;

template <int KB>
static int launch_bwd(const BwdArgs& args, cudaStream_t st) {
  const size_t smem = bwd_smem_bytes(KB);
  if (int rc = ensure_trap_buffer()) return rc;
  MSCS_CUDA(cudaFuncSetAttribute(k_sim_bwd<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_sim_bwd<KB><<<sm_count(), kBwdThreads, smem, st>>>(args);
  MSCS_LAUNCH_CHECK();
  return 0;
}

extern "C" int mscs_sim_backward(const mscs_sim_job* job, const float* grad_out, float* const* dF_sets,
                                 const int32_t* dF_ld, void* stream_) {
  int rc = validate_job(job);
  if (rc) return rc;
  MSCS_CHECK_ARG(job->work && grad_out && dF_sets && dF_ld, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream_;
  BwdPass passes[MSCS_MAX_PASSES];
  const int np = build_passes(job, passes);
  // the backward work tables live after the two forward tables in job->work
  size_t fwd_items = 0, ranges = 0;
  for (int t = 0; t < job->num_terms; ++t) {
    fwd_items += (size_t)ceil_div(job->terms[t].N2, 256);
    ranges += align_up(sizeof(int2) * (size_t)job->terms[t].N1, 64) +
              align_up(sizeof(int2) * (size_t)ceil_div(job->terms[t].N1, 128) * 4, 64);
  }
  char* w = (char*)job->work + 4096 + ranges +
            2 * (align_up(sizeof(WorkItem) * fwd_items, 64) + align_up(sizeof(int) * (fwd_items + 1), 64));
  BwdArgs args{};
  args.grad_out = grad_out;
  if (const char* e = getenv("MSCS_DEBUG_FLAGS")) args.flags = atoi(e);
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
  for (int i = 0; i < np; ++i) {
    const BwdPass& p = passes[i];
    const int xm = map_of(p.x_bf16, p.n_rows), ym = map_of(p.y_bf16, p.n_cols);
    if (xm == -2 || ym == -2) return -1;
    MSCS_CHECK_ARG(xm >= 0 && ym >= 0, "too many distinct operand matrices");
    MSCS_CHECK_ARG(dF_sets[p.row_set] && dF_ld[p.row_set] >= job->C_pad && dF_ld[p.row_set] % 4 == 0,
                   "pass %d: dF buffer of set %d missing or leading dimension < C_pad", i, p.row_set);
    args.p[i] = BwdDev{p.row_cls, p.col_seg, p.row_cs, p.row_cpn, p.row_neg, p.col_cs, p.col_cpn, p.col_neg,
                       dF_sets[p.row_set], dF_ld[p.row_set], p.n_rows, p.n_cols, p.self_mask, xm, ym,
                       p.scale_log2, p.out_scale};
    b.t[i] = BuildTerm{p.row_cls, p.col_seg, p.n_rows, p.n_cols, nitems, p.rb_lo, 0, 1 << 30};
    nitems += p.rb_hi - p.rb_lo;
  }
  b.num_terms = np; b.nitems = nitems; b.rows_per_item = 128; b.mode = 0;
  b.items = (WorkItem*)w; w += align_up(sizeof(WorkItem) * (size_t)nitems, 64);
  b.prefix = (int*)w;
  rc = launch_build_work(b, st);
  if (rc) return rc;
  args.work = WorkTable{b.items, b.prefix, nitems};
  switch (job->C_pad / 64) {
    case 1: return launch_bwd<1>(args, st);
    case 2: return launch_bwd<2>(args, st);
    case 3: return launch_bwd<3>(args, st);
    default: return launch_bwd<4>(args, st);
  }
}

// debug: read and reset the barrier wait profile of this translation unit (ns and count per tag % 32)
extern "C" int mscs_debug_wait_profile_bwd(unsigned long long* ns_out, unsigned long long* cnt_out) {
  MSCS_CUDA(cudaDeviceSynchronize());
  MSCS_CUDA(cudaMemcpyFromSymbol(ns_out, ptx::g_wait_ns, sizeof(unsigned long long) * 32));
  MSCS_CUDA(cudaMemcpyFromSymbol(cnt_out, ptx::g_wait_cnt, sizeof(unsigned long long) * 32));
  unsigned long long zero[32] = {};
  MSCS_CUDA(cudaMemcpyToSymbol(ptx::g_wait_ns, zero, sizeof(zero)));
  MSCS_CUDA(cudaMemcpyToSymbol(ptx::g_wait_cnt, zero, sizeof(zero)));
  return 0;
}
