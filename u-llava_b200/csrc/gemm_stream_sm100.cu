// Weight-streaming GEMM for the decode step (M <= 32 tokens):  D[M,N] = epi(X[M,K] * W[N,K]^T)
//
// Call sites (one token per sample, KV-cached decode): LLaMA q/k/v, o, gate/up, down projections and lm_head
// (hf:models/llama/modeling_llama.py:171-291, models/ullava_core.py:325) as driven by generate()
// (models/ullava.py:350-362).  2*M FLOP per 2-byte weight: HBM-bound, the job is to stream W exactly once at
// full bandwidth.
//
// Design (B200, sm_100a):
//   * swap-AB: the 128-row tcgen05 operand is a tile of W (K-major, TMA SWIZZLE_128B), the padded batch is the
//     UMMA N (16 or 32; rows >= M of X are zero-filled by TMA), accumulators [128 weight rows x BN] in TMEM;
//   * stream-K: the (weight tile, k-block) units are cut into gridDim.x equal contiguous ranges, one per SM, so
//     every SM streams the same number of bytes whatever N and K are (no wave quantisation, no split-K heuristics);
//   * a tile cut across CTAs is reduced IN the kernel: each contributor parks its fp32 partial in the L2-resident
//     workspace and bumps the tile's counter; the last one to arrive adds the others' partials to its registers and
//     runs the fused epilogue (bias / activation / SiLU*mul / residual, transposed store).  Nobody waits for anybody,
//     and there is no second kernel;
//   * programmatic dependent launch: weights do not depend on the previous kernel, so the producer warp fills the
//     whole shared-memory ring with W tiles BEFORE griddepcontrol.wait and only then loads X; launched with the
//     programmatic-serialization attribute this overlaps pipeline fill with the previous kernel's tail (the
//     norm / rope / attention kernels of the decode step call griddepcontrol.launch_dependents at their start).
//   * next-weight L2 prefetch: once its own loads are issued, each CTA pulls the tiles its successor (the next GEMM of
//     the decode chain, hinted by ullava_gemm_next_weight) will ask for right after its ring fill into L2
//     (cp.async.bulk.prefetch.tensor), so HBM stays busy through this kernel's reduction tail and the launch boundary;
//   * 8-stage TMA ring (128 KB of W in flight per SM, one CTA per SM so that the next GEMM's CTAs map 1:1 onto SMs),
//     warp-specialised like the large-M kernel: warp 0 producer, warp 1 MMA issuer, warp 2 TMEM allocator,
//     warps 4-7 epilogue; accumulator double buffered in TMEM; partial tiles are stored row-interleaved
//     ([BN/4][128] float4) so that a warp's store / load is one contiguous 512-byte run.
#include "gemm_stream.cuh"

namespace ullava {

// EK: 0 = bias / residual only, 1 = SiLU(gate) * up, 2 = bias + ReLU / GELU / quick-GELU + residual
template <typename T, int BN, int EK>
__global__ void __launch_bounds__(GS_THREADS, GS_CTAS_PER_SM)
gemm_stream_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                   const __grid_constant__ CUtensorMap tmP, const StreamParams p) {
  using S = StreamSmem<BN>;
  constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr uint32_t kIdesc = make_idesc_f16(T16<T>::kUmmaFormat, GS_BM, BN);

  extern __shared__ uint8_t gs_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gs_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + GS_STAGES;
  uint64_t* tmem_full = empty_bar + GS_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  volatile int* last_flag = reinterpret_cast<volatile int*>(tmem_ptr + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int U = p.units;
  const int lo = gs_lo(blockIdx.x, U, G), hi = gs_lo(blockIdx.x + 1, U, G);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GS_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4 * 32);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();  // the next kernel of the stream may start its own prologue / weight prefetch

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int pre_end = (hi - lo) < GS_STAGES ? hi : lo + GS_STAGES;
      // ring fill with weights only: independent of whatever kernel precedes this one
      for (int u = lo; u < pre_end; ++u) {
        const int t = u / p.kb_total, kb = u - t * p.kb_total;
        const int s = u - lo;
        mbar_expect_tx(&full_bar[s], S::kStageBytes);
        tma_load_2d_hint(smem + s * S::kStageBytes, &tmW, &full_bar[s], kb * GS_BK, t * GS_BM, kEvictFirst);
      }
      pdl_wait();  // activations (and everything else) of the previous kernel are now visible
      for (int u = lo; u < pre_end; ++u) {
        const int t = u / p.kb_total, kb = u - t * p.kb_total;
        const int s = u - lo;
        tma_load_2d_hint(smem + s * S::kStageBytes + S::kWBytes, &tmX, &full_bar[s], kb * GS_BK, 0, kEvictLast);
      }
      stage = (pre_end - lo) % GS_STAGES;
      phase = (pre_end - lo) == GS_STAGES ? 1u : 0u;
      for (int u = pre_end; u < hi; ++u) {
        const int t = u / p.kb_total, kb = u - t * p.kb_total;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sw = smem + stage * S::kStageBytes;
        mbar_expect_tx(&full_bar[stage], S::kStageBytes);
        tma_load_2d_hint(sw, &tmW, &full_bar[stage], kb * GS_BK, t * GS_BM, kEvictFirst);
        tma_load_2d_hint(sw + S::kWBytes, &tmX, &full_bar[stage], kb * GS_BK, 0, kEvictLast);
        if (++stage == GS_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      // All of this CTA's weights are in flight.  Keep HBM busy through the reduction / epilogue tail and the kernel
      // boundary: pull the tiles the next GEMM's CTA on this SM will ask for right after its ring fill into L2.
      if (p.pf_count > 0 && static_cast<int>(blockIdx.x) < p.pf_grid) {
        const int plo = gs_lo(blockIdx.x, p.pf_units, p.pf_grid) + p.pf_skip;
        const int phi = min(gs_lo(blockIdx.x + 1, p.pf_units, p.pf_grid), plo + p.pf_count);
        for (int u = plo; u < phi; ++u) {
          const int t = u / p.pf_kb_total, kb = u - t * p.pf_kb_total;
          tma_prefetch_2d(&tmP, kb * GS_BK, t * GS_BM);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int u = lo;
      while (u < hi) {
        const int t = u / p.kb_total;
        const int seg_end = min(hi, (t + 1) * p.kb_total);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int v = u; v < seg_end; ++v) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sw = smem_u32(smem + stage * S::kStageBytes);
          const uint64_t a_desc = make_kmajor_sw128_desc(sw);
          const uint64_t b_desc = make_kmajor_sw128_desc(sw + S::kWBytes);
#pragma unroll
          for (int k = 0; k < GS_BK / 16; ++k)
            umma_f16<1>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (v > u || k > 0) ? 1u : 0u);
          umma_commit<1>(&empty_bar[stage]);
          if (++stage == GS_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit<1>(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        u = seg_end;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread = one weight row (output column), BN batch entries =====================
    pdl_wait();  // no global write (output, partials, counters) before the previous kernel has fully finished
    const int q = warp & 3;
    const int etid = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    GsOut out;
    out.D = p.D; out.ldd = p.ldd; out.bias = p.bias; out.residual = p.residual; out.ldr = p.ldr;
    out.M = p.M; out.N = p.N; out.epilogue = p.epilogue; out.out_f32 = p.out_f32; out.kb_total = p.kb_total;
    out.partials = p.partials; out.counters = p.counters;
    int u = lo;
    while (u < hi) {
      const int t = u / p.kb_total;
      const int t0 = t * p.kb_total;
      const int seg_end = min(hi, t0 + p.kb_total);
      const bool whole = (u == t0) && (seg_end == t0 + p.kb_total);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      float v[BN];
      {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
        if constexpr (BN == 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        } else {
          uint32_t r[16];
          tmem_ld_32x16(taddr, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);  // accumulator is in registers: the MMA warp may reuse it
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }

      gs_finish_segment<T, BN, EK>(out, U, G, lo, t, whole, v, last_flag, q, lane, etid);
      u = seg_end;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}

// -----------------------------------------------------------------------------------------------------------------
// Host side
// -----------------------------------------------------------------------------------------------------------------
template <typename T, int BN, int EK>
static int stream_launch(const CUtensorMap& tw, const CUtensorMap& tx, const CUtensorMap& tp, const StreamParams& p,
                         int grid, bool pdl, cudaStream_t stream) {
  using S = StreamSmem<BN>;
  auto kern = gemm_stream_kernel<T, BN, EK>;
  static SmemOptIn opt_in;   // per device (common.cuh)
  { const int _st = ensure_dynamic_smem(kern, S::kTotal, opt_in); if (_st != OK) return _st; }
  return check_cuda(launch_pdl(kern, dim3(grid), dim3(GS_THREADS), S::kTotal, stream, pdl, tw, tx, tp, p),
                    "gemm_stream_kernel launch");
}

template <typename T>
static int stream_dispatch(int bn, int ek, const CUtensorMap& tw, const CUtensorMap& tx, const CUtensorMap& tp,
                           const StreamParams& p, int grid, bool pdl, cudaStream_t stream) {
  if (bn == 16) {
    if (ek == 0) return stream_launch<T, 16, 0>(tw, tx, tp, p, grid, pdl, stream);
    if (ek == 1) return stream_launch<T, 16, 1>(tw, tx, tp, p, grid, pdl, stream);
    return stream_launch<T, 16, 2>(tw, tx, tp, p, grid, pdl, stream);
  }
  if (ek == 0) return stream_launch<T, 32, 0>(tw, tx, tp, p, grid, pdl, stream);
  if (ek == 1) return stream_launch<T, 32, 1>(tw, tx, tp, p, grid, pdl, stream);
  return stream_launch<T, 32, 2>(tw, tx, tp, p, grid, pdl, stream);
}

size_t gemm_stream_workspace_bytes(int sm_count) {
  return kStreamCounterBytes + static_cast<size_t>(2) * GS_CTAS_PER_SM * sm_count * GS_BM * 32 * sizeof(float);
}

// a: the caller's GEMM (M <= 32 rows of X, N weight rows).  Uses ctx->workspace: [0, 64 KB) tile counters (kept at
// zero between launches), then the partial tiles.
int gemm_stream_run(Context* ctx, const GemmArgs& a, cudaStream_t stream) {
  const int bn = a.M <= 16 ? 16 : 32;
  ULLAVA_REQUIRE(a.M <= 32, "gemm_stream: M = %d > 32", a.M);
  if (a.epilogue == EPI_SILU_MUL) ULLAVA_REQUIRE((a.N % 32) == 0, "gemm_stream: SILU_MUL needs N %% 32 == 0");
  StreamParams p{};
  p.D = a.D; p.ldd = a.ldd; p.bias = a.bias; p.residual = a.residual; p.ldr = a.ldr;
  p.M = a.M; p.N = a.N; p.K = a.K; p.epilogue = a.epilogue; p.out_f32 = a.out_f32;
  p.num_t = (a.N + GS_BM - 1) / GS_BM;
  p.kb_total = (a.K + GS_BK - 1) / GS_BK;
  const long long units = static_cast<long long>(p.num_t) * p.kb_total;
  ULLAVA_REQUIRE(units * GS_CTAS_PER_SM * ctx->sm_count < (1ll << 31), "gemm_stream: %lld units exceed the 32-bit partition arithmetic", units);
  p.units = static_cast<int>(units);
  const int slots = GS_CTAS_PER_SM * ctx->sm_count;
  const int grid = p.units < slots ? p.units : slots;
  const size_t need = kStreamCounterBytes + static_cast<size_t>(2) * grid * GS_BM * bn * sizeof(float);
  if (need > ctx->workspace_bytes || static_cast<size_t>(p.num_t) * sizeof(int) > kStreamCounterBytes) {
    set_last_error("gemm_stream: workspace too small (%zu > %zu) or too many tiles (%d)", need, ctx->workspace_bytes,
                   p.num_t);
    return ERR_WORKSPACE;
  }
  p.counters = reinterpret_cast<int*>(ctx->workspace);
  p.partials = reinterpret_cast<float*>(static_cast<uint8_t*>(ctx->workspace) + kStreamCounterBytes);
  CUtensorMap tw, tx;
  int st = encode_tmap_2d(&tw, a.B, 2, a.K, a.N, a.ldb * 2, GS_BK, GS_BM, true);
  if (st) return st;
  st = encode_tmap_2d(&tx, a.A, 2, a.K, a.M, a.lda * 2, GS_BK, bn, true);
  if (st) return st;
  // next-weight hint -> prefetch-only tensor map over the next GEMM's W (same 64 x 128 box)
  CUtensorMap tp = tw;
  if (ctx->next_w != nullptr && ctx->prefetch_units > 0 && ctx->next_n > 0 && ctx->next_k > 0) {
    const int n_t = (ctx->next_n + GS_BM - 1) / GS_BM, n_kb = (ctx->next_k + GS_BK - 1) / GS_BK;
    const long long n_units = static_cast<long long>(n_t) * n_kb;
    if (n_units * slots < (1ll << 31) &&
        encode_tmap_2d(&tp, ctx->next_w, 2, ctx->next_k, ctx->next_n, ctx->next_ldb * 2, GS_BK, GS_BM, true) == OK) {
      p.pf_kb_total = n_kb;
      p.pf_units = static_cast<int>(n_units);
      p.pf_grid = p.pf_units < slots ? p.pf_units : slots;
      p.pf_skip = GS_STAGES;  // the ring fill is issued by the next kernel itself, before its griddepcontrol.wait
      p.pf_count = ctx->prefetch_units;
    } else {
      tp = tw;
    }
  }
  ctx->next_w = nullptr;
  const bool pdl = ctx->pdl != 0;
  const int ek = a.epilogue == EPI_SILU_MUL ? 1 : (a.epilogue == EPI_NONE ? 0 : 2);
  st = a.dtype == DT_BF16 ? stream_dispatch<__nv_bfloat16>(bn, ek, tw, tx, tp, p, grid, pdl, stream)
                          : stream_dispatch<__half>(bn, ek, tw, tx, tp, p, grid, pdl, stream);
  if (st == OK) ctx->launches += 1;
  return st;
}

}  // namespace ullava
