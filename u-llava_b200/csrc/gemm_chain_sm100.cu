// Decode-layer chain: the weight-streaming GEMMs of one LLaMA decode layer (and the RMSNorms between them) in ONE
// persistent kernel.
//
// Call site: one generated token of generate() (models/ullava.py:350-362 -> LlamaDecoderLayer,
// hf:models/llama/modeling_llama.py:292-333).  Between two single-query attention kernels a decode layer runs
//     o_proj (+residual) -> RMSNorm -> gate/up (SiLU*mul) -> down (+residual) -> RMSNorm -> q/k/v of the next layer
// on M <= 32 rows: 404 MB of weights per LLaMA-7B layer, 63 us at the HBM roofline.  As six kernels
// (gemm_stream_kernel x 4, norm_kernel x 2) the chain took 98 us: every kernel boundary costs a pipeline drain, a
// split-reduction tail, a launch and a ring fill during which HBM idles.
//
// Here the chain is a PROGRAM (ChainStep array in device memory, built once per decode session) executed by one
// kernel of one CTA per SM:
//   * the TMA ring, the TMEM accumulators and the barriers live across the whole chain;
//   * warp 0 streams W tiles of step s, s+1, ... back to back -- weights do not depend on activations -- and, whenever
//     the ring is full, pulls the tiles further ahead into L2 (cp.async.bulk.prefetch.tensor), so HBM never waits for
//     the consumer side;
//   * warp 3 loads the X tiles of a step once that step's input is complete.  "Complete" is a grid-wide condition:
//     every CTA bumps a counter when its share of a step is stored (gs_finish_segment, same stream-K split reduction as
//     gemm_stream_kernel, hence bit-identical results), consumers poll it with acquire loads -- a grid barrier that
//     only the threads that need the data wait on;
//   * an RMSNorm in front of a step is done by the epilogue warps of the first M CTAs (one row each, the reduction
//     tree of norm_kernel, bit-identical) and published through a second counter;
//   * warp 1 issues the tcgen05 MMAs (swap-AB, UMMA 128 x BN x 16), warps 4-7 are the epilogue.
// Stalls of the consumer side (barrier round trips, the norm) do not idle HBM as long as ring + L2 look-ahead cover them.
#include "gemm_stream.cuh"

#include <cstring>

namespace ullava {

struct alignas(128) ChainStep {
  CUtensorMap tmW;   // [N, K] weights, box 64 x 128, SWIZZLE_128B
  CUtensorMap tmX;   // [M, K] activations, box 64 x BN, SWIZZLE_128B
  void* D;
  int64_t ldd;
  const void* residual;
  int64_t ldr;
  int N, K, num_t, kb_total, units, ek, out_f32, has_norm;
  // RMSNorm in front of the step (has_norm): norm_dst[r] = rmsnorm(norm_src[r]) * norm_w, then read through tmX
  const void* norm_src;
  const void* norm_w;
  void* norm_dst;
  int norm_cols;
  float norm_eps;
};
static_assert(sizeof(ChainStep) % 128 == 0, "ChainStep must keep the tensor maps 128-byte aligned in an array");

struct ChainParams {
  const ChainStep* steps;
  int n_steps;
  int M;             // valid batch rows
  float* partials;   // stream-K partial tiles (context workspace)
  int* counters;     // per-tile arrival counters (context workspace), zero between uses
  int* sync;         // [2 * n_steps]: sync[2s] = CTAs done with step s, sync[2s+1] = rows normalised for step s; zeroed
                     // by the host side before every launch
  int lookahead;     // W tiles pulled into L2 ahead of the ring
};

static constexpr int kChainMaxVec = 4;  // norm_kernel's kMaxVec

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void spin_until(const int* p, int target) {
  while (ld_acquire(p) < target) __nanosleep(32);
}
// generic-proxy writes of other CTAs (acquired above) -> this thread's TMA (async proxy) reads
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// One row of RMSNorm by the 128 epilogue threads, with the reduction tree of norm_kernel<T, true> launched with
// VT = min(512, ceil32(cols / 8)) threads (what rmsnorm_run launches for M <= 32 rows): thread (warp w, lane l) plays
// the virtual warps w, w + 4, ...; sums are combined in the same order, so the result is bit-identical.
template <typename T>
__device__ __forceinline__ void chain_rmsnorm_row(const T* __restrict__ xr, const T* __restrict__ w, T* __restrict__ yr,
                                                  int cols, float eps, int etid, float* red) {
  const int nvec = cols >> 3;
  int VT = ((nvec + 31) / 32) * 32;
  if (VT > 512) VT = 512;
  while (VT * kChainMaxVec < nvec) VT += 32;
  const int nvw = VT >> 5;                    // virtual warps
  const int w4 = etid >> 5, l = etid & 31;
  // pass 1: sum of squares per virtual thread, warp_sum per virtual warp
  for (int vw = w4; vw < nvw; vw += 4) {
    const int vt = vw * 32 + l;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kChainMaxVec; ++i) {
      const int vi = vt + i * VT;
      if (vi < nvec) {
        const uint4 r = *reinterpret_cast<const uint4*>(xr + vi * 8);
        const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack2<T>(u[j]);
          sq += f.x * f.x + f.y * f.y;
        }
      }
    }
    sq = warp_sum(sq);
    if (l == 0) red[vw] = sq;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  float t = (l < nvw) ? red[l] : 0.f;
  t = warp_sum(t);
  const float rstd = rsqrtf(t / cols + eps);
  // pass 2: normalise (HF: weight * (x_fp32 * rstd).to(dtype) -- one rounding before the weight multiply)
  for (int vi = etid; vi < nvec; vi += 128) {
    const uint4 r = *reinterpret_cast<const uint4*>(xr + vi * 8);
    const uint4 wv = *reinterpret_cast<const uint4*>(w + vi * 8);
    const uint32_t u[4] = {r.x, r.y, r.z, r.w}, wu[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack2<T>(u[j]);
      const float2 g = unpack2<T>(wu[j]);
      const float2 n = unpack2<T>(pack2<T>(f.x * rstd, f.y * rstd));
      o[j] = pack2<T>(g.x * n.x, g.y * n.y);
    }
    *reinterpret_cast<uint4*>(yr + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");  // red[] is reused by the next row / step
}

template <int BN>
struct ChainSmem {
  static constexpr int kWBytes = GS_BM * GS_BK * 2;
  static constexpr int kXBytes = BN * GS_BK * 2;
  static constexpr int kStageBytes = kWBytes + kXBytes;
  static constexpr int kBarOffset = GS_STAGES * kStageBytes;
  // fullW, fullX, empty per stage; tmem_full / tmem_empty x 2; tmem ptr + last_flag; norm reduction scratch
  static constexpr int kTotal = kBarOffset + (3 * GS_STAGES + 4) * 8 + 32 + 16 * 4 + 1024;
};

template <typename T, int BN>
__global__ void __launch_bounds__(GS_THREADS, 1)
gemm_chain_kernel(const ChainParams p) {
  using S = ChainSmem<BN>;
  constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr uint32_t kIdesc = make_idesc_f16(T16<T>::kUmmaFormat, GS_BM, BN);

  extern __shared__ uint8_t gc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gc_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_w = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* full_x = full_w + GS_STAGES;
  uint64_t* empty_bar = full_x + GS_STAGES;
  uint64_t* tmem_full = empty_bar + GS_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  volatile int* last_flag = reinterpret_cast<volatile int*>(tmem_ptr + 1);
  float* red = reinterpret_cast<float*>(tmem_ptr + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int cta = blockIdx.x;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GS_STAGES; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&full_x[s], 1);
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
  pdl_launch_dependents();  // the attention kernel behind this chain may become resident and wait

  if (warp == 0) {
    // ===================== W producer: never waits for anything but ring slots =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // L2 look-ahead: a cursor (ps, pu) over this CTA's units of the WHOLE chain, in load order, kept in front of the
      // unit being loaded; n_loaded / n_pf count units passed by the ring loads / by the cursor
      int ps = -1, pu = 0, p_hi = 0, n_loaded = 0, n_pf = 0;
      auto pf_advance = [&]() {   // moves the cursor to this CTA's next unit; false at the end of the chain
        ++pu;
        while (pu >= p_hi) {
          if (++ps >= p.n_steps) return false;
          pu = gs_lo(cta, p.steps[ps].units, G);
          p_hi = gs_lo(cta + 1, p.steps[ps].units, G);
        }
        return true;
      };
      pu = -1;
      bool pf_live = p.lookahead > 0 && pf_advance();   // cursor on the first unit
      for (int s = 0; s < p.n_steps; ++s) {
        const ChainStep& st = p.steps[s];
        const int lo = gs_lo(cta, st.units, G), hi = gs_lo(cta + 1, st.units, G);
        if (lo < hi) tma_prefetch_desc(&st.tmW);
        for (int u = lo; u < hi; ++u) {
          while (pf_live && n_pf <= n_loaded) {   // the ring load itself covers this unit: skip it
            pf_live = pf_advance();
            ++n_pf;
          }
          while (!mbar_try_wait(&empty_bar[stage], phase ^ 1)) {
            // ring full (the consumer side is waiting for a grid-wide condition or reducing): keep HBM busy
            if (pf_live && n_pf - n_loaded <= p.lookahead) {
              const ChainStep& sp = p.steps[ps];
              const int t = pu / sp.kb_total, kb = pu - t * sp.kb_total;
              tma_prefetch_2d(&sp.tmW, kb * GS_BK, t * GS_BM);
              pf_live = pf_advance();
              ++n_pf;
            }
          }
          const int t = u / st.kb_total, kb = u - t * st.kb_total;
          mbar_expect_tx(&full_w[stage], S::kWBytes);
          tma_load_2d_hint(smem + stage * S::kStageBytes, &st.tmW, &full_w[stage], kb * GS_BK, t * GS_BM, kEvictFirst);
          ++n_loaded;
          if (++stage == GS_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== X producer: waits for the step's input to be complete grid-wide =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      pdl_wait();  // the chain's first input comes from the previous kernel
      for (int s = 0; s < p.n_steps; ++s) {
        const ChainStep& st = p.steps[s];
        const int lo = gs_lo(cta, st.units, G), hi = gs_lo(cta + 1, st.units, G);
        if (lo >= hi) continue;
        tma_prefetch_desc(&st.tmX);
        if (st.has_norm) spin_until(p.sync + 2 * s + 1, p.M);
        else if (s > 0) spin_until(p.sync + 2 * (s - 1), G);
        fence_proxy_async_global();
        for (int u = lo; u < hi; ++u) {
          const int kb = u % st.kb_total;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_x[stage], S::kXBytes);
          tma_load_2d_hint(smem + stage * S::kStageBytes + S::kWBytes, &st.tmX, &full_x[stage], kb * GS_BK, 0, kEvictLast);
          if (++stage == GS_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int s = 0; s < p.n_steps; ++s) {
        const ChainStep& st = p.steps[s];
        const int lo = gs_lo(cta, st.units, G), hi = gs_lo(cta + 1, st.units, G);
        int u = lo;
        while (u < hi) {
          const int t = u / st.kb_total;
          const int seg_end = min(hi, (t + 1) * st.kb_total);
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
          for (int v = u; v < seg_end; ++v) {
            mbar_wait(&full_w[stage], phase);
            mbar_wait(&full_x[stage], phase);
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
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps: norm rows, split reduction, fused epilogue, step completion =====================
    pdl_wait();  // no global write before the previous kernel has fully finished
    const int q = warp & 3;
    const int etid = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int s = 0; s < p.n_steps; ++s) {
      const ChainStep& st = p.steps[s];
      if (st.has_norm && cta < p.M) {
        if (s > 0) {
          if (etid == 0) spin_until(p.sync + 2 * (s - 1), G);   // the rows are complete once every CTA finished step s-1
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        const T* xr = reinterpret_cast<const T*>(st.norm_src) + static_cast<int64_t>(cta) * st.norm_cols;
        T* yr = reinterpret_cast<T*>(st.norm_dst) + static_cast<int64_t>(cta) * st.norm_cols;
        chain_rmsnorm_row<T>(xr, reinterpret_cast<const T*>(st.norm_w), yr, st.norm_cols, st.norm_eps, etid, red);
        if (etid == 0) {
          __threadfence();
          red_release_add(p.sync + 2 * s + 1, 1);
        }
      }
      const int U = st.units;
      const int lo = gs_lo(cta, U, G), hi = gs_lo(cta + 1, U, G);
      GsOut out;
      out.D = st.D; out.ldd = st.ldd; out.bias = nullptr; out.residual = st.residual; out.ldr = st.ldr;
      out.M = p.M; out.N = st.N; out.epilogue = EPI_NONE; out.out_f32 = st.out_f32; out.kb_total = st.kb_total;
      out.partials = p.partials; out.counters = p.counters;
      int u = lo;
      while (u < hi) {
        const int t = u / st.kb_total;
        const int t0 = t * st.kb_total;
        const int seg_end = min(hi, t0 + st.kb_total);
        const bool whole = (u == t0) && (seg_end == t0 + st.kb_total);
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
        if (st.ek == 1) gs_finish_segment<T, BN, 1>(out, U, G, lo, t, whole, v, last_flag, q, lane, etid);
        else gs_finish_segment<T, BN, 0>(out, U, G, lo, t, whole, v, last_flag, q, lane, etid);
        u = seg_end;
      }
      // this CTA's share of step s is stored (including every tile it finished as the last arriver)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (etid == 0) {
        __threadfence();
        red_release_add(p.sync + 2 * s, 1);
      }
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
size_t chain_step_bytes() { return sizeof(ChainStep); }

// Fills one step (host copy).  W [N, K] (ldb), X [M, K] (lda) read through TMA; D / residual as in gemm_stream_run.
int chain_encode_step(void* host_step, int bn, const void* W, int64_t ldb, const void* X, int64_t lda, int M, int N, int K,
                      void* D, int64_t ldd, const void* residual, int64_t ldr, int ek, int out_f32, const void* norm_src,
                      const void* norm_w, void* norm_dst, int norm_cols, float norm_eps) {
  ChainStep* s = static_cast<ChainStep*>(host_step);
  memset(s, 0, sizeof(ChainStep));
  s->D = D; s->ldd = ldd; s->residual = residual; s->ldr = ldr;
  s->N = N; s->K = K; s->ek = ek; s->out_f32 = out_f32;
  s->num_t = N > 0 ? (N + GS_BM - 1) / GS_BM : 0;
  s->kb_total = K > 0 ? (K + GS_BK - 1) / GS_BK : 1;
  const long long units = static_cast<long long>(s->num_t) * s->kb_total;
  ULLAVA_REQUIRE(units * 1024 < (1ll << 31), "chain: %lld units exceed the 32-bit partition arithmetic", units);
  s->units = static_cast<int>(units);
  s->has_norm = norm_src != nullptr;
  s->norm_src = norm_src; s->norm_w = norm_w; s->norm_dst = norm_dst; s->norm_cols = norm_cols; s->norm_eps = norm_eps;
  if (ek == 1) ULLAVA_REQUIRE((N % 32) == 0, "chain: SILU_MUL needs N %% 32 == 0");
  if (norm_src != nullptr)
    ULLAVA_REQUIRE(norm_cols > 0 && norm_cols % 8 == 0 && norm_cols <= 8 * kChainMaxVec * 512, "chain: bad norm width %d", norm_cols);
  if (N > 0) {
    int st = encode_tmap_2d(&s->tmW, W, 2, K, N, ldb * 2, GS_BK, GS_BM, true);
    if (st) return st;
    st = encode_tmap_2d(&s->tmX, X, 2, K, M, lda * 2, GS_BK, bn, true);
    if (st) return st;
  }
  return OK;
}

// steps_dev: n_steps ChainStep records in device memory; sync_dev: 2 * n_steps ints, zeroed (on `stream`) by the caller
// before this launch.
int gemm_chain_run(Context* ctx, const void* steps_dev, int n_steps, int M, int dtype, int* sync_dev, cudaStream_t stream) {
  ULLAVA_REQUIRE(steps_dev && sync_dev && n_steps > 0 && M > 0 && M <= 32, "chain: bad arguments");
  ULLAVA_REQUIRE(dtype == DT_BF16 || dtype == DT_F16, "chain: 16-bit dtypes only");
  const int bn = M <= 16 ? 16 : 32;
  const int G = GS_CTAS_PER_SM * ctx->sm_count;
  ULLAVA_REQUIRE(M <= G, "chain: fewer CTAs (%d) than rows (%d)", G, M);
  const size_t need = kStreamCounterBytes + static_cast<size_t>(2) * G * GS_BM * bn * sizeof(float);
  if (need > ctx->workspace_bytes) {
    set_last_error("chain: workspace too small (%zu > %zu)", need, ctx->workspace_bytes);
    return ERR_WORKSPACE;
  }
  ChainParams p{};
  p.steps = static_cast<const ChainStep*>(steps_dev);
  p.n_steps = n_steps;
  p.M = M;
  p.counters = reinterpret_cast<int*>(ctx->workspace);
  p.partials = reinterpret_cast<float*>(static_cast<uint8_t*>(ctx->workspace) + kStreamCounterBytes);
  p.sync = sync_dev;
  p.lookahead = ctx->prefetch_units > 0 ? 2 * ctx->prefetch_units : 0;
  ctx->next_w = nullptr;
  const bool pdl = ctx->pdl != 0;
  int st;
#define ULLAVA_CHAIN(TT, BNN)                                                                                        \
  do {                                                                                                              \
    auto kern = gemm_chain_kernel<TT, BNN>;                                                                         \
    static SmemOptIn opt_in;                                                                                        \
    st = ensure_dynamic_smem(kern, ChainSmem<BNN>::kTotal, opt_in);                                                 \
    if (st == OK)                                                                                                   \
      st = check_cuda(launch_pdl(kern, dim3(G), dim3(GS_THREADS), ChainSmem<BNN>::kTotal, stream, pdl, p),          \
                      "gemm_chain_kernel launch");                                                                  \
  } while (0)
  if (dtype == DT_BF16) {
    if (bn == 16) ULLAVA_CHAIN(__nv_bfloat16, 16); else ULLAVA_CHAIN(__nv_bfloat16, 32);
  } else {
    if (bn == 16) ULLAVA_CHAIN(__half, 16); else ULLAVA_CHAIN(__half, 32);
  }
#undef ULLAVA_CHAIN
  if (st == OK) ctx->launches += 1;
  return st;
}

}  // namespace ullava
