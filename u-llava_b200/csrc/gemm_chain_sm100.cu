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
  int* counters;     // [num_t] arrival counters of this step's tiles, zeroed by the host side before every launch
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
  int* sync;         // [2 * n_steps]: sync[2s] = CTAs done with step s, sync[2s+1] = rows normalised for step s; zeroed
                     // by the host side before every launch
  int lookahead;     // W tiles pulled into L2 ahead of the ring
  unsigned long long* trace;  // optional [grid][n_steps][8] globaltimer stamps (tools/bench_chain.py TRACE=1), else NULL
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CHAIN_TRACE(slot) \
  do { if (p.trace) p.trace[(static_cast<size_t>(cta) * p.n_steps + s) * 8 + (slot)] = gtime(); } while (0)

static constexpr int kChainMaxVec = 4;  // norm_kernel's kMaxVec

// Stream-K range of CTA `cta` in a step: the step's units are cut over Gs = min(units, grid) CTAs, exactly like a
// stand-alone gemm_stream_kernel launch (grid = min(units, SMs)) -- every participating CTA owns at least one unit, which
// the split reduction relies on (contributors of a tile = a contiguous CTA interval); the other CTAs sit the step out.
struct ChainRange { int lo, hi, Gs; };
__device__ __forceinline__ ChainRange chain_range(int units, int cta, int G) {
  ChainRange r;
  r.Gs = units < G ? units : G;
  if (cta < r.Gs) {
    r.lo = gs_lo(cta, units, r.Gs);
    r.hi = gs_lo(cta + 1, units, r.Gs);
  } else {
    r.lo = r.hi = 0;
  }
  return r;
}

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
__device__ __forceinline__ uint4 chain_norm_vec(uint4 r, uint4 wv, float rstd) {
  // HF: weight * (x_fp32 * rstd).to(dtype) -- one rounding before the weight multiply (as norm_kernel)
  const uint32_t u[4] = {r.x, r.y, r.z, r.w}, wu[4] = {wv.x, wv.y, wv.z, wv.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack2<T>(u[j]);
    const float2 g = unpack2<T>(wu[j]);
    const float2 n = unpack2<T>(pack2<T>(f.x * rstd, f.y * rstd));
    o[j] = pack2<T>(g.x * n.x, g.y * n.y);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

template <typename T>
__device__ __forceinline__ float chain_sumsq(uint4 r) {
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack2<T>(u[j]);
    sq += f.x * f.x + f.y * f.y;
  }
  return sq;
}

template <typename T>
__device__ __forceinline__ void chain_rmsnorm_row(const T* xr, const T* __restrict__ w, T* yr, int cols, float eps, int etid,
                                                  float* red) {
  const int nvec = cols >> 3;
  int VT = ((nvec + 31) / 32) * 32;
  if (VT > 512) VT = 512;
  while (VT * kChainMaxVec < nvec) VT += 32;
  const int nvw = VT >> 5;                    // virtual warps
  const int w4 = etid >> 5, l = etid & 31;
  // x is rewritten by other CTAs inside this kernel: read it past L1 (__ldcg)
  if (nvec <= VT && nvw <= 16) {
    // the common shapes (one vector per virtual thread, cols <= 4096): the whole row and its weights in ONE round of
    // loads, kept in registers for the second pass (virtual thread vt = etid + 128 k handles vector vt in both passes)
    uint4 r[4], wv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int vi = etid + 128 * k;
      const bool ok = (w4 + 4 * k < nvw) && vi < nvec;
      r[k] = ok ? __ldcg(reinterpret_cast<const uint4*>(xr + vi * 8)) : make_uint4(0, 0, 0, 0);
      wv[k] = ok ? *reinterpret_cast<const uint4*>(w + vi * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (w4 + 4 * k < nvw) {
        const float sq = warp_sum(chain_sumsq<T>(r[k]));
        if (l == 0) red[w4 + 4 * k] = sq;
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float t = (l < nvw) ? red[l] : 0.f;
    t = warp_sum(t);
    const float rstd = rsqrtf(t / cols + eps);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int vi = etid + 128 * k;
      if ((w4 + 4 * k < nvw) && vi < nvec) *reinterpret_cast<uint4*>(yr + vi * 8) = chain_norm_vec<T>(r[k], wv[k], rstd);
    }
  } else {
    for (int vw = w4; vw < nvw; vw += 4) {
      const int vt = vw * 32 + l;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < kChainMaxVec; ++i) {
        const int vi = vt + i * VT;
        if (vi < nvec) sq += chain_sumsq<T>(__ldcg(reinterpret_cast<const uint4*>(xr + vi * 8)));
      }
      sq = warp_sum(sq);
      if (l == 0) red[vw] = sq;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float t = (l < nvw) ? red[l] : 0.f;
    t = warp_sum(t);
    const float rstd = rsqrtf(t / cols + eps);
    for (int vi = etid; vi < nvec; vi += 128)
      *reinterpret_cast<uint4*>(yr + vi * 8) = chain_norm_vec<T>(__ldcg(reinterpret_cast<const uint4*>(xr + vi * 8)),
                                                                 *reinterpret_cast<const uint4*>(w + vi * 8), rstd);
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");  // red[] is reused by the next row / step
}

// Split reduction of one tile segment, cooperative flavour.  In gemm_stream_kernel the LAST contributor of a tile sums
// every partial (a serial chain: wait for the last arrival, then ceil(n / 2) dependent L2 round trips, then the whole
// epilogue) while the other contributors are done -- fine when nobody may wait, but in the chain every CTA has to wait
// for the step to complete anyway, and that tail (8-15 us on the slowest CTA, traced) was the longest part of every
// step boundary.  Here all n contributors of a tile arrive at about the same time (same number of units each), so they
// share the work: everyone parks its partial, waits until all n have arrived (spin on the tile counter), then
// contributor j sums -- in CTA order, so the result is bit-identical -- and finishes the batch-column groups g with
// g % n == j: one round of loads per CTA, 1 / n of the epilogue each.  Every step has its own tile counters (zeroed by
// the host side with the grid-barrier state), so nothing has to be reset inside the kernel.
// (b) of the cooperative split reduction: the epilogue of the batch-column groups in `mine` of tile t from v[].
template <typename T, int BN, int EK>
__device__ __forceinline__ void chain_epilogue(const GsOut& p, int t, uint32_t mine, float (&v)[BN], int q, int lane, int etid) {
  if (mine == 0u) return;
  const T* resid = reinterpret_cast<const T*>(p.residual);
  const int n_row = t * GS_BM + etid;  // weight row = output column
  const bool n_ok = n_row < p.N;
  if constexpr (EK == 1) {
    // SiLU(gate) * up: lanes 0-15 hold gate rows, lanes 16-31 the partner up rows (packed weight layout)
    const int oc = ((t * GS_BM + q * 32) >> 1) + (lane & 15);
    T* out = reinterpret_cast<T*>(p.D) + oc;
    const bool st_ok = lane < 16 && n_ok;
#pragma unroll
    for (int i = 0; i < BN; ++i) {
      if (mine & (1u << (i >> 2))) {       // CTA-uniform: all lanes take part in the shuffle
        const float up = __shfl_xor_sync(0xffffffffu, v[i], 16);
        const float gte = v[i];
        const float o = __fdividef(gte, 1.f + __expf(fminf(-gte, 80.f))) * up;
        if (st_ok && i < p.M) out[static_cast<int64_t>(i) * p.ldd] = T16<T>::from_f(o);
      }
    }
  } else if (n_ok) {
    if (resid != nullptr) {
      float rr[BN];
#pragma unroll
      for (int i = 0; i < BN; ++i)
        rr[i] = (i < p.M && (mine & (1u << (i >> 2)))) ? T16<T>::to_f(resid[static_cast<int64_t>(i) * p.ldr + n_row]) : 0.f;
#pragma unroll
      for (int i = 0; i < BN; ++i) v[i] += rr[i];
    }
    if (p.out_f32) {
      float* out = reinterpret_cast<float*>(p.D) + n_row;
#pragma unroll
      for (int i = 0; i < BN; ++i)
        if (i < p.M && (mine & (1u << (i >> 2)))) out[static_cast<int64_t>(i) * p.ldd] = v[i];
    } else {
      T* out = reinterpret_cast<T*>(p.D) + n_row;
#pragma unroll
      for (int i = 0; i < BN; ++i)
        if (i < p.M && (mine & (1u << (i >> 2)))) out[static_cast<int64_t>(i) * p.ldd] = T16<T>::from_f(v[i]);
    }
  }
}

// (a) park the partial of a cut tile and announce it -- WITHOUT waiting: a CTA parks every partial segment of the step
// first and only then waits for the other contributors (waiting per segment would chain CTA k's first segment to CTA
// k-1's last one across the whole grid).
template <int BN>
__device__ __forceinline__ void chain_park(const GsOut& p, int lo, int t, const float (&v)[BN], int etid) {
  const int t0 = t * p.kb_total;
  const int which = (lo >= t0) ? 0 : 1;
  float4* slot = reinterpret_cast<float4*>(p.partials) + static_cast<size_t>(2 * blockIdx.x + which) * (GS_BM * BN / 4) + etid;
#pragma unroll
  for (int i = 0; i < BN / 4; ++i) slot[i * GS_BM] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  asm volatile("bar.sync 1, 128;" ::: "memory");   // every thread's partial stores precede thread 0's release
  if (etid == 0) red_release_add(p.counters + t, 1);
}

// Sums, for up to GI column groups of this CTA (g = j + gi * n), the partials of contributors c_first .. c_last in CTA
// order; CI loads per group and round, all rounds' loads of a group issued before the first add.
template <int BN, int GI, int CI>
__device__ __forceinline__ void chain_sum_groups(const GsOut& p, int U, int G, int t0, int c_first, int c_last, int j, int n,
                                                 int etid, float (&v)[BN]) {
  constexpr int NG = BN / 4;
  const float4* base = reinterpret_cast<const float4*>(p.partials) + etid;
  float4 acc[GI];
#pragma unroll
  for (int gi = 0; gi < GI; ++gi) acc[gi] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c0 = c_first; c0 <= c_last; c0 += CI) {     // one round unless n > CI
    float4 f[GI][CI];
#pragma unroll
    for (int gi = 0; gi < GI; ++gi) {
      const int g = j + gi * n;
#pragma unroll
      for (int ci = 0; ci < CI; ++ci) {
        const int c = c0 + ci;
        const bool ok = g < NG && c <= c_last;
        const int slot = 2 * c + ((gs_lo(c, U, G) >= t0) ? 0 : 1);   // contributor c's slot: 0 when its range starts inside the tile
        f[gi][ci] = ok ? __ldcg(base + static_cast<size_t>(slot) * (GS_BM * BN / 4) + g * GS_BM) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int gi = 0; gi < GI; ++gi) {
#pragma unroll
      for (int ci = 0; ci < CI; ++ci) {
        if (c0 + ci <= c_last) {
          acc[gi].x += f[gi][ci].x; acc[gi].y += f[gi][ci].y; acc[gi].z += f[gi][ci].z; acc[gi].w += f[gi][ci].w;
        }
      }
    }
  }
#pragma unroll
  for (int gi = 0; gi < GI; ++gi) {
    const int g = j + gi * n;
    // v[] is indexed with a run-time g: select statically
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) {
      if (gg == g) { v[4 * gg] = acc[gi].x; v[4 * gg + 1] = acc[gi].y; v[4 * gg + 2] = acc[gi].z; v[4 * gg + 3] = acc[gi].w; }
    }
  }
}

// (c) wait until all n contributors of tile t have parked, sum -- in CTA order, bit-identical to gs_finish_segment --
// the column groups g with g % n == j (j = this CTA's rank among the contributors) and finish them.  Every load of the
// reduction (about NG + n float4 per thread whatever n is) and the residual values are issued in one round.
template <typename T, int BN, int EK>
__device__ __forceinline__ void chain_reduce(const GsOut& p, int U, int G, int t, int q, int lane, int etid) {
  constexpr int NG = BN / 4;
  const int t0 = t * p.kb_total;
  const int c_first = gs_owner(t0, U, G), c_last = gs_owner(t0 + p.kb_total - 1, U, G);
  const int n = c_last - c_first + 1, j = static_cast<int>(blockIdx.x) - c_first;
  uint32_t mine = 0u;
#pragma unroll
  for (int g = 0; g < NG; ++g)
    if (g % n == j) mine |= 1u << g;
  // residual values of my columns: independent of the other contributors, so they travel under the wait
  const T* resid = reinterpret_cast<const T*>(p.residual);
  const int n_row = t * GS_BM + etid;
  float rr[BN];
  if constexpr (EK != 1) {
#pragma unroll
    for (int i = 0; i < BN; ++i)
      rr[i] = (resid != nullptr && n_row < p.N && i < p.M && (mine & (1u << (i >> 2))))
                  ? T16<T>::to_f(resid[static_cast<int64_t>(i) * p.ldr + n_row]) : 0.f;
  }
  if (etid == 0) spin_until(p.counters + t, n);
  asm volatile("bar.sync 1, 128;" ::: "memory");   // thread 0's acquire precedes every thread's loads
  float v[BN];
#pragma unroll
  for (int i = 0; i < BN; ++i) v[i] = 0.f;
  if (mine != 0u) {
    // my groups are j, j + n, j + 2n, ...: at most GI of them, each summed over the n contributors.  (GI, CI) variants
    // keep the loads of ONE round in registers with compile-time indices: n <= 2 -> 4 x 2, n == 3 -> 3 x 3,
    // n <= 8 -> 2 x 8, else 1 x 16 (n > 16 cannot happen with >= 8 k-blocks per CTA and <= 172 per tile; looped anyway)
    if (n <= 2) chain_sum_groups<BN, 4, 2>(p, U, G, t0, c_first, c_last, j, n, etid, v);
    else if (n == 3) chain_sum_groups<BN, 3, 3>(p, U, G, t0, c_first, c_last, j, n, etid, v);
    else if (n <= 8) chain_sum_groups<BN, 2, 8>(p, U, G, t0, c_first, c_last, j, n, etid, v);
    else chain_sum_groups<BN, 1, 16>(p, U, G, t0, c_first, c_last, j, n, etid, v);
  }
  if constexpr (EK != 1) {
#pragma unroll
    for (int i = 0; i < BN; ++i) v[i] += rr[i];
  }
  GsOut pp = p;
  pp.residual = nullptr;                           // already added
  chain_epilogue<T, BN, EK>(pp, t, mine, v, q, lane, etid);
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
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // L2 look-ahead: a cursor (ps, pu) over this CTA's units of the WHOLE chain, in load order, kept in front of the
      // unit being loaded; n_loaded / n_pf count units passed by the ring loads / by the cursor
      int ps = -1, pu = 0, p_hi = 0, n_loaded = 0, n_pf = 0;
      auto pf_advance = [&]() {   // moves the cursor to this CTA's next unit; false at the end of the chain
        ++pu;
        while (pu >= p_hi) {
          if (++ps >= p.n_steps) return false;
          const ChainRange r = chain_range(p.steps[ps].units, cta, G);
          pu = r.lo;
          p_hi = r.hi;
        }
        return true;
      };
      pu = -1;
      bool pf_live = p.lookahead > 0 && pf_advance();   // cursor on the first unit
      for (int s = 0; s < p.n_steps; ++s) {
        const ChainStep& st = p.steps[s];
        const ChainRange rg = chain_range(st.units, cta, G);
        const int lo = rg.lo, hi = rg.hi;
        if (lo < hi) tma_prefetch_desc(&st.tmW);
        CHAIN_TRACE(5);
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
        CHAIN_TRACE(0);
      }
    }
  } else if (warp == 3) {
    // ===================== X producer: waits for the step's input to be complete grid-wide =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      pdl_wait();  // the chain's first input comes from the previous kernel
      for (int s = 0; s < p.n_steps; ++s) {
        const ChainStep& st = p.steps[s];
        const ChainRange rg = chain_range(st.units, cta, G);
        const int lo = rg.lo, hi = rg.hi;
        if (lo >= hi) continue;
        tma_prefetch_desc(&st.tmX);
        if (st.has_norm) spin_until(p.sync + 2 * s + 1, p.M);
        else if (s > 0) spin_until(p.sync + 2 * (s - 1), G);
        fence_proxy_async_global();
        CHAIN_TRACE(1);
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
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int s = 0; s < p.n_steps; ++s) {
        const ChainStep& st = p.steps[s];
        const ChainRange rg = chain_range(st.units, cta, G);
        const int lo = rg.lo, hi = rg.hi;
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
            if (v == lo) CHAIN_TRACE(2);
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
        CHAIN_TRACE(3);
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
        if (etid == 0) CHAIN_TRACE(6);
        chain_rmsnorm_row<T>(xr, reinterpret_cast<const T*>(st.norm_w), yr, st.norm_cols, st.norm_eps, etid, red);
        if (etid == 0) {
          red_release_add(p.sync + 2 * s + 1, 1);   // (the bar.sync at the end of the norm ordered every thread's stores)
          CHAIN_TRACE(7);
        }
      }
      const int U = st.units;
      const ChainRange rg = chain_range(U, cta, G);
      const int lo = rg.lo, hi = rg.hi, Gs = rg.Gs;
      GsOut out;
      out.D = st.D; out.ldd = st.ldd; out.bias = nullptr; out.residual = st.residual; out.ldr = st.ldr;
      out.M = p.M; out.N = st.N; out.epilogue = EPI_NONE; out.out_f32 = st.out_f32; out.kb_total = st.kb_total;
      out.partials = p.partials; out.counters = st.counters;
      int pend[2], n_pend = 0;   // cut tiles of this CTA's range (its first and / or last segment)
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
        if (whole) {
          if (st.ek == 1) chain_epilogue<T, BN, 1>(out, t, (1u << (BN / 4)) - 1u, v, q, lane, etid);
          else chain_epilogue<T, BN, 0>(out, t, (1u << (BN / 4)) - 1u, v, q, lane, etid);
        } else {
          chain_park<BN>(out, lo, t, v, etid);
          pend[n_pend++] = t;
        }
        u = seg_end;
      }
      for (int k = 0; k < n_pend; ++k) {
        if (st.ek == 1) chain_reduce<T, BN, 1>(out, U, Gs, pend[k], q, lane, etid);
        else chain_reduce<T, BN, 0>(out, U, Gs, pend[k], q, lane, etid);
      }
      // this CTA's share of step s is stored (including every tile it finished as the last arriver)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (etid == 0) {
        red_release_add(p.sync + 2 * s, 1);
        CHAIN_TRACE(4);
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
                      const void* norm_w, void* norm_dst, int norm_cols, float norm_eps, int* counters_dev) {
  ChainStep* s = static_cast<ChainStep*>(host_step);
  memset(s, 0, sizeof(ChainStep));
  s->counters = counters_dev;
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
  p.partials = reinterpret_cast<float*>(static_cast<uint8_t*>(ctx->workspace) + kStreamCounterBytes);
  p.sync = sync_dev;
  p.lookahead = ctx->prefetch_units > 0 ? 2 * ctx->prefetch_units : 0;
  p.trace = static_cast<unsigned long long*>(ctx->chain_trace);
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
