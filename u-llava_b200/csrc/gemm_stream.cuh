// Shared device code of the weight-streaming GEMMs (gemm_stream_sm100.cu, gemm_chain_sm100.cu): tile geometry, the
// stream-K partition, and the split reduction + fused epilogue of one tile segment.
#pragma once
#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

static constexpr int GS_BM = 128;     // weight rows per tile (UMMA M)
static constexpr int GS_BK = 64;      // one SWIZZLE_128B row
static constexpr int GS_STAGES = 8;     // 160 KB: one CTA per SM (PDL then maps the next GEMM's CTAs 1:1 onto SMs),
                                        // 128 KB of W in flight per SM.  Measured alternatives (decode-chain bench, B = 32):
                                        // 5 stages x 2 CTAs/SM with 148 ranges 139 us/layer, with 296 ranges 151 us/layer,
                                        // this configuration 109 us/layer.
static constexpr int GS_CTAS_PER_SM = 1;
static constexpr int GS_THREADS = 256;

struct StreamParams {
  void* D;
  int64_t ldd;
  const void* bias;
  const void* residual;
  int64_t ldr;
  int M, N, K;        // M = valid batch rows (<= BN), N = weight rows
  int epilogue, out_f32;
  int num_t;          // weight tiles
  int kb_total;       // k-blocks per tile
  int units;          // num_t * kb_total (host guarantees units * gridDim.x < 2^31)
  float* partials;    // [2 * gridDim.x][BN / 4][128] float4 (fp32 partial tiles, row-interleaved for coalescing)
  int* counters;      // [num_t], zero between launches
  // next-weight L2 prefetch (tmP): partition of the NEXT GEMM's units over ITS grid; this CTA pulls units
  // [lo' + pf_skip, lo' + pf_skip + pf_count) of the range CTA blockIdx.x of the next kernel will stream
  int pf_kb_total, pf_units, pf_grid, pf_skip, pf_count;
};

__device__ __forceinline__ int gs_lo(int c, int U, int G) {
  return static_cast<int>(static_cast<uint32_t>(c) * static_cast<uint32_t>(U) / static_cast<uint32_t>(G));
}
// CTA that owns unit u:  largest c with floor(c*U/G) <= u
__device__ __forceinline__ int gs_owner(int u, int U, int G) {
  return static_cast<int>((static_cast<uint32_t>(u + 1) * static_cast<uint32_t>(G) + static_cast<uint32_t>(U) - 1u) /
                          static_cast<uint32_t>(U)) - 1;
}

template <int BN>
struct StreamSmem {
  static constexpr int kWBytes = GS_BM * GS_BK * 2;
  static constexpr int kXBytes = BN * GS_BK * 2;
  static constexpr int kStageBytes = kWBytes + kXBytes;
  static constexpr int kBarOffset = GS_STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * GS_STAGES + 4) * 8 + 32 + 1024;
};

__device__ __forceinline__ float gs_act(float v, int epi) {
  switch (epi) {
    case EPI_RELU: return fmaxf(v, 0.f);
    case EPI_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    case EPI_QUICK_GELU: return __fdividef(v, 1.f + __expf(fminf(-1.702f * v, 80.f)));
    default: return v;
  }
}

// Everything after the accumulator of one segment (the k-blocks [u, seg_end) of weight tile t that this CTA streamed)
// has been read out of TMEM into v[]: park the partial and find out whether this CTA is the tile's last contributor
// (stream-K split), sum all contributions in CTA order, run the fused epilogue.  Shared by gemm_stream_kernel (one GEMM
// per launch) and gemm_chain_kernel (the GEMMs of a decode layer in one persistent kernel).
// EK: 0 = bias / residual only, 1 = SiLU(gate) * up, 2 = bias + ReLU / GELU / quick-GELU + residual
struct GsOut {
  void* D;
  int64_t ldd;
  const void* bias;
  const void* residual;
  int64_t ldr;
  int M, N;
  int epilogue, out_f32;
  int kb_total;
  float* partials;    // [2 * gridDim.x][BN / 4][128] float4 (fp32 partial tiles, row-interleaved for coalescing)
  int* counters;      // [num_t], zero between uses
};

template <typename T, int BN, int EK>
__device__ __forceinline__ void gs_finish_segment(const GsOut& p, int U, int G, int lo, int t, bool whole, float (&v)[BN],
                                                  volatile int* last_flag, int q, int lane, int etid) {
  const T* bias = reinterpret_cast<const T*>(p.bias);
  const T* resid = reinterpret_cast<const T*>(p.residual);
  const int t0 = t * p.kb_total;
  bool do_epilogue = whole;
  if (!whole) {
    // park the partial, then find out whether this CTA is the last contributor of tile t
    const int which = (lo >= t0) ? 0 : 1;  // 0: my range starts inside the tile, 1: it only ends there
    // slot layout [BN/4][128 rows] of float4: a warp's store / load covers 512 contiguous bytes
    float4* mine = reinterpret_cast<float4*>(p.partials) + static_cast<size_t>(2 * blockIdx.x + which) * (GS_BM * BN / 4) + etid;
#pragma unroll
    for (int i = 0; i < BN / 4; ++i) mine[i * GS_BM] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    // release/acquire through the tile counter: the CTA barrier orders every thread's partial stores before
    // thread 0's gpu-scope release, and the last arriver's acquire before every thread's loads (cumulativity)
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int c_first = gs_owner(t0, U, G), c_last = gs_owner(t0 + p.kb_total - 1, U, G);
    if (etid == 0) {
      int old;
      asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(p.counters + t) : "memory");
      *last_flag = (old == c_last - c_first) ? 1 : 0;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    do_epilogue = (*last_flag != 0);
    if (do_epilogue) {
      // sum ALL contributions (this CTA's own included) in CTA order, from memory: the result does not depend
      // on which contributor happened to arrive last, so greedy decode stays bit-reproducible
#pragma unroll
      for (int i = 0; i < BN; ++i) v[i] = 0.f;
      // two contributors in flight per step (one L2 round trip per pair instead of per contributor); the adds
      // still happen in CTA order
      int c_lo = gs_lo(c_first, U, G);
#pragma unroll 1
      for (int c = c_first; c <= c_last; c += 2) {
        const int w0 = (c_lo >= t0) ? 0 : 1;
        c_lo = gs_lo(c + 1, U, G);
        const bool two = c + 1 <= c_last;
        const int w1 = (c_lo >= t0) ? 0 : 1;
        c_lo = gs_lo(c + 2, U, G);
        const float4* o0 = reinterpret_cast<const float4*>(p.partials) + static_cast<size_t>(2 * c + w0) * (GS_BM * BN / 4) + etid;
        const float4* o1 = two ? reinterpret_cast<const float4*>(p.partials) + static_cast<size_t>(2 * (c + 1) + w1) * (GS_BM * BN / 4) + etid : o0;
        float4 f[BN / 4], g[BN / 4];
#pragma unroll
        for (int i = 0; i < BN / 4; ++i) f[i] = __ldcg(o0 + i * GS_BM);
#pragma unroll
        for (int i = 0; i < BN / 4; ++i) g[i] = __ldcg(o1 + i * GS_BM);
#pragma unroll
        for (int i = 0; i < BN / 4; ++i) {
          v[4 * i] += f[i].x; v[4 * i + 1] += f[i].y; v[4 * i + 2] += f[i].z; v[4 * i + 3] += f[i].w;
        }
        if (two) {
#pragma unroll
          for (int i = 0; i < BN / 4; ++i) {
            v[4 * i] += g[i].x; v[4 * i + 1] += g[i].y; v[4 * i + 2] += g[i].z; v[4 * i + 3] += g[i].w;
          }
        }
      }
      if (etid == 0) p.counters[t] = 0;  // ready for the next launch / graph replay
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // last_flag may be rewritten by the next segment
  }

  if (do_epilogue) {
    // Compact, branch-free-per-element stores: this code runs once or twice per CTA, i.e. always from a cold
    // instruction cache, so its size is what it costs.
    const int n = t * GS_BM + etid;  // weight row = output column
    const bool n_ok = n < p.N;
    if (bias != nullptr && n_ok) {
      const float bv = T16<T>::to_f(bias[n]);
#pragma unroll
      for (int i = 0; i < BN; ++i) v[i] += bv;
    }
    if constexpr (EK == 1) {
      // SiLU(gate) * up.  W rows are packed in blocks of 32 = 16 gate rows + the 16 matching up rows: lanes 0-15
      // hold gate, lanes 16-31 the partner up value; output column = block * 16 + lane
      const int oc = ((t * GS_BM + q * 32) >> 1) + (lane & 15);
      T* out = reinterpret_cast<T*>(p.D) + oc;
      const bool st_ok = lane < 16 && n_ok;
#pragma unroll
      for (int i = 0; i < BN; ++i) {
        const float up = __shfl_xor_sync(0xffffffffu, v[i], 16);
        const float g = v[i];
        const float o = __fdividef(g, 1.f + __expf(fminf(-g, 80.f))) * up;
        if (st_ok && i < p.M) out[static_cast<int64_t>(i) * p.ldd] = T16<T>::from_f(o);
      }
    } else {
      if constexpr (EK == 2) {
#pragma unroll
        for (int i = 0; i < BN; ++i) v[i] = gs_act(v[i], p.epilogue);
      }
      if (n_ok) {
        // D may alias the residual (in-place x += f(x)): read every residual value before the first store so
        // the loads are issued back to back instead of one full latency per row
        if (resid != nullptr) {
          float rr[BN];
#pragma unroll
          for (int i = 0; i < BN; ++i) rr[i] = i < p.M ? T16<T>::to_f(resid[static_cast<int64_t>(i) * p.ldr + n]) : 0.f;
#pragma unroll
          for (int i = 0; i < BN; ++i) v[i] += rr[i];
        }
        if (p.out_f32) {
          float* out = reinterpret_cast<float*>(p.D) + n;
#pragma unroll
          for (int i = 0; i < BN; ++i)
            if (i < p.M) out[static_cast<int64_t>(i) * p.ldd] = v[i];
        } else {
          T* out = reinterpret_cast<T*>(p.D) + n;
#pragma unroll
          for (int i = 0; i < BN; ++i)
            if (i < p.M) out[static_cast<int64_t>(i) * p.ldd] = T16<T>::from_f(v[i]);
        }
      }
    }
  }
}

}  // namespace ullava
