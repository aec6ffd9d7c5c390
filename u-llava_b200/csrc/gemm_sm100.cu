// tcgen05 GEMM for the u-LLaVA hot path:  D[M,N] = epi(A[M,K] * B[N,K]^T)
//
// Replaces every nn.Linear on the path (reference call sites: CLIP q/k/v/out/fc1/fc2
// hf:models/clip/modeling_clip.py:282-351, LLaMA q/k/v/o/gate/up/down
// hf:models/llama/modeling_llama.py:171-291, projector models/ullava_core.py:117-129,228,
// lm_head models/ullava_core.py:325, seg/det projectors models/ullava.py:83-132).
//
// Design (B200, sm_100a):
//   * persistent CTAs (one per SM), static tile scheduler with M-grouped rasterisation so
//     that the CTAs of one wave share A/B tiles through the 126 MB L2;
//   * warp-specialised: warp0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B, K-major
//     64-element slabs), warp1 = single-thread tcgen05.mma issuer, warp2 = TMEM allocator,
//     warps4-11 = epilogue (tcgen05.ld -> registers -> fused bias/activation/residual -> global), two warps per
//     TMEM lane quadrant; the activation is a template parameter and uses MUFU-based GELU / sigmoid so that the
//     epilogue of a 128 x 256 tile stays well under the tile's MMA time (a libm erff epilogue took 3.7x the MMA time);
//   * 128 x BN fp32 accumulator lives in TMEM, double buffered (2*BN columns) so the epilogue
//     of tile i overlaps the MMAs of tile i+1;
//   * STAGES-deep shared-memory ring guarded by full/empty mbarriers; tcgen05.commit releases a
//     slot when the MMAs that read it have retired;
//   * gemm_tcgen05_pair_kernel: the same pipeline on CTA pairs (cta_group::2, clusters of 2 SMs) with 256 x 256
//     tiles -- every product with >= 148 pair tiles (all ViT / prefill / SAM-encoder GEMMs) takes it;
//   * rasterisation group chosen per shape: when B does not fit in L2 beside the A panel a taller M group cuts the
//     DRAM re-reads of B;
//   * small-M (decode, M <= 32) products go to the weight-streaming kernel of gemm_stream_sm100.cu; force_splits keeps
//     a split-K flavour of the large-M kernel (fp32 workspace + splitk_reduce_kernel) for tests and odd shapes.
#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

static constexpr int BM = 128;       // UMMA M (rows of A per tile)
static constexpr int BK = 64;        // 64 x 16-bit = 128 B = one SWIZZLE_128B row
static constexpr int UMMA_K = 16;    // fixed for 16-bit inputs
static constexpr int kDefaultGroupM = 8;    // rasterisation group (ctx->group_m overrides)
static constexpr int kGemmThreads = 384;
static constexpr int kEpiWarp0 = 4;   // first epilogue warp
static constexpr int kEpiWarps = 8;   // two warps per TMEM lane quadrant, alternating 32-column chunks

struct GemmKernelParams {
  void* D;
  int64_t ldd;
  const void* bias;
  const void* residual;
  int64_t ldr;
  int M, N, K;
  int epilogue;
  int out_f32;
  int num_m, num_n;   // tile counts
  int kb_total;       // ceil(K / 64)
  int kb_per_split;   // k-blocks handled by one split
  int splits;         // >1: D is an fp32 workspace [splits][M][N], epilogue deferred
  int group_m;        // rasterisation group: M tiles swept together over N (their A panel stays L2 resident)
  uint64_t hint_a, hint_b;  // L2 eviction priority of the A / B tile loads (CTA-pair kernel)
  RopeFuse rope;            // EPI_QKV_ROPE only
  int tma_store;            // CTA-pair kernel: D leaves through shared memory + cp.async.bulk.tensor (16-bit outputs)
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem_ptr
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 4) * 8 + 16 + 1024 /*alignment slack*/;
};

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int GROUP_M, int& split, int& m_blk,
                                            int& n_blk) {
  const int per_split = num_m * num_n;
  split = t / per_split;
  int r = t - split * per_split;
  const int group = GROUP_M * num_n;
  const int g = r / group;
  const int first_m = g * GROUP_M;
  const int gm = min(num_m - first_m, GROUP_M);
  r -= g * group;
  m_blk = first_m + r % gm;
  n_blk = r / gm;
}

template <int EPI>
__device__ __forceinline__ float apply_act(float v) {
  if constexpr (EPI == EPI_RELU) return fmaxf(v, 0.f);
  else if constexpr (EPI == EPI_GELU) return fast_gelu(v);
  else if constexpr (EPI == EPI_QUICK_GELU) return fast_x_sigmoid(v, 1.702f * 1.4426950408889634f);
  else return v;
}
// runtime flavour for the (tiny) split-K reduce kernel
__device__ __forceinline__ float apply_act_rt(float v, int epi) {
  switch (epi) {
    case EPI_RELU: return fmaxf(v, 0.f);
    case EPI_GELU: return fast_gelu(v);
    case EPI_QUICK_GELU: return fast_x_sigmoid(v, 1.702f * 1.4426950408889634f);
    default: return v;
  }
}

template <typename T, int EPI, int CH>
__device__ __forceinline__ void epilogue_store(float (&v)[CH], const GemmKernelParams& p, const T* bias, const T* resid,
                                               bool split_mode, int split, int row, int n0) {

        if (split_mode) {
          float* out = reinterpret_cast<float*>(p.D) + (static_cast<int64_t>(split) * p.M + row) * p.N + n0;
          if (n0 + CH <= p.N && (p.N & 3) == 0) {
#pragma unroll
            for (int i = 0; i < CH; i += 4) *reinterpret_cast<float4*>(out + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
            for (int i = 0; i < CH; ++i)
              if (n0 + i < p.N) out[i] = v[i];
          }
          return;
        }

        const bool full_chunk = (n0 + CH <= p.N);
        if (bias != nullptr) {
          if (full_chunk && ((reinterpret_cast<uintptr_t>(bias + n0) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < CH; i += 8) {
              uint4 b = *reinterpret_cast<const uint4*>(bias + n0 + i);
              float2 f;
              f = unpack2<T>(b.x); v[i] += f.x; v[i + 1] += f.y;
              f = unpack2<T>(b.y); v[i + 2] += f.x; v[i + 3] += f.y;
              f = unpack2<T>(b.z); v[i + 4] += f.x; v[i + 5] += f.y;
              f = unpack2<T>(b.w); v[i + 6] += f.x; v[i + 7] += f.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < CH; ++i)
              if (n0 + i < p.N) v[i] += T16<T>::to_f(bias[n0 + i]);
          }
        }

        if constexpr (EPI == EPI_SILU_MUL) {
          // weight rows are packed so that every 32-column chunk holds 16 gate columns followed by
          // the 16 matching up columns; output has N/2 columns.
          if constexpr (CH == 32) {
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              o[i] = fast_x_sigmoid(v[i], 1.4426950408889634f) * v[16 + i];
            }
            const int on0 = n0 >> 1;
            if (p.out_f32) {
              float* out = reinterpret_cast<float*>(p.D) + static_cast<int64_t>(row) * p.ldd + on0;
#pragma unroll
              for (int i = 0; i < 16; ++i) out[i] = o[i];
            } else {
              T* out = reinterpret_cast<T*>(p.D) + static_cast<int64_t>(row) * p.ldd + on0;
              if (resid != nullptr) {
                const T* rr = resid + static_cast<int64_t>(row) * p.ldr + on0;
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] += T16<T>::to_f(rr[i]);
              }
              if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
#pragma unroll
                for (int i = 0; i < 16; i += 8) {
                  uint4 w;
                  w.x = pack2<T>(o[i], o[i + 1]);
                  w.y = pack2<T>(o[i + 2], o[i + 3]);
                  w.z = pack2<T>(o[i + 4], o[i + 5]);
                  w.w = pack2<T>(o[i + 6], o[i + 7]);
                  *reinterpret_cast<uint4*>(out + i) = w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) out[i] = T16<T>::from_f(o[i]);
              }
            }
          }
          return;
        }

        if constexpr (EPI != EPI_NONE) {
#pragma unroll
          for (int i = 0; i < CH; ++i) v[i] = apply_act<EPI>(v[i]);
        }

        if (resid != nullptr) {
          const T* rr = resid + static_cast<int64_t>(row) * p.ldr + n0;
          if (full_chunk && ((reinterpret_cast<uintptr_t>(rr) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < CH; i += 8) {
              uint4 b = *reinterpret_cast<const uint4*>(rr + i);
              float2 f;
              f = unpack2<T>(b.x); v[i] += f.x; v[i + 1] += f.y;
              f = unpack2<T>(b.y); v[i + 2] += f.x; v[i + 3] += f.y;
              f = unpack2<T>(b.z); v[i + 4] += f.x; v[i + 5] += f.y;
              f = unpack2<T>(b.w); v[i + 6] += f.x; v[i + 7] += f.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < CH; ++i)
              if (n0 + i < p.N) v[i] += T16<T>::to_f(rr[i]);
          }
        }

        if (p.out_f32) {
          float* out = reinterpret_cast<float*>(p.D) + static_cast<int64_t>(row) * p.ldd + n0;
          if (full_chunk && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < CH; i += 4) *reinterpret_cast<float4*>(out + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < CH; ++i)
              if (n0 + i < p.N) out[i] = v[i];
          }
        } else {
          T* out = reinterpret_cast<T*>(p.D) + static_cast<int64_t>(row) * p.ldd + n0;
          if (full_chunk && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < CH; i += 8) {
              uint4 w;
              w.x = pack2<T>(v[i], v[i + 1]);
              w.y = pack2<T>(v[i + 2], v[i + 3]);
              w.z = pack2<T>(v[i + 4], v[i + 5]);
              w.w = pack2<T>(v[i + 6], v[i + 7]);
              *reinterpret_cast<uint4*>(out + i) = w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < CH; ++i)
              if (n0 + i < p.N) out[i] = T16<T>::from_f(v[i]);
          }
        }
      }

template <typename T, int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmKernelParams p) {
  using S = GemmSmem<BN, STAGES>;
  constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t kIdesc = make_idesc_f16(T16<T>::kUmmaFormat, BM, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m * p.num_n * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kEpiWarps * 32);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_ptr, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int split, m_blk, n_blk;
        tile_coords(t, p.num_m, p.num_n, p.group_m, split, m_blk, n_blk);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          mbar_expect_tx(&full_bar[stage], S::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) {
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
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int split, m_blk, n_blk;
        tile_coords(t, p.num_m, p.num_n, p.group_m, split, m_blk, n_blk);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint64_t a_desc = make_kmajor_sw128_desc(sa);
          const uint64_t b_desc = make_kmajor_sw128_desc(sa + S::kABytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_f16<1>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit<1>(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit<1>(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue (8 warps: TMEM lane quadrant = warp % 4, column chunks alternate) =====================
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const T* bias = reinterpret_cast<const T*>(p.bias);
    const T* resid = reinterpret_cast<const T*>(p.residual);
    const bool split_mode = p.splits > 1;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int split, m_blk, n_blk;
      tile_coords(t, p.num_m, p.num_n, p.group_m, split, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
      constexpr int CH = (BN >= 32) ? 32 : 16;  // columns per tcgen05.ld
#pragma unroll 1
      for (int c = half; c < BN / CH; c += 2) {
        float v[CH];
        if constexpr (CH == 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        } else {
          uint32_t r[16];
          tmem_ld_32x16(taddr + c * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        }
        const int n0 = n_blk * BN + c * CH;
        if (row_ok && n0 < p.N) epilogue_store<T, EPI, CH>(v, p, bias, resid, split_mode, split, row, n0);
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
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

// -----------------------------------------------------------------------------
// CTA-pair flavour (tcgen05 cta_group::2): one 256 x 256 output tile per PAIR of SMs.
// Each CTA of the pair stages its own 128 rows of A and HALF of the B tile (128 weight rows); the leader's
// single MMA thread issues M = 256 instructions that read both halves of B from the two shared memories and
// write 128 accumulator rows into each CTA's TMEM.  Per output tile each SM therefore moves (128 + 128) x K
// operand bytes instead of (128 + 256) x K: 1/3 less L2 -> SM traffic for the same math, which is what bounds the
// 128 x 256 single-CTA tile (85 FLOP per operand byte -> ~19 TB/s of L2 bandwidth at tensor peak).
// Barriers: TMA of both CTAs completes on the LEADER's full barrier; tcgen05.commit multicasts to the empty /
// tmem_full barriers of both CTAs; the epilogue warps of both CTAs arrive on the leader's tmem_empty barrier.
// -----------------------------------------------------------------------------
static constexpr int kPairStages = 6;   // 6 x (16 KB A + 16 KB B half) = 192 KB
static constexpr int kPairBM = 256, kPairBN = 256;

struct PairSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (kPairBN / 2) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kPairStages * kStageBytes;
  // epilogue staging for the TMA store: 8 warps x 2 buffers x (32 rows x 64 B), SWIZZLE_64B
  static constexpr int kStoreOffset = kBarOffset + 1024;
  static constexpr int kStoreBytes = kEpiWarps * 2 * 2048;
  static constexpr int kTotal = kStoreOffset + kStoreBytes + 1024;
  static_assert((2 * kPairStages + 4) * 8 + 16 <= 1024, "barrier block");
};

// EPI_QKV_ROPE: the prefill qkv projection of LLaMA (hf:models/llama/modeling_llama.py:137-168, 240-262) with RoPE and
// the KV-cache write in the epilogue instead of a separate pass over q / k / v (rope_kvcache_kernel, elementwise.cu).
// `lo` / `hi` are the fp32 accumulators of one row for head dims [d0, d0 + 32) and [d0 + 64, d0 + 96) of one head (d0 =
// 0 or 32): exactly a rotation pair.  Same arithmetic as the separate kernel: the projection is rounded to the 16-bit
// type first (that is what the kernel read back), the rotation is fp32, one rounding at the end.  q goes to D (the
// packed qkv buffer the attention kernel reads), rotated k and v go straight to their cache rows.
template <typename T>
__device__ __forceinline__ void rope_pair_store(const uint32_t (&lo)[32], const uint32_t (&hi)[32],
                                                const GemmKernelParams& p, int row, int n0) {
  const RopeFuse& r = p.rope;
  const int hidden = r.heads * 128;
  const int region = n0 / hidden;                 // 0 = q, 1 = k, 2 = v
  const int col = n0 - region * hidden;
  const int head = col >> 7, d0 = col & 127;
  const int b = row / r.seq, t = row - b * r.seq, pos = r.pos0 + t;
  T* dst;
  if (region == 0) {
    dst = reinterpret_cast<T*>(p.D) + static_cast<int64_t>(row) * p.ldd + n0;
  } else {
    dst = reinterpret_cast<T*>(region == 1 ? r.k_cache : r.v_cache) + b * r.cache_bs + head * r.cache_hs +
          static_cast<int64_t>(pos) * 128 + d0;
  }
  if (region == 2) {
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      uint4 w0, w1;
      w0.x = pack2<T>(__uint_as_float(lo[k]), __uint_as_float(lo[k + 1]));
      w0.y = pack2<T>(__uint_as_float(lo[k + 2]), __uint_as_float(lo[k + 3]));
      w0.z = pack2<T>(__uint_as_float(lo[k + 4]), __uint_as_float(lo[k + 5]));
      w0.w = pack2<T>(__uint_as_float(lo[k + 6]), __uint_as_float(lo[k + 7]));
      w1.x = pack2<T>(__uint_as_float(hi[k]), __uint_as_float(hi[k + 1]));
      w1.y = pack2<T>(__uint_as_float(hi[k + 2]), __uint_as_float(hi[k + 3]));
      w1.z = pack2<T>(__uint_as_float(hi[k + 4]), __uint_as_float(hi[k + 5]));
      w1.w = pack2<T>(__uint_as_float(hi[k + 6]), __uint_as_float(hi[k + 7]));
      *reinterpret_cast<uint4*>(dst + k) = w0;
      *reinterpret_cast<uint4*>(dst + 64 + k) = w1;
    }
    return;
  }
  const float* cs = r.cos_t + static_cast<int64_t>(pos) * 64 + d0;
  const float* sn = r.sin_t + static_cast<int64_t>(pos) * 64 + d0;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const float4 c0 = *reinterpret_cast<const float4*>(cs + k), c1 = *reinterpret_cast<const float4*>(cs + k + 4);
    const float4 s0 = *reinterpret_cast<const float4*>(sn + k), s1 = *reinterpret_cast<const float4*>(sn + k + 4);
    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    uint32_t ol[4], ou[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 a = unpack2<T>(pack2<T>(__uint_as_float(lo[k + 2 * e]), __uint_as_float(lo[k + 2 * e + 1])));
      const float2 c = unpack2<T>(pack2<T>(__uint_as_float(hi[k + 2 * e]), __uint_as_float(hi[k + 2 * e + 1])));
      ol[e] = pack2<T>(a.x * cc[2 * e] - c.x * ss[2 * e], a.y * cc[2 * e + 1] - c.y * ss[2 * e + 1]);
      ou[e] = pack2<T>(c.x * cc[2 * e] + a.x * ss[2 * e], c.y * cc[2 * e + 1] + a.y * ss[2 * e + 1]);
    }
    *reinterpret_cast<uint4*>(dst + k) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    *reinterpret_cast<uint4*>(dst + 64 + k) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
  }
}

// Epilogue of one 32-row x 32-column chunk through shared memory and a TMA store (16-bit D, no SiLU-mul): bias,
// activation and residual in registers like epilogue_store, then each lane writes its row's 64 bytes into a
// SWIZZLE_64B staging tile (16-byte chunk k of row r at chunk k ^ ((r >> 1) & 3): conflict-free), and one lane issues
// cp.async.bulk.tensor for the whole tile -- full 64-byte row segments instead of four 16-byte stores per lane, and
// rows / columns beyond M / N are clipped by the tensor map.  The whole warp calls this.
template <typename T, int EPI>
__device__ __forceinline__ void epilogue_tma_store(float (&v)[32], const GemmKernelParams& p, const T* bias,
                                                   const T* resid, int row, bool row_ok, int row0, int n0,
                                                   uint8_t* stage, const CUtensorMap* tmD, int lane) {
  const bool full_chunk = (n0 + 32 <= p.N);
  if (bias != nullptr) {
    if (full_chunk && ((reinterpret_cast<uintptr_t>(bias + n0) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 b = *reinterpret_cast<const uint4*>(bias + n0 + i);
        float2 f;
        f = unpack2<T>(b.x); v[i] += f.x; v[i + 1] += f.y;
        f = unpack2<T>(b.y); v[i + 2] += f.x; v[i + 3] += f.y;
        f = unpack2<T>(b.z); v[i + 4] += f.x; v[i + 5] += f.y;
        f = unpack2<T>(b.w); v[i + 6] += f.x; v[i + 7] += f.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N) v[i] += T16<T>::to_f(bias[n0 + i]);
    }
  }
  if constexpr (EPI != EPI_NONE) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = apply_act<EPI>(v[i]);
  }
  if (resid != nullptr && row_ok) {
    const T* rr = resid + static_cast<int64_t>(row) * p.ldr + n0;
    if (full_chunk && ((reinterpret_cast<uintptr_t>(rr) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 b = *reinterpret_cast<const uint4*>(rr + i);
        float2 f;
        f = unpack2<T>(b.x); v[i] += f.x; v[i + 1] += f.y;
        f = unpack2<T>(b.y); v[i + 2] += f.x; v[i + 3] += f.y;
        f = unpack2<T>(b.z); v[i + 4] += f.x; v[i + 5] += f.y;
        f = unpack2<T>(b.w); v[i + 6] += f.x; v[i + 7] += f.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N) v[i] += T16<T>::to_f(rr[i]);
    }
  }
  uint8_t* my = stage + lane * 64;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint4 w;
    w.x = pack2<T>(v[8 * k], v[8 * k + 1]);
    w.y = pack2<T>(v[8 * k + 2], v[8 * k + 3]);
    w.z = pack2<T>(v[8 * k + 4], v[8 * k + 5]);
    w.w = pack2<T>(v[8 * k + 6], v[8 * k + 7]);
    *reinterpret_cast<uint4*>(my + ((k ^ sw) << 4)) = w;
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_4d(tmD, stage, n0, row0, 0, 0);
    tma_store_commit();
  }
}

template <typename T, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmD, const GemmKernelParams p) {
  using S = PairSmem;
  constexpr int STAGES = kPairStages;
  constexpr int BN = kPairBN;
  constexpr uint32_t kIdesc = make_idesc_f16(T16<T>::kUmmaFormat, kPairBM, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* store_stage = smem + S::kStoreOffset;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = p.num_m * p.num_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);    // leader's copy is the live one: one arrive.expect_tx for both CTAs' bytes
      mbar_init(&empty_bar[s], 1);   // one multicast tcgen05.commit per use, in each CTA
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * kEpiWarps);  // leader's copy: one elected arrive per epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<2>(tmem_ptr, 512);
  tc_fence_before();
  cluster_sync_all();   // both CTAs' barriers and TMEM exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own A rows, own half of B) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        int split, m_blk, n_blk;
        tile_coords(t, p.num_m, p.num_n, p.group_m, split, m_blk, n_blk);
        const int row_a = m_blk * kPairBM + static_cast<int>(rank) * BM;
        const int row_b = n_blk * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
          tma_load_2d_2sm_hint(sa, &tmA, &full_bar[stage], kb * BK, row_a, p.hint_a);
          tma_load_2d_2sm_hint(sa + S::kABytes, &tmB, &full_bar[stage], kb * BK, row_b, p.hint_b);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (rank == 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint64_t a_desc = make_kmajor_sw128_desc(sa);
          const uint64_t b_desc = make_kmajor_sw128_desc(sa + S::kABytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16<2>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit<2>(&empty_bar[stage]);   // frees the slot in BOTH CTAs
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit<2>(&tmem_full[acc]);       // accumulators of both CTAs complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue (both CTAs: own 128 rows x 256 columns) =====================
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const T* bias = reinterpret_cast<const T*>(p.bias);
    const T* resid = reinterpret_cast<const T*>(p.residual);
    for (int t = pair; t < num_tiles; t += num_pairs) {
      int split, m_blk, n_blk;
      tile_coords(t, p.num_m, p.num_n, p.group_m, split, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * kPairBM + static_cast<int>(rank) * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
      // This warp's chunks c = half, half + 2, ...: the TMEM load of the next chunk is in flight while the current one
      // goes through the activation and the stores, and the accumulator is handed back to the MMA thread as soon as
      // the LAST load has landed in registers -- not after the last store (with K = 1024 / 1280 a tile's main loop
      // is barely longer than its epilogue, so every cycle the accumulator is held stalls the tensor pipe).
      // Same-box A/B against load -> wait -> activation -> stores per chunk (tools/bench_gemm.py SHAPES=b32, and the
      // whole step): within +-2 % per shape either way, i.e. neutral at the power cap -- kept for the early release.
      if constexpr (EPI == EPI_QKV_ROPE) {
        // chunks (c, c + 2) are the two halves of a rotation pair: this warp takes c = half and c = half + 4
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
          const int c = half + 4 * pr;
          uint32_t lo[32], hi[32];
          tmem_ld_32x32(taddr + c * 32, lo);
          tmem_ld_32x32(taddr + (c + 2) * 32, hi);
          tmem_ld_wait();
          if (pr == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (rank == 0) mbar_arrive(&tmem_empty[acc]);
              else mbar_arrive_cluster(&tmem_empty[acc], 0);
            }
          }
          if (row_ok) rope_pair_store<T>(lo, hi, p, row, n_blk * BN + c * 32);
        }
      } else {
      constexpr int kChunks = BN / 64;
      uint32_t rn[32];
      tmem_ld_32x32(taddr + half * 32, rn);
#pragma unroll 1   // one copy of the activation code: unrolled four times the GELU epilogue ran 12 % slower
      for (int i = 0; i < kChunks; ++i) {
        const int c = half + 2 * i;
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(rn[k]);
        if (i + 1 < kChunks) {
          tmem_ld_32x32(taddr + (c + 2) * 32, rn);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&tmem_empty[acc]);
            else mbar_arrive_cluster(&tmem_empty[acc], 0);
          }
        }
        const int n0 = n_blk * BN + c * 32;
        if constexpr (EPI != EPI_SILU_MUL) {
          if (p.tma_store) {
            if (n0 < p.N) {
              // staging buffer i & 1 of this warp: the store issued from it two chunks ago must have read it
              if (lane == 0) tma_store_wait_read<1>();
              __syncwarp();
              epilogue_tma_store<T, EPI>(v, p, bias, resid, row, row_ok, row - lane, n0,
                                         store_stage + ((warp - kEpiWarp0) * 2 + (i & 1)) * 2048, &tmD, lane);
            }
            continue;
          }
        }
        if (row_ok && n0 < p.N) epilogue_store<T, EPI, 32>(v, p, bias, resid, false, 0, row, n0);
      }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait<0>();   // this thread's bulk stores have completed before the CTA may exit
  }

  tc_fence_before();
  cluster_sync_all();   // nobody leaves while the peer may still read its shared memory or signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

// -----------------------------------------------------------------------------
// Split-K reduce + epilogue.  partial: fp32 [splits][R][C].
//   transpose == 0: out[r][c]   (R = M rows, C = N cols)
//   transpose == 1: out[c][r]   (swap-AB: R = weight rows (N_out), C = padded batch; out is [batch][N_out])
// bias is indexed by the output column, residual has the output's layout.
// One block handles a 32 (r) x 32 (c) patch.
// -----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, int R, int C, int c_valid, void* __restrict__ out,
                     int64_t ldo, const T* __restrict__ bias, const T* __restrict__ resid, int64_t ldr, int epilogue,
                     int out_f32, int transpose) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31;
  const int ty = threadIdx.x >> 5;  // 0..7
  for (int rr = ty; rr < 32; rr += 8) {
    const int r = r0 + rr, c = c0 + tx;
    float s = 0.f;
    if (r < R && c < C) {
      const float* ptr = partial + static_cast<int64_t>(r) * C + c;
      for (int k = 0; k < splits; ++k) s += ptr[static_cast<int64_t>(k) * R * C];
    }
    tile[rr][tx] = s;
  }
  __syncthreads();
  if (!transpose) {
    // out[r][c]; threads along c
    for (int rr = ty; rr < 32; rr += 8) {
      const int r = r0 + rr;
      if (r >= R) continue;
      if (epilogue == EPI_SILU_MUL) {
        // chunk of 32 columns = 16 gate + 16 up -> 16 outputs
        if (tx < 16) {
          const int c = c0 + tx;
          if (c < c_valid) {
            const float g = tile[rr][tx], u = tile[rr][tx + 16];
            float o = __fdividef(g, 1.f + __expf(fminf(-g, 80.f))) * u;
            const int oc = (c0 >> 1) + tx;
            if (resid) o += T16<T>::to_f(resid[static_cast<int64_t>(r) * ldr + oc]);
            if (out_f32) reinterpret_cast<float*>(out)[static_cast<int64_t>(r) * ldo + oc] = o;
            else reinterpret_cast<T*>(out)[static_cast<int64_t>(r) * ldo + oc] = T16<T>::from_f(o);
          }
        }
        continue;
      }
      const int c = c0 + tx;
      if (c >= c_valid) continue;
      float v = tile[rr][tx];
      if (bias) v += T16<T>::to_f(bias[c]);
      v = apply_act_rt(v, epilogue);
      if (resid) v += T16<T>::to_f(resid[static_cast<int64_t>(r) * ldr + c]);
      if (out_f32) reinterpret_cast<float*>(out)[static_cast<int64_t>(r) * ldo + c] = v;
      else reinterpret_cast<T*>(out)[static_cast<int64_t>(r) * ldo + c] = T16<T>::from_f(v);
    }
  } else {
    // out[c][r]; threads along r (the contiguous output dimension)
    for (int cc = ty; cc < 32; cc += 8) {
      const int c = c0 + cc;
      if (c >= c_valid) continue;
      if (epilogue == EPI_SILU_MUL) {
        if (tx < 16) {
          const int r = r0 + tx;  // gate row; up row = r + 16 (same 32-row chunk)
          if (r < R) {
            const float g = tile[tx][cc], u = tile[tx + 16][cc];
            float o = __fdividef(g, 1.f + __expf(fminf(-g, 80.f))) * u;
            const int orow = (r0 >> 1) + tx;
            if (resid) o += T16<T>::to_f(resid[static_cast<int64_t>(c) * ldr + orow]);
            if (out_f32) reinterpret_cast<float*>(out)[static_cast<int64_t>(c) * ldo + orow] = o;
            else reinterpret_cast<T*>(out)[static_cast<int64_t>(c) * ldo + orow] = T16<T>::from_f(o);
          }
        }
        continue;
      }
      const int r = r0 + tx;
      if (r >= R) continue;
      float v = tile[tx][cc];
      if (bias) v += T16<T>::to_f(bias[r]);
      v = apply_act_rt(v, epilogue);
      if (resid) v += T16<T>::to_f(resid[static_cast<int64_t>(c) * ldr + r]);
      if (out_f32) reinterpret_cast<float*>(out)[static_cast<int64_t>(c) * ldo + r] = v;
      else reinterpret_cast<T*>(out)[static_cast<int64_t>(c) * ldo + r] = T16<T>::from_f(v);
    }
  }
}

// -----------------------------------------------------------------------------
// Host side
// -----------------------------------------------------------------------------
template <typename T, int BN, int STAGES, int EPI>
static int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKernelParams& p, int grid,
                          cudaStream_t stream) {
  using S = GemmSmem<BN, STAGES>;
  auto kern = gemm_tcgen05_kernel<T, BN, STAGES, EPI>;
  static SmemOptIn opt_in;   // per device (common.cuh)
  { const int _st = ensure_dynamic_smem(kern, S::kTotal, opt_in); if (_st != OK) return _st; }
  kern<<<grid, kGemmThreads, S::kTotal, stream>>>(ta, tb, p);
  return check_cuda(cudaGetLastError(), "gemm_tcgen05_kernel launch");
}

template <typename T, int EPI>
static int launch_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmKernelParams& p, int grid,
                     cudaStream_t stream) {
  switch (bn) {
    case 256: return launch_variant<T, 256, 4, EPI>(ta, tb, p, grid, stream);
    case 128: return launch_variant<T, 128, 6, EPI>(ta, tb, p, grid, stream);
    case 64: return launch_variant<T, 64, 8, EPI>(ta, tb, p, grid, stream);
    case 32: return launch_variant<T, 32, 8, EPI>(ta, tb, p, grid, stream);
    case 16: return launch_variant<T, 16, 8, EPI>(ta, tb, p, grid, stream);
    default: set_last_error("gemm: unsupported BN %d", bn); return ERR_UNSUPPORTED;
  }
}

template <typename T, int EPI>
static int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const GemmKernelParams& p,
                       int grid, cudaStream_t stream) {
  auto kern = gemm_tcgen05_pair_kernel<T, EPI>;
  static SmemOptIn opt_in;   // per device (common.cuh)
  { const int _st = ensure_dynamic_smem(kern, PairSmem::kTotal, opt_in); if (_st != OK) return _st; }
  kern<<<grid, kGemmThreads, PairSmem::kTotal, stream>>>(ta, tb, td, p);   // __cluster_dims__(2, 1, 1): grid is even
  return check_cuda(cudaGetLastError(), "gemm_tcgen05_pair_kernel launch");
}

template <typename T>
static int launch_pair_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td,
                           const GemmKernelParams& p, int grid, cudaStream_t stream) {
  switch (epi) {
    case EPI_NONE: return launch_pair<T, EPI_NONE>(ta, tb, td, p, grid, stream);
    case EPI_RELU: return launch_pair<T, EPI_RELU>(ta, tb, td, p, grid, stream);
    case EPI_GELU: return launch_pair<T, EPI_GELU>(ta, tb, td, p, grid, stream);
    case EPI_QUICK_GELU: return launch_pair<T, EPI_QUICK_GELU>(ta, tb, td, p, grid, stream);
    case EPI_SILU_MUL: return launch_pair<T, EPI_SILU_MUL>(ta, tb, td, p, grid, stream);
    case EPI_QKV_ROPE: return launch_pair<T, EPI_QKV_ROPE>(ta, tb, td, p, grid, stream);
    default: set_last_error("gemm: unknown epilogue %d", epi); return ERR_BAD_ARG;
  }
}

// the activation is compiled into the kernel; with split-K the kernel only parks fp32 partials (EPI_NONE)
template <typename T>
static int launch_epi(int epi, int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmKernelParams& p, int grid,
                      cudaStream_t stream) {
  switch (epi) {
    case EPI_NONE: return launch_bn<T, EPI_NONE>(bn, ta, tb, p, grid, stream);
    case EPI_RELU: return launch_bn<T, EPI_RELU>(bn, ta, tb, p, grid, stream);
    case EPI_GELU: return launch_bn<T, EPI_GELU>(bn, ta, tb, p, grid, stream);
    case EPI_QUICK_GELU: return launch_bn<T, EPI_QUICK_GELU>(bn, ta, tb, p, grid, stream);
    case EPI_SILU_MUL: return launch_bn<T, EPI_SILU_MUL>(bn, ta, tb, p, grid, stream);
    default: set_last_error("gemm: unknown epilogue %d", epi); return ERR_BAD_ARG;
  }
}

static int pick_bn_large(int N) {
  if (N > 128) return 256;
  if (N > 64) return 128;
  if (N > 32) return 64;
  if (N > 16) return 32;
  return 16;
}

// Large products run on CTA pairs (256 x 256 tile per two SMs) when there are enough pair tiles to fill the machine.
bool gemm_pair_eligible(const Context* ctx, int M, int N) {
  const int sms = ctx->sm_count;
  return ctx->gemm_pair != 0 && M > 32 && N >= 256 &&
         static_cast<long long>((M + kPairBM - 1) / kPairBM) * ((N + kPairBN - 1) / kPairBN) >= 2ll * (sms / 2);
}

int gemm_run(Context* ctx, const GemmArgs& a, cudaStream_t stream) {
  const RopeFuse* rope = ctx->rope_fuse;   // one-shot: whatever happens below, it never outlives this call
  ctx->rope_fuse = nullptr;
  ULLAVA_REQUIRE(a.A && a.B && a.D, "gemm: null operand");
  ULLAVA_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  ULLAVA_REQUIRE(a.dtype == DT_BF16 || a.dtype == DT_F16, "gemm: dtype must be bf16/f16");
  ULLAVA_REQUIRE((a.lda % 8) == 0 && (a.ldb % 8) == 0, "gemm: lda/ldb must be multiples of 8 elements (TMA 16 B stride)");
  ULLAVA_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0,
                 "gemm: A/B must be 16-byte aligned");
  ULLAVA_REQUIRE(a.lda >= a.K && a.ldb >= a.K, "gemm: leading dimension smaller than K");
  if (a.epilogue == EPI_SILU_MUL) ULLAVA_REQUIRE((a.N % 32) == 0, "gemm: SILU_MUL needs N %% 32 == 0");

  const int kb_total = (a.K + BK - 1) / BK;
  const int sms = ctx->sm_count;

  // Small-M products: swap operands so the weight matrix feeds the 128-row side.
  const bool swap = (a.M <= 32) && (a.N >= 128) && !a.no_swap;
  const double out_cols = a.epilogue == EPI_SILU_MUL ? a.N / 2.0 : a.N;
  ProfScope _ps(ctx, stream, swap ? ULLAVA_PROF_GEMM_STREAM : ULLAVA_PROF_GEMM_TENSOR, 2.0 * a.M * a.N * a.K,
                2.0 * (static_cast<double>(a.M) * a.K + static_cast<double>(a.N) * a.K) +
                    (a.out_f32 ? 4.0 : 2.0) * a.M * out_cols);
  GemmKernelParams p{};
  p.kb_total = kb_total;
  // Rasterisation group.  While a group of M tiles sweeps over N its A panel stays in L2 and B streams through, so B
  // is read from DRAM num_m / group_m times: when B is too large to live in L2 beside the panel (LLaMA qkv / gate-up
  // / down at prefill: 100 / 180 / 90 MB) a taller group cuts those re-reads (3.6 -> ~1 GB on gate/up, +6-7 % on
  // tools/bench_gemm.py), as long as the panel itself (group_m x 128 x K x 2 B) stays well inside L2.
  p.group_m = kDefaultGroupM;
  if (2.0 * a.N * a.K >= 48e6) {
    int g = 32;
    while (g > kDefaultGroupM && 256.0 * g * a.K > 48e6) g /= 2;
    p.group_m = g;
  }
  if (ctx->group_m > 0) p.group_m = ctx->group_m;
  // L2 eviction priorities (CTA-pair kernel).  Whichever operand is re-used ACROSS waves should survive the stream of
  // the other one: a B that fits in L2 beside a wave's working set (<= ~96 MB) while A is the big streamed operand
  // (LLaMA down: A 428 MB, B 90 MB) is loaded evict_last / A evict_first; when B is too large and the raster keeps a
  // tall A panel resident instead (qkv, gate/up), the panel is evict_last and B streams evict_first.
  // MEASURED (ncu dram__bytes_read, tools/gemm_traffic.py, round 2): the hints make it worse -- qkv 1.04 -> 2.68 GB,
  // gate/up 2.0 -> 5.1 GB, down 2.7 -> 4.0 GB read, 3-6 % slower: evict_first also throws out the tiles that the OTHER
  // CTAs of the same wave are about to re-read (the within-wave reuse is what keeps the traffic at 2-4x the operands).
  // Off by default; kept as an experiment switch.
  p.hint_a = p.hint_b = kEvictNormal;
  if (ctx->gemm_hints) {
    const double a_bytes = 2.0 * a.M * a.K, b_bytes = 2.0 * a.N * a.K;
    if (b_bytes <= 96e6 && a_bytes > 1.5 * b_bytes && b_bytes > 16e6) { p.hint_a = kEvictFirst; p.hint_b = kEvictLast; }
    else if (b_bytes >= 48e6) { p.hint_a = kEvictLast; p.hint_b = kEvictFirst; }
  }
  p.epilogue = a.epilogue;
  p.out_f32 = a.out_f32;
  p.bias = a.bias;
  p.residual = a.residual;
  p.ldr = a.ldr;
  CUtensorMap ta, tb;

  if (swap && a.force_splits > 0) {
    set_last_error("gemm: force_splits does not apply to the weight-streaming (M <= 32) path");
    return ERR_BAD_ARG;
  }
  if (swap && rope) {
    set_last_error("gemm: the fused RoPE epilogue does not apply to the weight-streaming (M <= 32) path");
    return ERR_BAD_ARG;
  }
  if (!swap) {
    ctx->next_w = nullptr;  // the next-weight hint only applies to the weight-streaming kernel
    // CTA-pair kernel (256 x 256 tile per two SMs) for the large products: enough pair tiles to fill the machine
    const bool pair_ok = !a.force_bn && a.force_splits <= 1 && gemm_pair_eligible(ctx, a.M, a.N);
    if (rope && (!pair_ok || a.epilogue != EPI_NONE || a.bias || a.residual || a.out_f32 || a.N != 3 * rope->heads * 128)) {
      set_last_error("gemm: the fused RoPE epilogue needs the CTA-pair kernel and a plain [rows, 3 * heads * 128] product");
      return ERR_BAD_ARG;
    }
    if (pair_ok) {
      p.M = a.M; p.N = a.N; p.K = a.K;
      p.num_m = (a.M + kPairBM - 1) / kPairBM;
      p.num_n = (a.N + kPairBN - 1) / kPairBN;
      p.kb_per_split = kb_total;
      p.splits = 1;
      p.D = a.D; p.ldd = a.ldd;
      p.group_m = (p.group_m + 1) / 2;   // groups are counted in 256-row tiles here
      int st = encode_tmap_2d(&ta, a.A, 2, a.K, a.M, a.lda * 2, BK, BM, true);
      if (st) return st;
      st = encode_tmap_2d(&tb, a.B, 2, a.K, a.N, a.ldb * 2, BK, kPairBN / 2, true);
      if (st) return st;
      const int tiles = p.num_m * p.num_n;
      const int grid = 2 * (tiles < sms / 2 ? tiles : sms / 2);
      int epi = a.epilogue;
      if (rope) {
        p.rope = *rope;
        epi = EPI_QKV_ROPE;
      }
      // D through shared memory + TMA store whenever it is a plain 16-bit matrix the tensor map can describe
      CUtensorMap td = ta;
      p.tma_store = 0;
      if (!rope && !a.out_f32 && a.epilogue != EPI_SILU_MUL && ctx->gemm_tma_store != 0 && (a.ldd % 8) == 0 &&
          (reinterpret_cast<uintptr_t>(a.D) & 15) == 0) {
        const uint64_t dd[4] = {static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.M), 1, 1};
        const uint64_t ds[3] = {static_cast<uint64_t>(a.ldd) * 2, static_cast<uint64_t>(a.ldd) * 2 * a.M,
                                static_cast<uint64_t>(a.ldd) * 2 * a.M};
        st = encode_tmap_4d(&td, a.D, dd, ds, 32, 32, 64);
        if (st) return st;
        p.tma_store = 1;
      }
      st = (a.dtype == DT_BF16) ? launch_pair_epi<__nv_bfloat16>(epi, ta, tb, td, p, grid, stream)
                                : launch_pair_epi<__half>(epi, ta, tb, td, p, grid, stream);
      if (st) return st;
      ctx->launches += 1;
      return OK;
    }
    const int bn = a.force_bn ? a.force_bn : pick_bn_large(a.N);
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.num_m = (a.M + BM - 1) / BM;
    p.num_n = (a.N + bn - 1) / bn;
    int splits = a.force_splits > 0 ? a.force_splits : 1;
    if (splits > kb_total) splits = kb_total;
    p.kb_per_split = (kb_total + splits - 1) / splits;
    splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.splits = splits;
    int st = encode_tmap_2d(&ta, a.A, 2, a.K, a.M, a.lda * 2, BK, BM, true);
    if (st) return st;
    st = encode_tmap_2d(&tb, a.B, 2, a.K, a.N, a.ldb * 2, BK, bn, true);
    if (st) return st;
    if (splits > 1) {
      const size_t need = kStreamCounterBytes + static_cast<size_t>(splits) * a.M * a.N * sizeof(float);
      if (need > ctx->workspace_bytes) {
        set_last_error("gemm: split-K workspace too small (%zu > %zu)", need, ctx->workspace_bytes);
        return ERR_WORKSPACE;
      }
      p.D = static_cast<uint8_t*>(ctx->workspace) + kStreamCounterBytes; p.ldd = a.N;
    } else {
      p.D = a.D; p.ldd = a.ldd;
    }
    const int tiles = p.num_m * p.num_n * splits;
    const int grid = tiles < sms ? tiles : sms;
    const int kepi = splits > 1 ? static_cast<int>(EPI_NONE) : a.epilogue;
    st = (a.dtype == DT_BF16) ? launch_epi<__nv_bfloat16>(kepi, bn, ta, tb, p, grid, stream)
                              : launch_epi<__half>(kepi, bn, ta, tb, p, grid, stream);
    if (st) return st;
    if (splits > 1) {
      dim3 g((a.M + 31) / 32, (a.N + 31) / 32);
      if (a.dtype == DT_BF16)
        splitk_reduce_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>(
            reinterpret_cast<const float*>(p.D), splits, a.M, a.N, a.N, a.D, a.ldd,
            reinterpret_cast<const __nv_bfloat16*>(a.bias), reinterpret_cast<const __nv_bfloat16*>(a.residual), a.ldr,
            a.epilogue, a.out_f32, 0);
      else
        splitk_reduce_kernel<__half><<<g, 256, 0, stream>>>(
            reinterpret_cast<const float*>(p.D), splits, a.M, a.N, a.N, a.D, a.ldd,
            reinterpret_cast<const __half*>(a.bias), reinterpret_cast<const __half*>(a.residual), a.ldr, a.epilogue,
            a.out_f32, 0);
      ctx->launches += 2;
      return check_cuda(cudaGetLastError(), "splitk_reduce launch");
    }
    ctx->launches += 1;
    return OK;
  }

  // ---- small-M (decode) products: dedicated weight-streaming kernel (gemm_stream_sm100.cu) ----
  return gemm_stream_run(ctx, a, stream);
}

}  // namespace ullava
