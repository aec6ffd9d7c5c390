// Attention kernels of the u-LLaVA hot path.
//
//   attention_run        : flash-style softmax(Q K^T * scale) V, fp32 online softmax, non-causal
//                          (CLIP ViT, hd 64, S = 577; hf:models/clip/modeling_clip.py:282-336) and
//                          causal (LLaMA prefill, hd 128, S = 608; hf:models/llama/modeling_llama.py:199-291).
//   attention_decode_run : single-query attention against the KV cache (LLaMA decode step).
//
// Attention is ~1-3 % of the path's FLOPs, so round 1 keeps it on the
// warp-level tensor-core path (mma.sync.m16n8k16 + ldmatrix + cp.async double buffering): K/V tiles
// are read from HBM/L2 exactly once per 64-row query block, nothing of size S x S is materialised.
// The tcgen05/TMEM version (S in TMEM, P fed back as the A operand) is the planned follow-up.
#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

// ---- small PTX helpers ------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- flash forward ------------------------------------------------------------------------------
static constexpr int FA_BM = 64;   // query rows per CTA (4 warps x 16)
static constexpr int FA_BN = 64;   // keys per tile
static constexpr int FA_THREADS = 128;

struct FlashParams {
  const void* q; int64_t q_bs, q_rs, q_hs;
  const void* k; int64_t k_bs, k_rs, k_hs;
  const void* v; int64_t v_bs, v_rs, v_hs;
  void* o; int64_t o_bs, o_rs, o_hs;
  int seq_q, seq_k, causal, q_pos0;
  float scale_log2;  // scale * log2(e)
};

// shared tile [rows][HD] 16-bit with a 16-byte-chunk XOR swizzle (conflict-free ldmatrix).
// HD = 80 (SAM ViT-H heads: 10 chunks per row, not a power of two) uses rows padded to 176 bytes instead:
// 8 consecutive rows then start at banks {0,12,24,4,16,28,8,20}, again conflict-free for ldmatrix.
template <int HD>
__host__ __device__ constexpr int fa_row_bytes() { return HD == 80 ? 176 : HD * 2; }
template <int HD>
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) {
  if constexpr (HD == 80) {
    return static_cast<uint32_t>(row * 176 + (chunk << 4));
  } else {
    constexpr int CPR = HD / 8;                    // 16-byte chunks per row
    constexpr int MASK = CPR >= 8 ? 7 : CPR - 1;   // XOR only within the row's own chunks
    return static_cast<uint32_t>(row * (HD * 2) + (((chunk & ~MASK) | ((chunk ^ row) & MASK)) << 4));
  }
}

template <typename T, int HD>
__device__ __forceinline__ void load_tile(uint32_t smem_base, const T* g, int64_t row_stride, int row0, int nrows_valid,
                                          int tid) {
  constexpr int CPR = HD / 8;  // 16-byte chunks per row
  for (int i = tid; i < FA_BN * CPR; i += FA_THREADS) {
    const int r = i / CPR, c = i - r * CPR;
    const bool ok = (row0 + r) < nrows_valid;
    const T* src = g + static_cast<int64_t>(ok ? (row0 + r) : 0) * row_stride + c * 8;
    cp_async16(smem_base + sw_off<HD>(r, c), src, ok);
  }
}

template <typename T, int HD>
__global__ void __launch_bounds__(FA_THREADS)
flash_fwd_kernel(const FlashParams p) {
  extern __shared__ __align__(128) uint8_t fa_smem[];
  constexpr int TILE_BYTES = FA_BN * fa_row_bytes<HD>();
  const uint32_t sQ = smem_u32(fa_smem);
  const uint32_t sK = sQ + TILE_BYTES;            // 2 buffers
  const uint32_t sV = sK + 2 * TILE_BYTES;        // 2 buffers

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * FA_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const T* qg = static_cast<const T*>(p.q) + b * p.q_bs + h * p.q_hs;
  const T* kg = static_cast<const T*>(p.k) + b * p.k_bs + h * p.k_hs;
  const T* vg = static_cast<const T*>(p.v) + b * p.v_bs + h * p.v_hs;
  T* og = static_cast<T*>(p.o) + b * p.o_bs + h * p.o_hs;

  int k_end = p.seq_k;
  if (p.causal) k_end = min(k_end, p.q_pos0 + m0 + FA_BM);
  const int n_tiles = (k_end + FA_BN - 1) / FA_BN;

  // prologue: Q tile + first K/V tile
  load_tile<T, HD>(sQ, qg, p.q_rs, m0, p.seq_q, tid);
  load_tile<T, HD>(sK, kg, p.k_rs, 0, p.seq_k, tid);
  load_tile<T, HD>(sV, vg, p.v_rs, 0, p.seq_k, tid);
  cp_async_commit();

  constexpr int KS = HD / 16;  // k-steps of QK^T
  uint32_t qf[KS][4];
  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  const int g = lane >> 2, tq = lane & 3;
  const int qrow0 = m0 + warp * 16 + g;  // this thread's rows: qrow0 and qrow0 + 8

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      load_tile<T, HD>(sK + (buf ^ 1) * TILE_BYTES, kg, p.k_rs, (t + 1) * FA_BN, p.seq_k, tid);
      load_tile<T, HD>(sV + (buf ^ 1) * TILE_BYTES, vg, p.v_rs, (t + 1) * FA_BN, p.seq_k, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (t == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = ks * 2 + (lane >> 4);
        ldsm_x4(sQ + sw_off<HD>(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[FA_BN / 8][4];
#pragma unroll
    for (int i = 0; i < FA_BN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    const uint32_t kb = sK + buf * TILE_BYTES;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int nb = 0; nb < FA_BN / 16; ++nb) {
        uint32_t r0, r1, r2, r3;
        const int row = nb * 16 + (lane & 7) + (lane >> 4) * 8;
        const int chunk = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(kb + sw_off<HD>(row, chunk), r0, r1, r2, r3);
        mma16816<T>(s[nb * 2], qf[ks], r0, r1);
        mma16816<T>(s[nb * 2 + 1], qf[ks], r2, r3);
      }
    }

    // ---- mask + online softmax ----
    const int key0 = t * FA_BN;
    const bool need_mask = (key0 + FA_BN > p.seq_k) || (p.causal && (key0 + FA_BN - 1 > p.q_pos0 + m0 + warp * 16));
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < FA_BN / 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float val = s[nb][e] * p.scale_log2;
        if (need_mask) {
          const int key = key0 + nb * 8 + tq * 2 + (e & 1);
          const int qr = qrow0 + (e >> 1) * 8;
          if (key >= p.seq_k || (p.causal && key > p.q_pos0 + qr)) val = -INFINITY;
        }
        s[nb][e] = val;
        mx[e >> 1] = fmaxf(mx[e >> 1], val);
      }
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mnew[r] = fmaxf(m_run[r], mx[r]);
      // fully masked so far: keep everything at zero without producing NaN from (-inf) - (-inf)
      const float msafe = (mnew[r] == -INFINITY) ? 0.f : mnew[r];
      corr[r] = exp2f(m_run[r] - msafe);
      m_run[r] = mnew[r];
      mnew[r] = msafe;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[FA_BN / 16][4];
#pragma unroll
    for (int nb = 0; nb < FA_BN / 8; ++nb) {
      const float p0 = exp2f(s[nb][0] - mnew[0]);
      const float p1 = exp2f(s[nb][1] - mnew[0]);
      const float p2 = exp2f(s[nb][2] - mnew[1]);
      const float p3 = exp2f(s[nb][3] - mnew[1]);
      // round P to the 16-bit dtype first so that the row sum matches what the PV MMA consumes
      const uint32_t lo = pack2<T>(p0, p1), hi = pack2<T>(p2, p3);
      const float2 flo = unpack2<T>(lo), fhi = unpack2<T>(hi);
      rs[0] += flo.x + flo.y;
      rs[1] += fhi.x + fhi.y;
      pf[nb >> 1][(nb & 1) * 2 + 0] = lo;
      pf[nb >> 1][(nb & 1) * 2 + 1] = hi;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o_acc[i][0] *= corr[0];
      o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1];
      o_acc[i][3] *= corr[1];
    }

    // ---- O += P V ----
    const uint32_t vb = sV + buf * TILE_BYTES;
#pragma unroll
    for (int kk = 0; kk < FA_BN / 16; ++kk) {
#pragma unroll
      for (int db = 0; db < HD / 16; ++db) {
        uint32_t r0, r1, r2, r3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = db * 2 + (lane >> 4);
        ldsm_x4_t(vb + sw_off<HD>(row, chunk), r0, r1, r2, r3);
        mma16816<T>(o_acc[db * 2], pf[kk], r0, r1);
        mma16816<T>(o_acc[db * 2 + 1], pf[kk], r2, r3);
      }
    }
    __syncthreads();  // everyone done with buffer `buf` before it is refilled (iteration t+1 prefetches into it)
  }

  // ---- finalise: O / l ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    const int col = i * 8 + tq * 2;
    if (qrow0 < p.seq_q)
      *reinterpret_cast<uint32_t*>(og + static_cast<int64_t>(qrow0) * p.o_rs + col) =
          pack2<T>(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
    if (qrow0 + 8 < p.seq_q)
      *reinterpret_cast<uint32_t*>(og + static_cast<int64_t>(qrow0 + 8) * p.o_rs + col) =
          pack2<T>(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
  }
}

template <typename T, int HD>
static int flash_launch(const FlashParams& p, int batch, int heads, cudaStream_t stream) {
  constexpr int smem = 5 * FA_BN * HD * 2;
  auto kern = flash_fwd_kernel<T, HD>;
  static SmemOptIn opt_in;   // per device (common.cuh)
  { const int _st = ensure_dynamic_smem(kern, smem, opt_in); if (_st != OK) return _st; }
  dim3 grid((p.seq_q + FA_BM - 1) / FA_BM, heads, batch);
  kern<<<grid, FA_THREADS, smem, stream>>>(p);
  return check_cuda(cudaGetLastError(), "flash_fwd launch");
}

int attention_run(Context* ctx, const AttnArgs& a, cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_ATTN_PREFILL, (a.causal ? 2.0 : 4.0) * a.batch * a.heads * (double)a.seq_q * a.seq_k * a.head_dim, 2.0 * a.batch * a.heads * a.head_dim * (2.0 * a.seq_q + 2.0 * a.seq_k));
  ULLAVA_REQUIRE(a.q && a.k && a.v && a.o, "attention: null pointer");
  ULLAVA_REQUIRE(a.batch >= 0 && a.heads > 0 && a.seq_q >= 0 && a.seq_k > 0, "attention: bad shape");
  const int64_t strides[] = {a.q_bs, a.q_rs, a.q_hs, a.k_bs, a.k_rs, a.k_hs, a.v_bs, a.v_rs, a.v_hs, a.o_bs, a.o_rs, a.o_hs};
  for (int64_t s : strides) ULLAVA_REQUIRE(s % 8 == 0, "attention: strides must be multiples of 8 elements");
  ULLAVA_REQUIRE(((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) |
                   reinterpret_cast<uintptr_t>(a.v) | reinterpret_cast<uintptr_t>(a.o)) & 15) == 0,
                 "attention: pointers must be 16-byte aligned");
  if (a.batch == 0 || a.seq_q == 0) return OK;
  // large tiles go to the tcgen05/TMEM kernel; tiny problems (a handful of query rows or keys, e.g. the SAM mask
  // decoder's 7 prompt tokens) would leave a 128 x 128 tile almost empty and stay on the warp-level kernel
  if (ctx->attn_impl != 1 && fmha_supported(a, false, 0) && (ctx->attn_impl == 2 || (a.seq_q >= 32 && a.seq_k >= 32)))
    return fmha_run(ctx, a, nullptr, nullptr, 0, nullptr, stream);
  FlashParams p;
  p.q = a.q; p.q_bs = a.q_bs; p.q_rs = a.q_rs; p.q_hs = a.q_hs;
  p.k = a.k; p.k_bs = a.k_bs; p.k_rs = a.k_rs; p.k_hs = a.k_hs;
  p.v = a.v; p.v_bs = a.v_bs; p.v_rs = a.v_rs; p.v_hs = a.v_hs;
  p.o = a.o; p.o_bs = a.o_bs; p.o_rs = a.o_rs; p.o_hs = a.o_hs;
  p.seq_q = a.seq_q; p.seq_k = a.seq_k; p.causal = a.causal; p.q_pos0 = a.q_pos0;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  int st;
  if (a.dtype == DT_BF16) {
    if (a.head_dim == 64) st = flash_launch<__nv_bfloat16, 64>(p, a.batch, a.heads, stream);
    else if (a.head_dim == 128) st = flash_launch<__nv_bfloat16, 128>(p, a.batch, a.heads, stream);
    else if (a.head_dim == 32) st = flash_launch<__nv_bfloat16, 32>(p, a.batch, a.heads, stream);
    else if (a.head_dim == 16) st = flash_launch<__nv_bfloat16, 16>(p, a.batch, a.heads, stream);
    else { set_last_error("attention: head_dim %d not compiled (16/32/64/128)", a.head_dim); return ERR_UNSUPPORTED; }
  } else if (a.dtype == DT_F16) {
    if (a.head_dim == 64) st = flash_launch<__half, 64>(p, a.batch, a.heads, stream);
    else if (a.head_dim == 128) st = flash_launch<__half, 128>(p, a.batch, a.heads, stream);
    else if (a.head_dim == 32) st = flash_launch<__half, 32>(p, a.batch, a.heads, stream);
    else if (a.head_dim == 16) st = flash_launch<__half, 16>(p, a.batch, a.heads, stream);
    else { set_last_error("attention: head_dim %d not compiled (16/32/64/128)", a.head_dim); return ERR_UNSUPPORTED; }
  } else {
    set_last_error("attention: unsupported dtype %d", a.dtype);
    return ERR_UNSUPPORTED;
  }
  if (st == OK) ctx->launches++;
  return st;
}

// ---- flash forward with SAM's decomposed relative-position bias ----------------------------------
// Attention of the SAM ViT image encoder (segment_anything/modeling/image_encoder.py:196-260, bias :355-392):
//   softmax( scale * q k^T + q . Rh[qh - kh + S - 1] + q . Rw[qw - kw + S - 1] ) v
// over an S x S token grid (S = 14 inside a window, 64 for the global blocks), non causal.
// The two bias products are themselves Q x table^T contractions, so the prologue runs the tables through
// the same mma.sync path as the keys and parks the [64 x (2S-1)] results of this CTA's query rows in
// shared memory; the main loop gathers bias(q, k) = Ph[q][qh-kh+S-1] + Pw[q][qw-kw+S-1] from there.
// o_row_map (optional) scatters output rows: window-layout query row -> un-partitioned token row (-1 = pad).
struct RelPosParams {
  const void* rel_h;        // [2S-1, HD] 16-bit
  const void* rel_w;        // [2S-1, HD] 16-bit
  int S;
  const int32_t* o_row_map; // [batch * seq_q] or nullptr
};

template <typename T, int HD>
__global__ void __launch_bounds__(FA_THREADS)
flash_relpos_kernel(const FlashParams p, const RelPosParams rp) {
  extern __shared__ __align__(128) uint8_t fa_smem[];
  constexpr int TILE_BYTES = FA_BN * fa_row_bytes<HD>();
  const uint32_t sQ = smem_u32(fa_smem);
  const uint32_t sK = sQ + TILE_BYTES;            // 2 buffers
  const uint32_t sV = sK + 2 * TILE_BYTES;        // 2 buffers
  float* sP = reinterpret_cast<float*>(fa_smem + 5 * TILE_BYTES);
  const int S = rp.S, n_rel = 2 * S - 1;
  const int pstride = 2 * n_rel + 1;              // odd: rows land on different banks

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * FA_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const T* qg = static_cast<const T*>(p.q) + b * p.q_bs + h * p.q_hs;
  const T* kg = static_cast<const T*>(p.k) + b * p.k_bs + h * p.k_hs;
  const T* vg = static_cast<const T*>(p.v) + b * p.v_bs + h * p.v_hs;
  const int n_tiles = (p.seq_k + FA_BN - 1) / FA_BN;
  const int g = lane >> 2, tq = lane & 3;
  constexpr int KS = HD / 16;

  // ---- Q tile -> registers ----
  load_tile<T, HD>(sQ, qg, p.q_rs, m0, p.seq_q, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int chunk = ks * 2 + (lane >> 4);
    ldsm_x4(sQ + sw_off<HD>(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }

  // ---- bias prologue: P = Q x table^T for both tables ----
  for (int tb = 0; tb < 2; ++tb) {
    const T* table = static_cast<const T*>(tb == 0 ? rp.rel_h : rp.rel_w);
    for (int c0 = 0; c0 < n_rel; c0 += FA_BN) {
      __syncthreads();  // previous contents of sK[0] consumed
      load_tile<T, HD>(sK, table, HD, c0, n_rel, tid);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      float s[FA_BN / 8][4];
#pragma unroll
      for (int i = 0; i < FA_BN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int nb = 0; nb < FA_BN / 16; ++nb) {
          uint32_t r0, r1, r2, r3;
          const int row = nb * 16 + (lane & 7) + (lane >> 4) * 8;
          const int chunk = ks * 2 + ((lane >> 3) & 1);
          ldsm_x4(sK + sw_off<HD>(row, chunk), r0, r1, r2, r3);
          mma16816<T>(s[nb * 2], qf[ks], r0, r1);
          mma16816<T>(s[nb * 2 + 1], qf[ks], r2, r3);
        }
      }
#pragma unroll
      for (int nb = 0; nb < FA_BN / 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = c0 + nb * 8 + tq * 2 + (e & 1);
          const int ql = warp * 16 + g + (e >> 1) * 8;
          if (col < n_rel) sP[ql * pstride + tb * n_rel + col] = s[nb][e];
        }
      }
    }
  }
  __syncthreads();

  // ---- main loop ----
  load_tile<T, HD>(sK, kg, p.k_rs, 0, p.seq_k, tid);
  load_tile<T, HD>(sV, vg, p.v_rs, 0, p.seq_k, tid);
  cp_async_commit();

  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const int qrow0 = m0 + warp * 16 + g;  // this thread's rows: qrow0 and qrow0 + 8
  // per-row bases into sP: index = base_h - kh  and  base_w - kw
  int base_h[2], base_w[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qr = min(qrow0 + r * 8, p.seq_q - 1);
    const int ql = warp * 16 + g + r * 8;
    base_h[r] = ql * pstride + qr / S + S - 1;
    base_w[r] = ql * pstride + n_rel + qr % S + S - 1;
  }
  constexpr float kLog2e = 1.4426950408889634f;

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      load_tile<T, HD>(sK + (buf ^ 1) * TILE_BYTES, kg, p.k_rs, (t + 1) * FA_BN, p.seq_k, tid);
      load_tile<T, HD>(sV + (buf ^ 1) * TILE_BYTES, vg, p.v_rs, (t + 1) * FA_BN, p.seq_k, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    float s[FA_BN / 8][4];
#pragma unroll
    for (int i = 0; i < FA_BN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    const uint32_t kb = sK + buf * TILE_BYTES;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int nb = 0; nb < FA_BN / 16; ++nb) {
        uint32_t r0, r1, r2, r3;
        const int row = nb * 16 + (lane & 7) + (lane >> 4) * 8;
        const int chunk = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(kb + sw_off<HD>(row, chunk), r0, r1, r2, r3);
        mma16816<T>(s[nb * 2], qf[ks], r0, r1);
        mma16816<T>(s[nb * 2 + 1], qf[ks], r2, r3);
      }
    }

    const int key0 = t * FA_BN;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < FA_BN / 8; ++nb) {
#pragma unroll
      for (int e01 = 0; e01 < 2; ++e01) {
        const int key = key0 + nb * 8 + tq * 2 + e01;
        int kh, kw;
        if (S == 64) { kh = key >> 6; kw = key & 63; } else { kh = key / S; kw = key - kh * S; }
        const bool ok = key < p.seq_k;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int e = r * 2 + e01;
          float val = -INFINITY;
          if (ok) val = s[nb][e] * p.scale_log2 + (sP[base_h[r] - kh] + sP[base_w[r] - kw]) * kLog2e;
          s[nb][e] = val;
          mx[r] = fmaxf(mx[r], val);
        }
      }
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mnew[r] = fmaxf(m_run[r], mx[r]);
      const float msafe = (mnew[r] == -INFINITY) ? 0.f : mnew[r];
      corr[r] = exp2f(m_run[r] - msafe);
      m_run[r] = mnew[r];
      mnew[r] = msafe;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[FA_BN / 16][4];
#pragma unroll
    for (int nb = 0; nb < FA_BN / 8; ++nb) {
      const float p0 = exp2f(s[nb][0] - mnew[0]);
      const float p1 = exp2f(s[nb][1] - mnew[0]);
      const float p2 = exp2f(s[nb][2] - mnew[1]);
      const float p3 = exp2f(s[nb][3] - mnew[1]);
      const uint32_t lo = pack2<T>(p0, p1), hi = pack2<T>(p2, p3);
      const float2 flo = unpack2<T>(lo), fhi = unpack2<T>(hi);
      rs[0] += flo.x + flo.y;
      rs[1] += fhi.x + fhi.y;
      pf[nb >> 1][(nb & 1) * 2 + 0] = lo;
      pf[nb >> 1][(nb & 1) * 2 + 1] = hi;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o_acc[i][0] *= corr[0];
      o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1];
      o_acc[i][3] *= corr[1];
    }
    const uint32_t vb = sV + buf * TILE_BYTES;
#pragma unroll
    for (int kk = 0; kk < FA_BN / 16; ++kk) {
#pragma unroll
      for (int db = 0; db < HD / 16; ++db) {
        uint32_t r0, r1, r2, r3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = db * 2 + (lane >> 4);
        ldsm_x4_t(vb + sw_off<HD>(row, chunk), r0, r1, r2, r3);
        mma16816<T>(o_acc[db * 2], pf[kk], r0, r1);
        mma16816<T>(o_acc[db * 2 + 1], pf[kk], r2, r3);
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv[2] = {l_run[0] > 0.f ? 1.f / l_run[0] : 0.f, l_run[1] > 0.f ? 1.f / l_run[1] : 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qr = qrow0 + r * 8;
    if (qr >= p.seq_q) continue;
    T* orow;
    if (rp.o_row_map) {
      const int dst = rp.o_row_map[static_cast<int64_t>(b) * p.seq_q + qr];
      if (dst < 0) continue;  // padded window token: dropped by window_unpartition
      orow = static_cast<T*>(p.o) + static_cast<int64_t>(dst) * p.o_rs + h * p.o_hs;
    } else {
      orow = static_cast<T*>(p.o) + b * p.o_bs + static_cast<int64_t>(qr) * p.o_rs + h * p.o_hs;
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i)
      *reinterpret_cast<uint32_t*>(orow + i * 8 + tq * 2) = pack2<T>(o_acc[i][r * 2] * inv[r], o_acc[i][r * 2 + 1] * inv[r]);
  }
}

template <typename T, int HD>
static int relpos_launch(const FlashParams& p, const RelPosParams& rp, int batch, int heads, cudaStream_t stream) {
  const int smem = 5 * FA_BN * fa_row_bytes<HD>() + FA_BM * (2 * (2 * rp.S - 1) + 1) * 4;
  auto kern = flash_relpos_kernel<T, HD>;
  static SmemOptIn opt_in;   // per device (common.cuh)
  { const int _st = ensure_dynamic_smem(kern, smem, opt_in); if (_st != OK) return _st; }
  dim3 grid((p.seq_q + FA_BM - 1) / FA_BM, heads, batch);
  kern<<<grid, FA_THREADS, smem, stream>>>(p, rp);
  return check_cuda(cudaGetLastError(), "flash_relpos launch");
}

int attention_relpos_run(Context* ctx, const AttnArgs& a, const void* rel_h, const void* rel_w, int S,
                         const int32_t* o_row_map, cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_ATTN_PREFILL, 4.0 * a.batch * a.heads * (double)a.seq_q * a.seq_k * a.head_dim,
                2.0 * a.batch * a.heads * a.head_dim * (2.0 * a.seq_q + 2.0 * a.seq_k));
  ULLAVA_REQUIRE(a.q && a.k && a.v && a.o && rel_h && rel_w, "attention_relpos: null pointer");
  ULLAVA_REQUIRE(S > 0 && S <= 64 && a.seq_q == S * S && a.seq_k == S * S, "attention_relpos: seq must be S*S, S <= 64");
  const int64_t strides[] = {a.q_bs, a.q_rs, a.q_hs, a.k_bs, a.k_rs, a.k_hs, a.v_bs, a.v_rs, a.v_hs, a.o_bs, a.o_rs, a.o_hs};
  for (int64_t st : strides) ULLAVA_REQUIRE(st % 8 == 0, "attention_relpos: strides must be multiples of 8 elements");
  if (a.batch == 0) return OK;
  // 14 x 14 windows with head_dim 80 and back-to-back tables: dedicated single-pass kernel
  if (ctx->attn_impl != 1 && fmha_window_supported(a, rel_h, rel_w, S)) return fmha_window_run(ctx, a, rel_h, o_row_map, stream);
  // auto: the 64 x 64 global grid goes to the online-softmax tcgen05/TMEM kernel; other small grids (two key tiles per
  // CTA, prologue dominated) are faster on the warp-level kernel
  if (ctx->attn_impl != 1 && fmha_supported(a, true, S) && (ctx->attn_impl == 2 || S == 64))
    return fmha_run(ctx, a, rel_h, rel_w, S, o_row_map, stream);
  FlashParams p;
  p.q = a.q; p.q_bs = a.q_bs; p.q_rs = a.q_rs; p.q_hs = a.q_hs;
  p.k = a.k; p.k_bs = a.k_bs; p.k_rs = a.k_rs; p.k_hs = a.k_hs;
  p.v = a.v; p.v_bs = a.v_bs; p.v_rs = a.v_rs; p.v_hs = a.v_hs;
  p.o = a.o; p.o_bs = a.o_bs; p.o_rs = a.o_rs; p.o_hs = a.o_hs;
  p.seq_q = a.seq_q; p.seq_k = a.seq_k; p.causal = 0; p.q_pos0 = 0;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  RelPosParams rp{rel_h, rel_w, S, o_row_map};
  int st;
#define ULLAVA_RP(TT) \
  if (a.head_dim == 80) st = relpos_launch<TT, 80>(p, rp, a.batch, a.heads, stream); \
  else if (a.head_dim == 64) st = relpos_launch<TT, 64>(p, rp, a.batch, a.heads, stream); \
  else if (a.head_dim == 32) st = relpos_launch<TT, 32>(p, rp, a.batch, a.heads, stream); \
  else { set_last_error("attention_relpos: head_dim %d not compiled (32/64/80)", a.head_dim); return ERR_UNSUPPORTED; }
  if (a.dtype == DT_BF16) { ULLAVA_RP(__nv_bfloat16) }
  else if (a.dtype == DT_F16) { ULLAVA_RP(__half) }
  else { set_last_error("attention_relpos: unsupported dtype"); return ERR_UNSUPPORTED; }
#undef ULLAVA_RP
  if (st == OK) ctx->launches++;
  return st;
}

// streaming 16-byte load: read once, do not keep in L1
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// ---- decode attention --------------------------------------------------------------------------
// One CTA per (head, sample).  Phase 1: scores (LPK lanes per key, 16 elements per lane) into shared
// memory; phase 2: block softmax; phase 3: P V with 16-byte V loads, 128/(HD/8) key groups reduced
// through shared memory.  HBM-bound: K and V are each read exactly once.
static constexpr int DEC_THREADS = 128;

// kRope: the kernel also does the step's RoPE + KV-cache write (rope_kvcache_kernel for one new token per sample):
// q points at the PACKED qkv row [q | k | v] (each `qkv_hd` = heads * HD wide); the first HD/4 threads rotate q and k
// of this (sample, head) with the cos / sin row of position ctx_len - 1 (rotate_half,
// hf:models/llama/modeling_llama.py:137-168; same expressions and one rounding to T as rope_kvcache_kernel), write
// rotated k and v into cache row ctx_len - 1 and keep q, k, v in shared memory; the cached rows 0 .. ctx_len - 2 are
// streamed as before and the new row is taken from shared memory.  One kernel less per layer of the decode step.
template <typename T, int HD, bool kRope>
__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_kernel(const T* __restrict__ q, int64_t q_bs, T* __restrict__ kc, T* __restrict__ vc,
                   int64_t cache_bs, int64_t cache_hs, T* __restrict__ o, int64_t o_bs, int ctx_len,
                   float scale_log2, const int32_t* __restrict__ ctx_dev, const float* __restrict__ cos_t,
                   const float* __restrict__ sin_t, int qkv_hd, const int32_t* __restrict__ ctx_off) {
  pdl_launch_dependents();  // the o-projection GEMM may start prefetching its weights under this kernel
  pdl_wait();               // launched with programmatic serialization: resident before its predecessor ends
  if (ctx_dev) ctx_len = *ctx_dev + 1;  // CUDA-graph decode: keys 0..pos are attended, pos read from device memory
  if (ctx_off) ctx_len += ctx_off[blockIdx.y];  // per-sample offset (<= 0): right-padded prompts of different lengths
  extern __shared__ float dec_smem[];   // [ctx_len] scores, then [groups][HD] partial outputs
  __shared__ float red[DEC_THREADS / 32];
  __shared__ float bcast[2];
  __shared__ __align__(16) T s_new[kRope ? 3 * HD : 8];   // rotated q, rotated k, v of the new token
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* qp = q + b * q_bs + h * HD;
  const T* kp = kc + b * cache_bs + h * cache_hs;
  const T* vp = vc + b * cache_bs + h * cache_hs;
  const int ctx_main = kRope ? ctx_len - 1 : ctx_len;   // rows read from the cache

  if constexpr (kRope) {
    constexpr int half = HD / 2, NP = HD / 16;   // NP (lo, hi) chunk pairs of 8 elements per vector
    const int pos = ctx_len - 1;
    if (tid < 2 * NP) {
      const bool is_k = tid >= NP;
      const int j = (is_k ? tid - NP : tid) * 8;
      const T* src = qp + (is_k ? qkv_hd : 0);
      const float* cs = cos_t + static_cast<int64_t>(pos) * half;
      const float* sn = sin_t + static_cast<int64_t>(pos) * half;
      const uint4 lo = *reinterpret_cast<const uint4*>(src + j);
      const uint4 hi = *reinterpret_cast<const uint4*>(src + j + half);
      const uint32_t l[4] = {lo.x, lo.y, lo.z, lo.w}, u[4] = {hi.x, hi.y, hi.z, hi.w};
      uint32_t ol[4], ou[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = unpack2<T>(l[e]), c = unpack2<T>(u[e]);
        const float c0 = cs[j + 2 * e], c1 = cs[j + 2 * e + 1];
        const float s0 = sn[j + 2 * e], s1 = sn[j + 2 * e + 1];
        ol[e] = pack2<T>(a.x * c0 - c.x * s0, a.y * c1 - c.y * s1);
        ou[e] = pack2<T>(c.x * c0 + a.x * s0, c.y * c1 + a.y * s1);
      }
      const uint4 vlo = make_uint4(ol[0], ol[1], ol[2], ol[3]), vhi = make_uint4(ou[0], ou[1], ou[2], ou[3]);
      T* dst = s_new + (is_k ? HD : 0);
      *reinterpret_cast<uint4*>(dst + j) = vlo;
      *reinterpret_cast<uint4*>(dst + j + half) = vhi;
      if (is_k) {
        T* d = kc + b * cache_bs + h * cache_hs + static_cast<int64_t>(pos) * HD + j;
        *reinterpret_cast<uint4*>(d) = vlo;
        *reinterpret_cast<uint4*>(d + half) = vhi;
      }
    } else if (tid < 2 * NP + HD / 8) {
      const int e = (tid - 2 * NP) * 8;
      const uint4 v = *reinterpret_cast<const uint4*>(qp + 2 * qkv_hd + e);
      *reinterpret_cast<uint4*>(s_new + 2 * HD + e) = v;
      *reinterpret_cast<uint4*>(vc + b * cache_bs + h * cache_hs + static_cast<int64_t>(pos) * HD + e) = v;
    }
    __syncthreads();
    qp = s_new;   // the rotated query
  }

  constexpr int LPK = HD / 16;             // lanes per key
  constexpr int KPW = 32 / LPK;            // keys per warp iteration
  const int sub = lane % LPK, kin = lane / LPK;
  float qf[16];
  {
    const uint4 a = *reinterpret_cast<const uint4*>(qp + sub * 16);
    const uint4 c = *reinterpret_cast<const uint4*>(qp + sub * 16 + 8);
    const uint32_t u[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float2 f = unpack2<T>(u[e]);
      qf[2 * e] = f.x;
      qf[2 * e + 1] = f.y;
    }
  }
  float lmax = -INFINITY;
  // KU key rows per lane in flight (2 x 16 B each): the loop is latency-bound otherwise (one CTA holds only 4 warps)
  constexpr int KU = 4;
  constexpr int KSTEP = (DEC_THREADS / 32) * KPW;
  for (int j0 = warp * KPW; j0 < ctx_main; j0 += KU * KSTEP) {
    uint4 ka[KU], kb[KU];
#pragma unroll
    for (int r = 0; r < KU; ++r) {
      const int j = j0 + r * KSTEP + kin;
      if (j < ctx_main) {
        const T* kr = kp + static_cast<int64_t>(j) * HD + sub * 16;
        ka[r] = ld_stream(reinterpret_cast<const uint4*>(kr));
        kb[r] = ld_stream(reinterpret_cast<const uint4*>(kr + 8));
      } else {
        ka[r] = make_uint4(0u, 0u, 0u, 0u);
        kb[r] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int r = 0; r < KU; ++r) {
      const int j = j0 + r * KSTEP + kin;
      const uint32_t u[8] = {ka[r].x, ka[r].y, ka[r].z, ka[r].w, kb[r].x, kb[r].y, kb[r].z, kb[r].w};
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 f = unpack2<T>(u[e]);
        acc += f.x * qf[2 * e] + f.y * qf[2 * e + 1];
      }
#pragma unroll
      for (int off = LPK / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      acc *= scale_log2;
      if (j < ctx_main) {
        if (sub == 0) dec_smem[j] = acc;
        lmax = fmaxf(lmax, acc);
      }
    }
  }
  if constexpr (kRope) {
    if (warp == 0) {   // the new key, from shared memory (its cache row was written by this CTA, not read back)
      const T* kr = s_new + HD + sub * 16;
      const uint4 a = *reinterpret_cast<const uint4*>(kr);
      const uint4 c = *reinterpret_cast<const uint4*>(kr + 8);
      const uint32_t u[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 f = unpack2<T>(u[e]);
        acc += f.x * qf[2 * e] + f.y * qf[2 * e + 1];
      }
#pragma unroll
      for (int off = LPK / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      acc *= scale_log2;
      if (lane == 0) dec_smem[ctx_len - 1] = acc;
      lmax = fmaxf(lmax, acc);
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
    for (int i = 1; i < DEC_THREADS / 32; ++i) m = fmaxf(m, red[i]);
    bcast[0] = m;
  }
  __syncthreads();
  const float m = bcast[0];
  float lsum = 0.f;
  for (int j = tid; j < ctx_len; j += DEC_THREADS) {
    // P rounded to the 16-bit dtype like the eager reference (softmax in fp32, cast, then P V)
    const float pj = exp2f(dec_smem[j] - m);
    lsum += pj;
    dec_smem[j] = pj;
  }
  lsum = warp_sum(lsum);
  __syncthreads();
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < DEC_THREADS / 32; ++i) s += red[i];
    bcast[1] = 1.f / s;
  }
  __syncthreads();
  const float inv = bcast[1];

  constexpr int TPR = HD / 8;              // threads per V row
  constexpr int GROUPS = DEC_THREADS / TPR;
  const int grp = tid / TPR, dv = (tid % TPR) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  constexpr int VU = 8;  // V rows per thread in flight (16 B each); rows are accumulated in increasing j
  for (int j0 = grp; j0 < ctx_main; j0 += VU * GROUPS) {
    uint4 va[VU];
#pragma unroll
    for (int r = 0; r < VU; ++r) {
      const int j = j0 + r * GROUPS;
      va[r] = j < ctx_main ? ld_stream(reinterpret_cast<const uint4*>(vp + static_cast<int64_t>(j) * HD + dv))
                           : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int r = 0; r < VU; ++r) {
      const int j = j0 + r * GROUPS;
      const float pj = j < ctx_main ? T16<T>::to_f(T16<T>::from_f(dec_smem[j] * inv)) : 0.f;
      const uint32_t u[4] = {va[r].x, va[r].y, va[r].z, va[r].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack2<T>(u[e]);
        acc[2 * e] += pj * f.x;
        acc[2 * e + 1] += pj * f.y;
      }
    }
  }
  if constexpr (kRope) {
    // the new value row: last row of its group, exactly where the cache loop would have added it
    if (grp == (ctx_len - 1) % GROUPS) {
      const float pj = T16<T>::to_f(T16<T>::from_f(dec_smem[ctx_len - 1] * inv));
      const uint4 v = *reinterpret_cast<const uint4*>(s_new + 2 * HD + dv);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack2<T>(u[e]);
        acc[2 * e] += pj * f.x;
        acc[2 * e + 1] += pj * f.y;
      }
    }
  }
  __syncthreads();  // scores no longer needed: reuse shared memory for the cross-group reduction
  float* part = dec_smem;
#pragma unroll
  for (int e = 0; e < 8; ++e) part[grp * HD + dv + e] = acc[e];
  __syncthreads();
  for (int d = tid; d < HD; d += DEC_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int gI = 0; gI < GROUPS; ++gI) s += part[gI * HD + d];
    o[b * o_bs + h * HD + d] = T16<T>::from_f(s);
  }
}

template <typename T, int HD, bool kRope>
static int decode_launch_k(const void* q, int64_t q_bs, const void* kc, const void* vc, int64_t cache_bs,
                           int64_t cache_hs, void* o, int64_t o_bs, int batch, int heads, int ctx_len, float scale,
                           cudaStream_t stream, const int32_t* ctx_dev, int max_ctx, bool pdl, const float* cos_t,
                           const float* sin_t, const int32_t* ctx_off) {
  constexpr int GROUPS = DEC_THREADS / (HD / 8);
  const int smem_ctx = ctx_dev ? max_ctx : ctx_len;  // with a device-side length the buffer covers the whole cache
  size_t smem = sizeof(float) * static_cast<size_t>(smem_ctx > GROUPS * HD ? smem_ctx : GROUPS * HD);
  auto kern = attn_decode_kernel<T, HD, kRope>;
  if (smem > 48 * 1024) {
    ULLAVA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  dim3 grid(heads, batch);
  return check_cuda(launch_pdl(kern, grid, dim3(DEC_THREADS), smem, stream, pdl, static_cast<const T*>(q), q_bs,
                               static_cast<T*>(const_cast<void*>(kc)), static_cast<T*>(const_cast<void*>(vc)), cache_bs,
                               cache_hs, static_cast<T*>(o), o_bs, ctx_len, scale * 1.4426950408889634f, ctx_dev, cos_t,
                               sin_t, heads * HD, ctx_off),
                    "attn_decode launch");
}

template <typename T, int HD>
static int decode_launch(const void* q, int64_t q_bs, const void* kc, const void* vc, int64_t cache_bs,
                         int64_t cache_hs, void* o, int64_t o_bs, int batch, int heads, int ctx_len, float scale,
                         cudaStream_t stream, const int32_t* ctx_dev, int max_ctx, bool pdl,
                         const float* cos_t = nullptr, const float* sin_t = nullptr, const int32_t* ctx_off = nullptr) {
  if (cos_t != nullptr)
    return decode_launch_k<T, HD, true>(q, q_bs, kc, vc, cache_bs, cache_hs, o, o_bs, batch, heads, ctx_len, scale,
                                        stream, ctx_dev, max_ctx, pdl, cos_t, sin_t, ctx_off);
  return decode_launch_k<T, HD, false>(q, q_bs, kc, vc, cache_bs, cache_hs, o, o_bs, batch, heads, ctx_len, scale, stream,
                                       ctx_dev, max_ctx, pdl, nullptr, nullptr, ctx_off);
}

// rope_cos / rope_sin != nullptr: fused RoPE + KV-cache write of the new token (q = packed qkv row, see the kernel)
int attention_decode_run(Context* ctx, const void* q, int64_t q_bs, const void* kc, const void* vc, int64_t cache_bs,
                         int64_t cache_hs, void* o, int64_t o_bs, int batch, int heads, int head_dim, int ctx_len,
                         float scale, int dtype, cudaStream_t stream, const int32_t* ctx_dev, int max_ctx,
                         const float* rope_cos, const float* rope_sin, const int32_t* ctx_offset) {
  if (ctx_dev) ctx_len = max_ctx;
  ULLAVA_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr), "attention_decode: cos and sin go together");
  ProfScope _ps(ctx, stream, ULLAVA_PROF_ATTN_DECODE, 4.0 * batch * heads * (double)ctx_len * head_dim, 4.0 * batch * heads * (double)ctx_len * head_dim);
  ULLAVA_REQUIRE(q && kc && vc && o, "attention_decode: null pointer");
  ULLAVA_REQUIRE(ctx_len > 0 && ctx_len <= 16384, "attention_decode: ctx_len %d out of range", ctx_len);
  ULLAVA_REQUIRE(q_bs % 8 == 0 && cache_bs % 8 == 0 && cache_hs % 8 == 0, "attention_decode: bad strides");
  if (batch == 0) return OK;
  int st;
#define ULLAVA_DEC(TT, HDIM) \
  st = decode_launch<TT, HDIM>(q, q_bs, kc, vc, cache_bs, cache_hs, o, o_bs, batch, heads, ctx_len, scale, stream, \
                               ctx_dev, max_ctx, ctx->pdl != 0, rope_cos, rope_sin, ctx_offset)
  if (dtype == DT_BF16) {
    if (head_dim == 128) ULLAVA_DEC(__nv_bfloat16, 128);
    else if (head_dim == 64) ULLAVA_DEC(__nv_bfloat16, 64);
    else if (head_dim == 32) ULLAVA_DEC(__nv_bfloat16, 32);
    else if (head_dim == 16) ULLAVA_DEC(__nv_bfloat16, 16);
    else { set_last_error("attention_decode: head_dim %d not compiled", head_dim); return ERR_UNSUPPORTED; }
  } else if (dtype == DT_F16) {
    if (head_dim == 128) ULLAVA_DEC(__half, 128);
    else if (head_dim == 64) ULLAVA_DEC(__half, 64);
    else if (head_dim == 32) ULLAVA_DEC(__half, 32);
    else if (head_dim == 16) ULLAVA_DEC(__half, 16);
    else { set_last_error("attention_decode: head_dim %d not compiled", head_dim); return ERR_UNSUPPORTED; }
  } else {
    set_last_error("attention_decode: unsupported dtype");
    return ERR_UNSUPPORTED;
  }
#undef ULLAVA_DEC
  if (st == OK) ctx->launches++;
  return st;
}

}  // namespace ullava
