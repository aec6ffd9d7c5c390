// tcgen05 / TMEM attention for SAM ViT-H *windowed* blocks: 14 x 14 windows (196 tokens), head_dim 80,
// decomposed relative-position bias (segment_anything/modeling/image_encoder.py:196-260, 355-392; window partition
// :263-318).  28 of the 32 encoder blocks use it: 25 windows x 16 heads per image.
//
// A window is short (196 keys), so instead of the online-softmax pipeline of fmha_sm100.cu this kernel does the
// whole row in one pass and spends its effort on hiding latency with a second CTA per SM:
//   * one CTA = one (window, head): K, V and the tables are loaded once and serve both 128-row query tiles (rows
//     0-127, then 128-195), one after the other through the same 256 TMEM columns; 2 CTAs per SM (113 KB shared
//     memory each), so one CTA's TMA latency / softmax runs under the other's MMAs.  (One CTA per query tile loaded
//     K / V twice and paid the whole launch-load-compute-store latency chain per tile: 715 us per encoder block
//     for 32 images against 630 - 650 now -- the chain per tile, not the traffic, is what costs);
//   * ONE MMA chain gives scores AND both bias tables: the B operand is [K (196 rows); pad; Rh (27 rows); Rw (27 rows)]
//     = 254 of 256 rows, so S_ext = Q B^T [128 x 256] holds q.k in columns 0-195, q.Rh in 200-226, q.Rw in 227-253;
//   * softmax: one thread per row; the row's 54 table values go through a private shared-memory row to become
//     A_h[kh], A_w[kw] in 28 registers; the key column -> (kh, kw) mapping is static (unrolled), so the bias costs one
//     packed add per two scores; P (16-bit) overwrites S in place, O accumulates in columns 112-191 of the same
//     TMEM region (free once the scores are consumed);
//   * O += P V with P as the TMEM A operand and V as MN-major B (13 k-steps of 16 keys; V rows >= 196 are TMA
//     zero-fill), output rows scattered through the window_unpartition row map.
#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

static constexpr int FW_S = 14;                 // window side
static constexpr int FW_SEQ = FW_S * FW_S;      // 196 tokens
static constexpr int FW_HD = 80;
static constexpr int FW_NREL = 2 * FW_S - 1;    // 27 table rows
static constexpr int FW_THREADS = 192;
#ifndef FW_EX2_FMA
#define FW_EX2_FMA 0                        // of every 8 pairs of exponentials, how many avoid the MUFU
#endif
static constexpr int FW_TMEM_COLS = 256;
static constexpr int FW_COL_TAB = 200;          // q.Rh at columns 200..226, q.Rw at 227..253
static constexpr int FW_COL_O = 112;            // O accumulator: columns 112..191
static constexpr int FW_KROWS = 256;            // rows of the extended B operand
static constexpr int FW_VROWS = 208;            // 13 k-steps of 16 keys
static constexpr int FW_Q_BYTES = 128 * FW_HD * 2;                // 20480
static constexpr int FW_K_BYTES = FW_KROWS * FW_HD * 2;           // 40960
static constexpr int FW_V_BYTES = FW_VROWS * FW_HD * 2;           // 33280
static constexpr int FW_K_SLAB = FW_KROWS * 128;                  // 32768: 64-column slab, then the 16-column tail
static constexpr int FW_V_SLAB = FW_VROWS * 128;                  // 26624
static constexpr int FW_SCR_STRIDE = 2 * FW_S + 1;                // 29 floats: odd, conflict-free per-row scratch
// the second query tile has 68 rows: its box is 72 rows (the MMA reads 128 -- whatever follows in shared memory gives
// rows nobody looks at), padded so that K stays 1024-byte aligned; two CTAs of 106 KB fit one SM
static constexpr int FW_Q2_ROWS = 72;
static constexpr int FW_Q2_BYTES = FW_Q2_ROWS * FW_HD * 2;         // 11520
static constexpr int FW_Q2_SPACE = 12288;
static constexpr int FW_SMEM = 1024 + FW_Q_BYTES + FW_Q2_SPACE + FW_K_BYTES + FW_V_BYTES + 8 * 8 + 16;
static_assert(128 * FW_SCR_STRIDE * 4 <= FW_Q_BYTES, "the bias scratch lives in the consumed first Q tile");
static_assert(2 * (FW_SMEM + 1024) <= 228 * 1024, "two CTAs per SM");

struct FwMaps {
  CUtensorMap q, qt, q2, q2t, k, kt, v, vt, tab, tabt;  // *t = 16-column tail slab (SWIZZLE_32B); q2 = 72-row box
};

struct FwParams {
  void* o;
  int64_t o_bs, o_rs, o_hs;
  const int32_t* o_row_map;
  float scale_log2;
  long long* trace;   // debug (ullava_debug_fmha_trace): clock64 stamps of CTA (0, 0, 0), [tile][16]
};
#define FW_TRACE(t, slot)                                                                                \
  do {                                                                                                    \
    if (p.trace && blockIdx.y == 0 && blockIdx.z == 0) p.trace[(t) * 16 + (slot)] = clock64();             \
  } while (0)

template <typename T>
__global__ void __launch_bounds__(FW_THREADS, 2)
fmha_window_kernel(const __grid_constant__ FwMaps maps, const FwParams p) {
  extern __shared__ uint8_t fw_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fw_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + FW_Q_BYTES + FW_Q2_SPACE;
  uint8_t* sV = sK + FW_K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FW_V_BYTES);
  uint64_t* qk_full = bars;       // Q + K + tables landed
  uint64_t* v_full = bars + 1;
  uint64_t* s_full = bars + 2;    // S_ext complete
  uint64_t* p_full = bars + 3;    // P written (128 arrivals)
  uint64_t* o_full = bars + 4;    // O complete
  uint64_t* o_read = bars + 5;    // O of the first tile read back: the TMEM columns may take the second tile's scores
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    tma_prefetch_desc(&maps.tab);
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_read, 128);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr, FW_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      FW_TRACE(0, 0);
      mbar_expect_tx(qk_full, FW_Q_BYTES + FW_Q2_BYTES + FW_SEQ * FW_HD * 2 + 2 * FW_NREL * FW_HD * 2);
      tma_load_4d(sQ, &maps.q, qk_full, 0, 0, h, b);
      tma_load_4d(sQ + 128 * 128, &maps.qt, qk_full, 64, 0, h, b);
      tma_load_4d(sQ + FW_Q_BYTES, &maps.q2, qk_full, 0, 128, h, b);                  // rows 128..199, >= 196 zero fill
      tma_load_4d(sQ + FW_Q_BYTES + FW_Q2_ROWS * 128, &maps.q2t, qk_full, 64, 128, h, b);
      tma_load_4d(sK, &maps.k, qk_full, 0, 0, h, b);                                  // rows 0..195
      tma_load_4d(sK + FW_K_SLAB, &maps.kt, qk_full, 64, 0, h, b);
      tma_load_4d(sK + FW_COL_TAB * 128, &maps.tab, qk_full, 0, 0, 0, 0);             // rows 200..253 = [Rh; Rw]
      tma_load_4d(sK + FW_K_SLAB + FW_COL_TAB * 32, &maps.tabt, qk_full, 64, 0, 0, 0);
      mbar_expect_tx(v_full, FW_V_BYTES);
      tma_load_4d(sV, &maps.v, v_full, 0, 0, h, b);                                   // rows >= 196: zero fill
      tma_load_4d(sV + FW_V_SLAB, &maps.vt, v_full, 64, 0, h, b);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (elect_one()) {
      const uint32_t q_s = smem_u32(sQ), k_s = smem_u32(sK), v_s = smem_u32(sV);
      constexpr uint32_t idesc_qk = make_idesc_f16(T16<T>::kUmmaFormat, 128, FW_KROWS);
      constexpr uint32_t idesc_pv = make_idesc_f16(T16<T>::kUmmaFormat, 128, 64) | (1u << 16);
      constexpr uint32_t idesc_pvt = make_idesc_f16(T16<T>::kUmmaFormat, 128, 16) | (1u << 16);
      mbar_wait(qk_full, 0);
      FW_TRACE(0, 1);
      // 16 keys further = + 2048 B / + 512 B in the descriptors' 16-byte-unit address field
      const uint64_t bd0 = make_smem_desc(v_s, 16, 1024, 2);
      const uint64_t bt0 = make_smem_desc(v_s + FW_V_SLAB, 16, 256, 6);
      const uint64_t kd = make_smem_desc(k_s, 16, 1024, 2);
      const uint64_t kt = make_smem_desc(k_s + FW_K_SLAB, 16, 256, 6);
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        if (t) mbar_wait(o_read, 0);
        tc_fence_after();
        {
          const uint64_t a = make_smem_desc(q_s + t * FW_Q_BYTES, 16, 1024, 2);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16<1>(tmem_base, a + 2u * k, kd + 2u * k, idesc_qk, k ? 1u : 0u);
          const uint64_t at = make_smem_desc(q_s + t * FW_Q_BYTES + (t ? FW_Q2_ROWS : 128) * 128, 16, 256, 6);
          umma_f16<1>(tmem_base, at, kt, idesc_qk, 1u);
        }
        umma_commit<1>(s_full);
        FW_TRACE(t, 2);
        if (t == 0) mbar_wait(v_full, 0);
        mbar_wait(p_full, t);
        FW_TRACE(t, 7);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < FW_VROWS / 16; ++ks) {
          umma_f16_ts(tmem_base + FW_COL_O, tmem_base + ks * 8, bd0 + static_cast<uint64_t>(ks * (2048 >> 4)), idesc_pv,
                      ks ? 1u : 0u);
          umma_f16_ts(tmem_base + FW_COL_O + 64, tmem_base + ks * 8, bt0 + static_cast<uint64_t>(ks * (512 >> 4)),
                      idesc_pvt, ks ? 1u : 0u);
        }
        umma_commit<1>(o_full);
        FW_TRACE(t, 8);
      }
    }
  } else {
    // ===================== softmax + epilogue: one thread per query row =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
    const int m0 = t * 128;
    const int qi = m0 + row;
    const bool warp_active = (m0 + quad * 32) < FW_SEQ;   // warp-uniform: any valid row in this warp
    float l_sum = 1.f;

    // destination row first: the row-map lookup is a dependent global load, keep it off the end of the chain
    T* orow = nullptr;
    if (qi < FW_SEQ) {
      if (p.o_row_map) {
        const int dst = p.o_row_map[static_cast<int64_t>(b) * FW_SEQ + qi];
        if (dst >= 0) orow = static_cast<T*>(p.o) + static_cast<int64_t>(dst) * p.o_rs + h * p.o_hs;
      } else {
        orow = static_cast<T*>(p.o) + b * p.o_bs + static_cast<int64_t>(qi) * p.o_rs + h * p.o_hs;
      }
    }

    mbar_wait(s_full, t);
    if (threadIdx.x == 64) FW_TRACE(t, 3);
    tc_fence_after();
    if (warp_active) {
      // ---- bias tables of this row: TMEM -> private shared row -> A_h[kh], A_w[kw] (log2 units) ----
      constexpr float kLog2e = 1.4426950408889634f;
      // the row keeps the 14 + 14 table values its (qh, qw) needs: table index qh .. qh + 13 -> scratch 0 .. 13.  The
      // scratch is the first Q tile (consumed: its S_ext is complete); K must survive for the second tile.
      float* scr = reinterpret_cast<float*>(sQ) + row * FW_SCR_STRIDE;
      const int qc = min(qi, FW_SEQ - 1);
      const int qh = qc / FW_S, qw = qc - qh * FW_S;
      {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(trow + 192, r0);
        tmem_ld_32x32(trow + 224, r1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 2 * FW_NREL; ++i) {     // table value i sits in column 200 + i
          const float v = __uint_as_float(i < 24 ? r0[8 + i] : r1[i - 24]);
          const int dh = i - qh, dw = i - FW_NREL - qw;
          if (i < FW_NREL) {
            if (dh >= 0 && dh < FW_S) scr[dh] = v;
          } else {
            if (dw >= 0 && dw < FW_S) scr[FW_S + dw] = v;
          }
        }
      }
      float ah[FW_S], aw[FW_S];
#pragma unroll
      for (int i = 0; i < FW_S; ++i) {
        ah[i] = scr[FW_S - 1 - i] * kLog2e;
        aw[i] = scr[FW_S + FW_S - 1 - i] * kLog2e;
      }
      const float sl2 = p.scale_log2;
      const uint64_t sl2v = pk2(sl2, sl2);
      if (threadIdx.x == 64) FW_TRACE(t, 4);

      // ---- pass 1: row maximum over the 196 key columns (column -> (kh, kw) is static) ----
      float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(trow + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int col = c * 32 + i;
          if (col < FW_SEQ) {
            const int kh = col / FW_S, kw = col % FW_S;   // kw even: col + 1 is in the same grid row
            const uint64_t bias = fadd2(pk2(aw[kw], aw[kw + 1]), pk2(ah[kh], ah[kh]));
            const uint64_t x = ffma2(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sl2v, bias);
            float x0, x1;
            upk2(x, x0, x1);
            mxa[(i >> 1) & 3] = fmaxf(mxa[(i >> 1) & 3], fmaxf(x0, x1));
          }
        }
      }
      const float m = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
      if (threadIdx.x == 64) FW_TRACE(t, 5);
#pragma unroll
      for (int i = 0; i < FW_S; ++i) ah[i] -= m;

      // ---- pass 2: P = 2^(x - m), row sum, P -> TMEM over the consumed score columns ----
      uint64_t sum[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(trow + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int col = c * 32 + i;
          if (col < FW_SEQ) {
            const int kh = col / FW_S, kw = col % FW_S;
            const uint64_t bias = fadd2(pk2(aw[kw], aw[kw + 1]), pk2(ah[kh], ah[kh]));
            const uint64_t x = ffma2(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sl2v, bias);
            float x0, x1;
            upk2(x, x0, x1);
            float p0, p1;
            if (((i >> 1) & 7) < FW_EX2_FMA) {   // part of the exponentials on the FMA pipe (ex2_fma2, common.cuh)
              const uint64_t pe = ex2_fma2<std::is_same<T, __half>::value ? 4 : 3>(x0, x1);
              upk2(pe, p0, p1);
              sum[(i >> 1) & 1] = fadd2(sum[(i >> 1) & 1], pe);
            } else {
              p0 = ex2_approx(x0);
              p1 = ex2_approx(x1);
              sum[(i >> 1) & 1] = fadd2(sum[(i >> 1) & 1], pk2(p0, p1));
            }
            pk[i >> 1] = pack2<T>(p0, p1);
          } else {
            pk[i >> 1] = 0u;  // keys 196..207 (zero V rows) and the table columns
          }
        }
        tmem_st_32x16(trow + c * 16, pk);
      }
      float s0, s1, s2, s3;
      upk2(sum[0], s0, s1);
      upk2(sum[1], s2, s3);
      l_sum = (s0 + s1) + (s2 + s3);
      tmem_st_wait();
    }
    tc_fence_before();
    mbar_arrive(p_full);
    if (threadIdx.x == 64) FW_TRACE(t, 6);

    // ---- epilogue: O / l -> global (rows scattered through the window_unpartition map) ----
    mbar_wait(o_full, t);
    if (threadIdx.x == 64) FW_TRACE(t, 9);
    tc_fence_after();
    if (warp_active) {
      const float inv = 1.f / l_sum;
      // all five TMEM loads in flight before the first store (one wait instead of five round trips)
      uint32_t r[FW_HD];
#pragma unroll
      for (int c = 0; c < FW_HD / 16; ++c)
        tmem_ld_32x16(trow + FW_COL_O + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&r[c * 16]));
      tmem_ld_wait();
      if (orow) {
#pragma unroll
        for (int c = 0; c < FW_HD / 8; ++c) {
          uint4 w;
          w.x = pack2<T>(__uint_as_float(r[c * 8 + 0]) * inv, __uint_as_float(r[c * 8 + 1]) * inv);
          w.y = pack2<T>(__uint_as_float(r[c * 8 + 2]) * inv, __uint_as_float(r[c * 8 + 3]) * inv);
          w.z = pack2<T>(__uint_as_float(r[c * 8 + 4]) * inv, __uint_as_float(r[c * 8 + 5]) * inv);
          w.w = pack2<T>(__uint_as_float(r[c * 8 + 6]) * inv, __uint_as_float(r[c * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 8) = w;
        }
      }
    }
    if (threadIdx.x == 64) FW_TRACE(t, 10);
    if (t == 0) {
      tc_fence_before();
      mbar_arrive(o_read);
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, FW_TMEM_COLS);
  }
}

// -----------------------------------------------------------------------------------------------------------------
// Host side
// -----------------------------------------------------------------------------------------------------------------
static int fw_map(CUtensorMap* main_map, CUtensorMap* tail_map, const void* base, int64_t rows, int64_t heads,
                  int64_t batch, int64_t rs, int64_t hs, int64_t bs, uint32_t box_rows) {
  const uint64_t dims[4] = {static_cast<uint64_t>(FW_HD), static_cast<uint64_t>(rows), static_cast<uint64_t>(heads),
                            static_cast<uint64_t>(batch)};
  const uint64_t st[3] = {static_cast<uint64_t>(rs > 0 ? rs : FW_HD) * 2, static_cast<uint64_t>(hs > 0 ? hs : FW_HD) * 2,
                          static_cast<uint64_t>(bs > 0 ? bs : FW_HD) * 2};
  int e = encode_tmap_4d(main_map, base, dims, st, 64, box_rows, 128);
  if (e) return e;
  return encode_tmap_4d(tail_map, base, dims, st, 16, box_rows, 32);
}

// Eligible: 14 x 14 grid, head_dim 80, and the two tables stored back to back ([Rh; Rw], as the model packs them).
bool fmha_window_supported(const AttnArgs& a, const void* rel_h, const void* rel_w, int S) {
  if (S != FW_S || a.head_dim != FW_HD || a.seq_q != FW_SEQ || a.seq_k != FW_SEQ || a.causal) return false;
  if (a.dtype != DT_BF16 && a.dtype != DT_F16) return false;
  if (a.batch > 65535 || a.heads > 65535) return false;
  return static_cast<const uint8_t*>(rel_w) == static_cast<const uint8_t*>(rel_h) + FW_NREL * FW_HD * 2;
}

int fmha_window_run(Context* ctx, const AttnArgs& a, const void* rel_h, const int32_t* o_row_map, cudaStream_t stream) {
  FwMaps maps;
  int e;
  if ((e = fw_map(&maps.q, &maps.qt, a.q, FW_SEQ, a.heads, a.batch, a.q_rs, a.q_hs, a.q_bs, 128))) return e;
  if ((e = fw_map(&maps.q2, &maps.q2t, a.q, FW_SEQ, a.heads, a.batch, a.q_rs, a.q_hs, a.q_bs, FW_Q2_ROWS))) return e;
  if ((e = fw_map(&maps.k, &maps.kt, a.k, FW_SEQ, a.heads, a.batch, a.k_rs, a.k_hs, a.k_bs, FW_SEQ))) return e;
  if ((e = fw_map(&maps.v, &maps.vt, a.v, FW_SEQ, a.heads, a.batch, a.v_rs, a.v_hs, a.v_bs, FW_VROWS))) return e;
  if ((e = fw_map(&maps.tab, &maps.tabt, rel_h, 2 * FW_NREL, 1, 1, FW_HD, 0, 0, 2 * FW_NREL))) return e;
  FwParams p;
  p.o = a.o; p.o_bs = a.o_bs; p.o_rs = a.o_rs; p.o_hs = a.o_hs;
  p.o_row_map = o_row_map;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.trace = static_cast<long long*>(ctx->fmha_trace);
  dim3 grid(1, a.heads, a.batch);
  int st;
  if (a.dtype == DT_BF16) {
    auto kern = fmha_window_kernel<__nv_bfloat16>;
    static SmemOptIn opt_in;   // per device (common.cuh)
    { const int _st = ensure_dynamic_smem(kern, FW_SMEM, opt_in); if (_st != OK) return _st; }
    kern<<<grid, FW_THREADS, FW_SMEM, stream>>>(maps, p);
    st = check_cuda(cudaGetLastError(), "fmha_window launch");
  } else {
    auto kern = fmha_window_kernel<__half>;
    static SmemOptIn opt_in;   // per device (common.cuh)
    { const int _st = ensure_dynamic_smem(kern, FW_SMEM, opt_in); if (_st != OK) return _st; }
    kern<<<grid, FW_THREADS, FW_SMEM, stream>>>(maps, p);
    st = check_cuda(cudaGetLastError(), "fmha_window launch");
  }
  if (st == OK) ctx->launches++;
  return st;
}

}  // namespace ullava
