// tcgen05 / TMEM flash attention for the u-LLaVA hot path (sm_100a).
//
//   softmax(scale * Q K^T [+ rel-pos bias] [causal mask]) V      fp32 scores, fp32 online softmax
//
// Call sites it serves (reference file:line):
//   * CLIP ViT-L/14 self-attention, hd 64, S = 577        hf:models/clip/modeling_clip.py:282-336
//   * LLaMA prefill causal attention, hd 128, S = 608     hf:models/llama/modeling_llama.py:199-291
//   * SAM ViT-H windowed (14x14) and global (64x64) attention with decomposed relative-position bias,
//     hd 80                                               segment_anything/modeling/image_encoder.py:196-260,355-392
//
// Work item = one 128-row query tile of one (batch, head).  Without bias the kernel is persistent: one CTA per SM works
// through the items blockIdx.x, blockIdx.x + gridDim.x, ... and all pipelines run on across items (fm_item below); with the
// rel-pos bias one CTA per item.  384 threads in three warpgroups:
//   warps 0-7   softmax, two warps per TMEM lane quadrant (w and w + 4; registers raised to 232 with setmaxnreg):
//                 without bias the two take ALTERNATING key tiles, one thread per row (the row maximum in use is handed
//                 from tile to tile through shared memory, partial row sums are added at the end), so one warp's
//                 exponential pass runs beside the other's TMEM load + maximum pass;
//                 with the rel-pos bias both work on every tile, 64 key columns each (the bias registers do not fit
//                 beside 128 scores), and exchange partial maxima through one 64-thread named barrier per tile.
//               The scores stay in registers between the max pass and the exp pass; 3 of every 8 pairs of exponentials
//               are computed on the FMA pipe (the MUFU, 16 ex2 / clk / SM, is the pipe that bounds a tile), the row sum
//               is kept as partial sums, P is stored to TMEM.  O lives in TMEM for the whole CTA; it is only rescaled
//               when the running maximum grew by more than 2^8 ("lazy rescale" - the stale maximum is used consistently
//               for P and the row sum, so the result is exact), which after the first tiles practically never happens;
//   warp 8      TMA producer: Q once, then K / V tiles of 128 keys through a STAGES-deep ring whose K and V slots are
//               freed separately (cp.async.bulk.tensor.4d over the strided [d, token, head, batch] view, SWIZZLE_128B
//               slabs of 64 columns plus, for hd 80, one SWIZZLE_32B slab of 16 columns);
//   warp 9      TMEM allocator + Q K^T issuer:  S_j = Q K_j^T   (SS form, N = 128) into buffer j % 3 as soon as
//               P_{j-3} V_{j-3} has freed it, i.e. up to three tiles ahead of the softmax;
//   warp 10     P V issuer:  O += P_j V_j   (TS form: A = P_j in TMEM, B = V smem MN-major, N = hd); P_j overwrites the
//               first 64 columns of its own S buffer as packed 16-bit pairs.
//               (Two issuing threads: one thread stalls on its uniform registers until the tensor pipe has taken its
//               instructions over, and spent ~1300 cycles a tile on 512 cycles of tensor work.  Both issue under
//               elect.sync -- under `lane == 0` the compiler wraps every tcgen05.mma in a broadcast loop.)
//   warp 11     idle (its registers go to the softmax warpgroups).
// Rel-pos bias: the prologue runs Q Rh^T and Q Rw^T through the same MMA path into the two S buffers; each softmax
// thread turns its two rows into per-row tables A_h[kh], A_w[kw] (x log2 e, fp32, shared memory), so that in the
// main loop bias(q, k) = A_h[kh(k)] + A_w[kw(k)].  For the 64x64 global grid a key tile is exactly two grid rows:
// A_w sits in 64 registers and A_h costs two shared loads per tile.
#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

static constexpr int FM_BM = 128;       // query rows per CTA (= TMEM lanes)
static constexpr int FM_BN = 128;       // keys per tile
// Softmax warps per TMEM lane quadrant (warps w, w + 4): two.  Four were no faster (profiles/r02_ncu_fmha.md).
static constexpr int fm_tpr(int hd) { return 2 + 0 * hd; }
// Warps 0-7 softmax (two warpgroups), warp 8 TMA producer, warps 9 / 10 Q K^T / P V issuers, warp 11 idle: the third warpgroup
// hands its registers to the softmax warpgroups (setmaxnreg 40 / 232), whose threads hold 128 scores each.
static constexpr int FM_TMA_WARP = 8, FM_MMA_WARP = 9, FM_PV_WARP = 10;   // FM_MMA_WARP: TMEM owner, Q K^T issuer
static constexpr int FM_REGS_SOFTMAX = 232, FM_REGS_OTHER = 40;
static constexpr int fm_threads(int hd) { return 128 * fm_tpr(hd) + 128 + 0 * hd; }
// exchange of the partial row maxima (and, at the end, row sums) among the threads of a row: [tile parity][part][row]
// floats; hd 128 uses one parity and a second barrier per tile
static constexpr int fm_xch_floats(int hd) { return (hd == 128 ? 1 : 2) * fm_tpr(hd) * FM_BM; }
static constexpr int FM_TMEM_COLS = 512;
static constexpr int FM_SBUFS = 3;      // S / P buffers in TMEM
static constexpr int FM_COL_S = 0;      // S buffers: columns [0,128), [128,256), [256,384)
static constexpr int FM_COL_O = 384;    // O accumulator: columns [384, 384 + hd)
static constexpr float FM_RESCALE_THRESHOLD = 8.f;  // log2 units

struct FmhaMaps {
  CUtensorMap q, qt, k, kt, v, vt, rh, rht, rw, rwt;  // *t = tail slab (hd % 64 columns, SWIZZLE_32B)
};

struct FmhaParams {
  void* o;
  int64_t o_bs, o_rs, o_hs;
  const int32_t* o_row_map;
  int seq_q, seq_k, causal, q_pos0;
  float scale_log2;
  int S;  // rel-pos grid side (RP != 0)
  int n_qt, heads, n_items;   // work items = (query tile, head, batch), query tile fastest
  int sms;                    // host side only: CTAs of the persistent launch
  long long* trace;  // debug (ullava_debug_fmha_trace): clock64 stamps of CTA (0, 0, 0), [tile < 64][48], else NULL
};
#define FMHA_TRACE(tile, slot)                                                                      \
  do {                                                                                              \
    if (p.trace && blockIdx.x == 0 && ic == 0 && (tile) < 64)                                       \
      p.trace[(tile) * 48 + (slot)] = clock64();                                                     \
  } while (0)

template <int HD>
struct FmhaCfg {
  static constexpr int NS = HD / 64;                      // 64-column SWIZZLE_128B slabs
  static constexpr int TAIL = HD % 64;                    // 0 or 16 columns in a SWIZZLE_32B slab
  static constexpr int SLAB = 128 * 128;                  // bytes: 128 rows x 128 B
  static constexpr int TAIL_BYTES = 128 * TAIL * 2;
  static constexpr int TILE = NS * SLAB + TAIL_BYTES;     // one Q / K / V tile
  static constexpr int STAGES = HD == 64 ? 4 : 3;
  static constexpr int BAR_BYTES = (1 + 4 * STAGES + 3 * FM_SBUFS + 1 + 1 + 2) * 8 + 16;
  static_assert(TAIL == 0 || TAIL == 16, "head_dim must be 64, 80 or 128");
  static constexpr int smem_bytes(int table_floats) {
    return 1024 + TILE * (1 + 2 * STAGES) + (table_floats + fm_xch_floats(HD)) * 4 + BAR_BYTES;
  }
};

// ---- MMA issue helpers (one thread) -----------------------------------------------------------------------------
template <typename T, int HD>
__device__ __forceinline__ void fmha_issue_qk(uint32_t d_tmem, uint32_t q_smem, uint32_t k_smem) {
  using C = FmhaCfg<HD>;
  constexpr uint32_t idesc = make_idesc_f16(T16<T>::kUmmaFormat, FM_BM, FM_BN);
#pragma unroll
  for (int s = 0; s < C::NS; ++s) {
    const uint64_t a = make_smem_desc(q_smem + s * C::SLAB, 16, 1024, 2);
    const uint64_t b = make_smem_desc(k_smem + s * C::SLAB, 16, 1024, 2);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16<1>(d_tmem, a + 2u * k, b + 2u * k, idesc, (s | k) ? 1u : 0u);
  }
  if constexpr (C::TAIL != 0) {
    const uint64_t a = make_smem_desc(q_smem + C::NS * C::SLAB, 16, 256, 6);
    const uint64_t b = make_smem_desc(k_smem + C::NS * C::SLAB, 16, 256, 6);
    umma_f16<1>(d_tmem, a, b, idesc, 1u);
  }
}

template <typename T, int HD>
__device__ __forceinline__ void fmha_issue_pv(uint32_t o_tmem, uint32_t p_tmem, uint32_t v_smem, bool first_tile) {
  using C = FmhaCfg<HD>;
  // B = V tile [128 keys x hd], hd contiguous: MN-major
  constexpr uint32_t idesc_main = make_idesc_f16(T16<T>::kUmmaFormat, FM_BM, C::NS * 64) | (1u << 16);
  constexpr uint32_t idesc_tail = make_idesc_f16(T16<T>::kUmmaFormat, FM_BM, 16) | (1u << 16);
  // one descriptor per slab kind; 16 keys further = + 2048 B (SWIZZLE_128B slab) / + 512 B (tail slab) in the
  // 16-byte-unit address field, which cannot carry out of its 14 bits for a valid shared address
  const uint64_t b0 = make_smem_desc(v_smem, C::SLAB, 1024, 2);
  const uint64_t bt0 = make_smem_desc(v_smem + C::NS * C::SLAB, 16, 256, 6);
#pragma unroll
  for (int ks = 0; ks < FM_BN / 16; ++ks) {
    const uint32_t acc = (first_tile && ks == 0) ? 0u : 1u;
    umma_f16_ts(o_tmem, p_tmem + ks * 8, b0 + static_cast<uint64_t>(ks * (2048 >> 4)), idesc_main, acc);
    if constexpr (C::TAIL != 0)
      umma_f16_ts(o_tmem + C::NS * 64, p_tmem + ks * 8, bt0 + static_cast<uint64_t>(ks * (512 >> 4)), idesc_tail, acc);
  }
}

// Lazy rescale of O (TMEM) and the running row sum when the row maximum of tile j (`mx`, log2 units) exceeds the
// maximum in use by more than the threshold.  Warp-uniform control flow around the TMEM accesses.
// The TPR threads of a row take the same decision (same mx, same history); thread `part` rescales the O chunks c with
// c % TPR == part.
template <int HD, int TPR>
__device__ __forceinline__ void fmha_rescale(int j, int g, float mx, float& m_used, float& l_run, uint32_t o_taddr,
                                             uint64_t* pv_done, int part) {
  bool grow;
  float alpha = 1.f;
  if (j == 0) {
    m_used = (mx == -INFINITY) ? 0.f : mx;
    grow = false;
  } else {
    grow = mx > m_used + FM_RESCALE_THRESHOLD;
    if (grow) {
      alpha = ex2_approx(m_used - mx);
      m_used = mx;
      l_run *= alpha;
    }
  }
  if (__any_sync(0xffffffffu, grow)) {
    // O holds tiles 0..j-1.  One barrier per S buffer: Q K^T runs FM_SBUFS tiles ahead, so S_j having landed only
    // proves P V of tile j - FM_SBUFS complete, and a single barrier's parity could be two completions behind
    mbar_wait(&pv_done[(g - 1) % FM_SBUFS], static_cast<uint32_t>((g - 1) / FM_SBUFS) & 1u);   // g = running tile index
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) {
      if (c % TPR == part) {
        uint32_t r[16];
        tmem_ld_32x16(o_taddr + c * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
        tmem_st_32x16(o_taddr + c * 16, r);
      }
    }
  }
}

// A CTA works through the items blockIdx.x, blockIdx.x + gridDim.x, ... (one item per CTA for the rel-pos variants); the
// pipelines run on across items: the K / V ring and the three S buffers are indexed by running tile counters, so the
// next item's Q / K / V loads and first Q K^T run under the current item's last tiles and its epilogue.
struct FmItem {
  int m0, h, b, n_tiles;
};
__device__ __forceinline__ FmItem fm_item(const FmhaParams& p, int item) {
  FmItem w;
  const int x = item % p.n_qt;
  const int rest = item / p.n_qt;
  w.h = rest % p.heads;
  w.b = rest / p.heads;
  const int qt = p.causal ? (p.n_qt - 1 - x) : x;   // heavy tiles first
  w.m0 = qt * FM_BM;
  int k_end = p.seq_k;
  if (p.causal) k_end = min(k_end, p.q_pos0 + w.m0 + FM_BM);
  w.n_tiles = (k_end + FM_BN - 1) / FM_BN;
  return w;
}

// RP: 0 = no bias (causal allowed), 1 = rel-pos bias on a generic S x S grid, 2 = rel-pos bias, S == 64
// (RP != 0: one item per CTA -- the item loops below then have a static trip count of one and carry no state)
// EMU: of every 8 pairs of exponentials, EMU are computed on the FMA pipe (ex2_fma2) and 8 - EMU on the MUFU
template <typename T, int HD, int RP, int EMU, bool ALT>
__global__ void __launch_bounds__(fm_threads(HD), 1)
fmha_tcgen05_kernel(const __grid_constant__ FmhaMaps maps, const FmhaParams p) {
  using C = FmhaCfg<HD>;
  constexpr int ST = C::STAGES;
  extern __shared__ uint8_t fm_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fm_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + C::TILE;  // stage s: K at sKV + s*2*TILE, V right after
  const int S = RP ? p.S : 0;
  const int pstride = 2 * S + 1;  // odd: the 32 rows of a warp hit 32 different banks
  float* tabs = reinterpret_cast<float*>(smem + C::TILE * (1 + 2 * ST));
  float* xch = tabs + (RP ? FM_BM * pstride : 0);
  constexpr int TPR = 2;                                    // == fm_tpr(HD)
  constexpr int CPT = ALT ? FM_BN : FM_BN / TPR;            // key columns per softmax thread and tile
  constexpr int kXchFloats = (HD == 128 ? 1 : 2) * TPR * FM_BM;   // == fm_xch_floats(HD)
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + kXchFloats);
  constexpr bool kXchDouble = HD != 128;
  bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bars) + 7) & ~uintptr_t(7));
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* v_full = k_full + ST;
  uint64_t* k_empty = v_full + ST;
  uint64_t* v_empty = k_empty + ST;
  uint64_t* s_full = v_empty + ST;
  uint64_t* p_full = s_full + FM_SBUFS;
  uint64_t* pv_done = p_full + FM_SBUFS;
  uint64_t* pro_done = pv_done + FM_SBUFS;
  uint64_t* o_final = pro_done + 1;
  uint64_t* q_free = o_final + 1;    // the item's last Q K^T has completed: the Q tile may be overwritten
  uint64_t* o_free = q_free + 1;     // the item's O has been read back: the next item's P V may overwrite it
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kRP = RP ? 1 : 0;

  if (warp == FM_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < FM_SBUFS; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], ALT ? 128 : 128 * TPR);
    }
    for (int i = 0; i < FM_SBUFS; ++i) mbar_init(&pv_done[i], 1);
    mbar_init(pro_done, 128 * TPR);
    mbar_init(o_final, 1);
    mbar_init(q_free, 1);
    mbar_init(o_free, 128 * TPR);
    fence_mbar_init();
  }
  if (warp == FM_MMA_WARP) tmem_alloc<1>(tmem_ptr, FM_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= 8) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FM_REGS_OTHER));
  if (warp == FM_TMA_WARP) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int kv_base = kRP;
      for (int item = blockIdx.x, ic = 0; RP != 0 ? ic < 1 : item < p.n_items; item += gridDim.x, ++ic) {
        const FmItem w = fm_item(p, item);
        const int m0 = w.m0, h = w.h, b = w.b, n_tiles = w.n_tiles;
        if (ic > 0) mbar_wait(q_free, static_cast<uint32_t>(ic - 1) & 1u);
        mbar_expect_tx(q_full, C::TILE);
#pragma unroll
        for (int s = 0; s < C::NS; ++s) tma_load_4d(sQ + s * C::SLAB, &maps.q, q_full, s * 64, m0, h, b);
        if constexpr (C::TAIL != 0) tma_load_4d(sQ + C::NS * C::SLAB, &maps.qt, q_full, C::NS * 64, m0, h, b);
        if constexpr (RP != 0) {
          // pipeline slot 0: the two rel-pos tables take the place of a K and a V tile (rows >= 2S-1 are zero-filled)
          uint8_t* sk = sKV;
          uint8_t* sv = sk + C::TILE;
          mbar_expect_tx(&k_full[0], C::TILE);
#pragma unroll
          for (int s = 0; s < C::NS; ++s) tma_load_4d(sk + s * C::SLAB, &maps.rh, &k_full[0], s * 64, 0, 0, 0);
          if constexpr (C::TAIL != 0) tma_load_4d(sk + C::NS * C::SLAB, &maps.rht, &k_full[0], C::NS * 64, 0, 0, 0);
          mbar_expect_tx(&v_full[0], C::TILE);
#pragma unroll
          for (int s = 0; s < C::NS; ++s) tma_load_4d(sv + s * C::SLAB, &maps.rw, &v_full[0], s * 64, 0, 0, 0);
          if constexpr (C::TAIL != 0) tma_load_4d(sv + C::NS * C::SLAB, &maps.rwt, &v_full[0], C::NS * 64, 0, 0, 0);
        }
        // K runs one tile ahead of V: a K slot is free again once its QK^T has completed, a V slot once its P V has
        auto load_k = [&](int j) {
          const int it = kv_base + j, st = it % ST;
          mbar_wait(&k_empty[st], (static_cast<uint32_t>(it / ST) & 1u) ^ 1u);
          FMHA_TRACE(j, 0);
          uint8_t* sk = sKV + st * 2 * C::TILE;
          mbar_expect_tx(&k_full[st], C::TILE);
#pragma unroll
          for (int s = 0; s < C::NS; ++s) tma_load_4d(sk + s * C::SLAB, &maps.k, &k_full[st], s * 64, j * FM_BN, h, b);
          if constexpr (C::TAIL != 0) tma_load_4d(sk + C::NS * C::SLAB, &maps.kt, &k_full[st], C::NS * 64, j * FM_BN, h, b);
        };
        load_k(0);
        for (int j = 0; j < n_tiles; ++j) {
          if (j + 1 < n_tiles) load_k(j + 1);
          const int it = kv_base + j, st = it % ST;
          mbar_wait(&v_empty[st], (static_cast<uint32_t>(it / ST) & 1u) ^ 1u);
          uint8_t* sv = sKV + st * 2 * C::TILE + C::TILE;
          mbar_expect_tx(&v_full[st], C::TILE);
#pragma unroll
          for (int s = 0; s < C::NS; ++s) tma_load_4d(sv + s * C::SLAB, &maps.v, &v_full[st], s * 64, j * FM_BN, h, b);
          if constexpr (C::TAIL != 0) tma_load_4d(sv + C::NS * C::SLAB, &maps.vt, &v_full[st], C::NS * 64, j * FM_BN, h, b);
        }
        kv_base += n_tiles;
      }
    }
  } else if (warp == FM_MMA_WARP) {
    // ===================== MMA issuer (one thread) =====================
    if (elect_one()) {
      const uint32_t q_s = smem_u32(sQ);
      const uint32_t kv_s = smem_u32(sKV);
      int kv_base = kRP, g_base = 0;
      for (int item = blockIdx.x, ic = 0; RP != 0 ? ic < 1 : item < p.n_items; item += gridDim.x, ++ic) {
        const int n_tiles = fm_item(p, item).n_tiles;
        mbar_wait(q_full, static_cast<uint32_t>(ic) & 1u);
        if constexpr (RP != 0) {
          mbar_wait(&k_full[0], 0);
          mbar_wait(&v_full[0], 0);
          tc_fence_after();
          fmha_issue_qk<T, HD>(tmem_base + FM_COL_S, q_s, kv_s);                       // Ph = Q Rh^T
          umma_commit<1>(&s_full[0]);
          fmha_issue_qk<T, HD>(tmem_base + FM_COL_S + FM_BN, q_s, kv_s + C::TILE);     // Pw = Q Rw^T
          umma_commit<1>(&s_full[1]);
          umma_commit<1>(&k_empty[0]);
          umma_commit<1>(&v_empty[0]);
          mbar_wait(pro_done, 0);  // softmax threads have moved both tables out of TMEM
        }
        // S of the g-th tile this CTA processes goes to buffer g % 3 as soon as P V of tile g - 3 has completed (pv_done
        // of that buffer), so Q K^T runs up to three tiles ahead of the softmax -- across item boundaries too -- and a
        // softmax group finds its next S complete when it hands in P.
        // Q K^T and P V are issued by two different warps: the issuing thread stalls on its uniform registers until
        // the tensor pipe has taken the instructions over, and one thread doing both spent ~1300 cycles per tile on
        // 512 cycles of tensor work.
        for (int j = 0; j < n_tiles; ++j) {
          const int it = kv_base + j, st = it % ST;
          const int g = g_base + j, sb = g % FM_SBUFS;
          mbar_wait(&k_full[st], static_cast<uint32_t>(it / ST) & 1u);
          FMHA_TRACE(j, 1);
          if (g >= FM_SBUFS) mbar_wait(&pv_done[sb], static_cast<uint32_t>(g / FM_SBUFS - 1) & 1u);   // P of tile g - 3 consumed
          tc_fence_after();
          fmha_issue_qk<T, HD>(tmem_base + FM_COL_S + sb * FM_BN, q_s, kv_s + st * 2 * C::TILE);
          umma_commit<1>(&s_full[sb]);
          umma_commit<1>(&k_empty[st]);
          FMHA_TRACE(j, 2);
        }
        umma_commit<1>(q_free);
        kv_base += n_tiles;
        g_base += n_tiles;
      }
    }
  } else if (warp == FM_PV_WARP) {
    // ===================== P V issuer (one thread) =====================
    if (elect_one()) {
      const uint32_t kv_s = smem_u32(sKV);
      const uint32_t o_tmem = tmem_base + FM_COL_O;
      int kv_base = kRP, g_base = 0;
      for (int item = blockIdx.x, ic = 0; RP != 0 ? ic < 1 : item < p.n_items; item += gridDim.x, ++ic) {
        const int n_tiles = fm_item(p, item).n_tiles;
        if (ic > 0) mbar_wait(o_free, static_cast<uint32_t>(ic - 1) & 1u);   // the previous item's O is in registers
        for (int j = 0; j < n_tiles; ++j) {
          const int it = kv_base + j, st = it % ST;
          const int g = g_base + j, sb = g % FM_SBUFS;
          mbar_wait(&v_full[st], static_cast<uint32_t>(it / ST) & 1u);
          FMHA_TRACE(j, 3);
          mbar_wait(&p_full[sb], static_cast<uint32_t>(g / FM_SBUFS) & 1u);
          FMHA_TRACE(j, 4);
          tc_fence_after();
          fmha_issue_pv<T, HD>(o_tmem, tmem_base + FM_COL_S + sb * FM_BN, kv_s + st * 2 * C::TILE + C::TILE, j == 0);
          umma_commit<1>(&v_empty[st]);
          umma_commit<1>(&pv_done[sb]);
          FMHA_TRACE(j, 14);
        }
        // the epilogue needs its own barrier: a softmax thread can be several completions behind pv_done
        umma_commit<1>(o_final);
        kv_base += n_tiles;
        g_base += n_tiles;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FM_REGS_SOFTMAX));
    // ===================== softmax / correction / epilogue: two warps per TMEM lane quadrant =====================
    // Softmax warps w and w + 4 share the 32 rows of TMEM lane quadrant w % 4 (a warp may only touch that quadrant)
    // and one SM sub-partition.  Per tile the sub-partition's MUFU is the bound (128 x 128 exponentials at 16 / clk /
    // SM = 1024 cycles against 512 - 1024 cycles of tcgen05 work), so the two warps must keep it busy in turn:
    //   ALT  (group g = 0 / 1 takes the tiles j = g, g + 2, ...; one thread owns a whole row of its tiles): while one
    //        group is in its exponential pass the other loads S of the next tile and finds its row maximum.  The
    //        reference maximum of a row is handed from tile to tile through shared memory (one float per row, a
    //        bar.arrive / bar.sync pair of named barriers per tile and quadrant): tile j reads the maximum in use
    //        after tile j - 1, grows it lazily, publishes it.  Each thread keeps the partial row sum of its own
    //        tiles relative to the last maximum it saw and converts when that changes; the partial sums are added at
    //        the end.
    //   !ALT (both warps work on every tile, 64 key columns each, exchange their partial maxima, one barrier of 64
    //        threads per tile): both warps are in the same phase at the same time -- kept for the rel-pos variants
    //        whose per-thread bias registers do not fit beside 128 scores.
    // The scores stay in registers between the max pass and the exp pass; P (which overwrites the first 64 columns
    // of the S buffer as 16-bit pairs) is only written after every thread of the row has read its scores.
    const int grp = warp >> 2;             // which of the two warps of the quadrant
    const int part = ALT ? 0 : grp;              // which CPT-column part of the tile
    const int quad = warp & 3;                   // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;            // row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t o_taddr = tmem_base + lane_off + FM_COL_O;
    const float sl2 = p.scale_log2;
    float* Ah = tabs + row * pstride;
    float* Aw = Ah + S;
    float aw[RP == 2 ? 64 : 1];                  // S == 64: the bias of grid column kw = key % 64
    constexpr int kOChunks = HD / 16;            // O columns in chunks of 16: chunk c belongs to group c % 2
    constexpr int NH = CPT / 64;                 // 64-column halves (= grid rows when S == 64) per thread and tile
#define FMHA_ROW_SYNC() asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(64) : "memory")
    // ALT: barrier 1 + 2 * quad + g is "group g published"; the publisher arrives, the reader syncs
#define FMHA_PUB_ARRIVE(g) asm volatile("bar.arrive %0, %1;" ::"r"(1 + 2 * quad + (g)), "n"(64) : "memory")
#define FMHA_PUB_SYNC(g) asm volatile("bar.sync %0, %1;" ::"r"(1 + 2 * quad + (g)), "n"(64) : "memory")

    int g_base = 0;                              // tiles this CTA has processed before the current item
    for (int item = blockIdx.x, ic = 0; RP != 0 ? ic < 1 : item < p.n_items; item += gridDim.x, ++ic) {
    const FmItem w = fm_item(p, item);
    const int m0 = w.m0, h = w.h, b = w.b, n_tiles = w.n_tiles;
    const int qrow = m0 + row;
    if constexpr (RP != 0) {
      constexpr float kLog2e = 1.4426950408889634f;
      const int qr = min(qrow, p.seq_q - 1);
      const int qh = qr / S, qw = qr - qh * S;
      mbar_wait(&s_full[0], 0);
      mbar_wait(&s_full[1], 0);
      tc_fence_after();
      {
        // group 0 turns Q Rh^T into A_h, group 1 turns Q Rw^T into A_w
        const int tb = grp;
        const uint32_t ts = tmem_base + lane_off + FM_COL_S + tb * FM_BN;
        float* dst = tb ? Aw : Ah;
        const int base = (tb ? qw : qh) + S - 1;   // table column r  ->  index base - r
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          if (c * 32 >= 2 * S - 1) break;
          uint32_t r[32];
          tmem_ld_32x32(ts + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int idx = base - (c * 32 + i);
            if (idx >= 0 && idx < S) dst[idx] = __uint_as_float(r[i]) * kLog2e;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(pro_done);
      asm volatile("bar.sync %0, %1;" ::"r"(9 + quad), "n"(64) : "memory");   // both tables of the row are complete
      if constexpr (RP == 2) {
#pragma unroll
        for (int i = 0; i < 64; ++i) aw[i] = Aw[i];   // S == 64: my columns start at a multiple of 64
      }
    }

    float m_used = 0.f, l_run = 0.f;
    bool first_mine = true;                      // ALT: l_run is still empty
    // ALT: the groups alternate on the CTA's running tile count, so the alternation carries on across items
    for (int j = ALT ? ((grp ^ g_base) & 1) : 0; j < n_tiles; j += ALT ? 2 : 1) {
      const int g = g_base + j;                  // running tile index: S buffer and barrier phases
      const int buf = g % FM_SBUFS;
      const uint32_t ts = tmem_base + lane_off + FM_COL_S + buf * FM_BN;
      const int key0 = j * FM_BN + CPT * part;   // first key of this thread's part
      const bool need_mask = (j * FM_BN + FM_BN > p.seq_k) || (p.causal && (j * FM_BN + FM_BN - 1 > p.q_pos0 + m0));
      mbar_wait(&s_full[buf], static_cast<uint32_t>(g / FM_SBUFS + (buf < 2 ? kRP : 0)) & 1u);
      if (lane == 0) FMHA_TRACE(j, 32 + warp);
      tc_fence_after();
      uint32_t r[CPT];
#pragma unroll
      for (int c = 0; c < CPT / 32; ++c)
        tmem_ld_32x32(ts + CPT * part + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&r[32 * c]));
      tmem_ld_wait();
      float ah[NH];                              // S == 64: half h of my part lies in grid row 2 j + ...
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) ah[hf] = 0.f;
      if constexpr (RP == 2) {
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) ah[hf] = Ah[min(2 * j + (CPT * part) / 64 + hf, S - 1)];
      }
      const int key_lim = p.causal ? min(p.seq_k, p.q_pos0 + qrow + 1) : p.seq_k;   // keys < key_lim are visible
      int kh0 = 0, kw0 = 0;
      if constexpr (RP == 1) {
        kh0 = key0 / S;
        kw0 = key0 - kh0 * S;
      }

      // ---- pass 1: my scores in log2 units (in place, bias and mask applied; the per-half constant ah of the
      //      64 x 64 grid is added later) and their maximum.  Without bias or mask the raw scores stay as they are
      //      and the scale is folded into pass 2. ----
      const bool raw_scores = (RP == 0) && !need_mask;
      float mx = -INFINITY;
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) {
        float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (raw_scores) {
#pragma unroll
          for (int i = 64 * hf; i < 64 * hf + 64; i += 2)
            mxa[(i >> 1) & 3] = fmaxf(mxa[(i >> 1) & 3], fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
        } else if (RP == 2 && !need_mask) {
          const uint64_t sl2v = pk2(sl2, sl2);
#pragma unroll
          for (int i = 64 * hf; i < 64 * hf + 64; i += 2) {
            const uint64_t x = ffma2(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sl2v,
                                     pk2(aw[i & 63], aw[(i + 1) & 63]));
            float x0, x1;
            upk2(x, x0, x1);
            r[i] = __float_as_uint(x0);
            r[i + 1] = __float_as_uint(x1);
            mxa[(i >> 1) & 3] = fmaxf(mxa[(i >> 1) & 3], fmaxf(x0, x1));
          }
        } else {
          int kh = kh0, kw = kw0;
          if constexpr (RP == 1) {
            if (hf == 1) {                        // NH == 2: advance the grid position by 64 keys
              kw += 64;
              while (kw >= S) { kw -= S; ++kh; }
            }
          }
#pragma unroll   // fully: r[] must be indexed with compile-time constants to stay in registers
          for (int i = 64 * hf; i < 64 * hf + 64; ++i) {
            float x;
            if constexpr (RP == 0) x = __uint_as_float(r[i]) * sl2;
            else if constexpr (RP == 2) x = fmaf(__uint_as_float(r[i]), sl2, aw[i & 63]);
            else x = fmaf(__uint_as_float(r[i]), sl2, Ah[min(kh, S - 1)] + Aw[kw]);
            if (key0 + i >= key_lim) x = -INFINITY;
            r[i] = __float_as_uint(x);
            mxa[i & 3] = fmaxf(mxa[i & 3], x);
            if constexpr (RP == 1) {
              if (++kw == S) { kw = 0; ++kh; }
            }
          }
        }
        const float mh = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
        mx = fmaxf(mx, raw_scores ? mh * sl2 : mh + ah[hf]);   // scale > 0; ah == 0 unless RP == 2
      }

      if constexpr (!ALT) {
        // ---- the row maximum of the whole tile: exchange with the thread that holds the other columns ----
        const int xs = kXchDouble ? (j & 1) * 2 : 0;
        xch[(xs + part) * FM_BM + row] = mx;
        FMHA_ROW_SYNC();
        mx = fmaxf(mx, xch[(xs + (part ^ 1)) * FM_BM + row]);
        if constexpr (!kXchDouble) FMHA_ROW_SYNC();   // single slot: read before the next tile's write
        if (lane == 0) FMHA_TRACE(j, 24 + warp);
        fmha_rescale<HD, 2>(j, g, mx, m_used, l_run, o_taddr, pv_done, part);
      } else {
        // ---- the maximum in use: take it over from tile j - 1 (the other group), grow it lazily, hand it on ----
        float m_prev = 0.f;
        if (j > 0) {
          FMHA_PUB_SYNC(grp ^ 1);
          m_prev = xch[row];
        }
        bool grow = false;
        float m_new;
        if (j == 0) {
          m_new = (mx == -INFINITY) ? 0.f : mx;
        } else {
          grow = mx > m_prev + FM_RESCALE_THRESHOLD;
          m_new = grow ? mx : m_prev;
        }
        xch[row] = m_new;
        FMHA_PUB_ARRIVE(grp);
        if (lane == 0) FMHA_TRACE(j, 24 + warp);
        if (!first_mine && m_used != m_new) l_run *= ex2_approx(m_used - m_new);   // my partial sum follows the maximum
        first_mine = false;
        m_used = m_new;
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? ex2_approx(m_prev - m_new) : 1.f;
          mbar_wait(&pv_done[(g - 1) % FM_SBUFS], static_cast<uint32_t>((g - 1) / FM_SBUFS) & 1u);  // O holds tiles 0..j-1
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < kOChunks; ++c) {
            uint32_t o[16];
            tmem_ld_32x16(o_taddr + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x16(o_taddr + c * 16, o);
          }
        }
      }

      if (lane == 0) FMHA_TRACE(j, 40 + warp);
      // ---- pass 2: p = 2^(x - m), partial row sum, P -> TMEM (packed pairs: words [CPT/2 * part, +CPT/2) of the S buffer) ----
      uint64_t sum[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
      auto exp_pair = [&](float x0, float x1, int slot, uint32_t& packed) {
        float p0, p1;                                               // 2^(-inf) = 0 for masked keys
        if ((slot & 7) < EMU) {
          const uint64_t pe = ex2_fma2<std::is_same<T, __half>::value ? 4 : 3>(x0, x1);
          upk2(pe, p0, p1);
          sum[slot & 1] = fadd2(sum[slot & 1], pe);
        } else {
          p0 = ex2_approx(x0);
          p1 = ex2_approx(x1);
          sum[slot & 1] = fadd2(sum[slot & 1], pk2(p0, p1));
        }
        packed = pack2<T>(p0, p1);
      };
      {
        // one FFMA2 per pair either way (x * 1 + off is exactly x + off): a select between an FFMA2 and an FADD2 made
        // the compiler funnel every exponential through the same two temporaries, a serial chain of ~18 cycles a pair
        const float mul = raw_scores ? sl2 : 1.f;
        const uint64_t sl2v = pk2(mul, mul);
#pragma unroll
        for (int c = 0; c < CPT / 32; ++c) {
          const float off = ah[c / 2] - m_used;
          const uint64_t offv = pk2(off, off);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int e = c * 32 + i;
            const uint64_t x = ffma2(pk2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), sl2v, offv);
            float x0, x1;
            upk2(x, x0, x1);
            exp_pair(x0, x1, i >> 1, pk[i >> 1]);
          }
          tmem_st_32x16(ts + (CPT / 2) * part + c * 16, pk);
        }
      }
      {
        float s0, s1, s2, s3;
        upk2(sum[0], s0, s1);
        upk2(sum[1], s2, s3);
        l_run += (s0 + s1) + (s2 + s3);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[buf]);
      if (lane == 0) FMHA_TRACE(j, 16 + warp);
    }

    // ---- epilogue: O / l -> global; the row sum is the sum of the two threads' partial sums ----
    float inv;
    if constexpr (!ALT) {
      const int ls = kXchDouble ? (n_tiles & 1) * 2 : 0;   // the parity the last tile did not use (its reads may be in flight)
      xch[(ls + part) * FM_BM + row] = l_run;
      FMHA_ROW_SYNC();
      const float l_tot = xch[ls * FM_BM + row] + xch[(ls + 1) * FM_BM + row];
      inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
    } else {
      // the group of the last tile holds the final maximum; the other one converts its partial sum to it
      const int gl = (g_base + n_tiles - 1) & 1;
      if (grp != gl) {
        FMHA_PUB_SYNC(gl);
        const float m_fin = xch[row];
        xch[FM_BM + row] = first_mine ? 0.f : l_run * ex2_approx(m_used - m_fin);
        FMHA_PUB_ARRIVE(grp);
        FMHA_PUB_SYNC(gl);
        inv = xch[FM_BM + row];
      } else {
        FMHA_PUB_SYNC(grp ^ 1);
        const float l_tot = l_run + xch[FM_BM + row];
        inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
        xch[FM_BM + row] = inv;   // not the hand-over slot: the next item's first tile may write that one any time
        FMHA_PUB_ARRIVE(grp);
      }
    }
    T* orow = nullptr;
    if (qrow < p.seq_q) {
      if (p.o_row_map) {
        const int dst = p.o_row_map[static_cast<int64_t>(b) * p.seq_q + qrow];
        if (dst >= 0) orow = static_cast<T*>(p.o) + static_cast<int64_t>(dst) * p.o_rs + h * p.o_hs;
      } else {
        orow = static_cast<T*>(p.o) + b * p.o_bs + static_cast<int64_t>(qrow) * p.o_rs + h * p.o_hs;
      }
    }
    mbar_wait(o_final, static_cast<uint32_t>(ic) & 1u);
    tc_fence_after();
    {
      // this group's chunks (c % 2 == grp): all TMEM loads in flight before the first store
      constexpr int kMine = (kOChunks + 1) / 2;
      uint32_t r[kMine * 16];
#pragma unroll
      for (int i = 0; i < kMine; ++i) {
        const int c = 2 * i + grp;
        if (c < kOChunks) tmem_ld_32x16(o_taddr + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&r[i * 16]));
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free);                       // O is in registers: the next item's P V may start
      if (orow) {
#pragma unroll
        for (int i = 0; i < kMine; ++i) {
          const int c = 2 * i + grp;
          if (c < kOChunks) {
#pragma unroll
            for (int hv = 0; hv < 2; ++hv) {
              const uint32_t* v = &r[i * 16 + hv * 8];
              uint4 w;
              w.x = pack2<T>(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
              w.y = pack2<T>(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
              w.z = pack2<T>(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
              w.w = pack2<T>(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
              *reinterpret_cast<uint4*>(orow + c * 16 + hv * 8) = w;
            }
          }
        }
      }
    }
    g_base += n_tiles;
    }   // items
#undef FMHA_ROW_SYNC
#undef FMHA_PUB_ARRIVE
#undef FMHA_PUB_SYNC
  }

  tc_fence_before();
  __syncthreads();
  if (warp == FM_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, FM_TMEM_COLS);
  }
}

// -----------------------------------------------------------------------------------------------------------------
// Host side
// -----------------------------------------------------------------------------------------------------------------
static int fmha_map(CUtensorMap* main_map, CUtensorMap* tail_map, const void* base, int hd, int64_t rows, int64_t heads,
                    int64_t batch, int64_t rs, int64_t hs, int64_t bs) {
  const uint64_t dims[4] = {static_cast<uint64_t>(hd), static_cast<uint64_t>(rows), static_cast<uint64_t>(heads),
                            static_cast<uint64_t>(batch)};
  // a dimension of extent 1 never uses its stride; keep it a valid multiple of 16 bytes
  const uint64_t st[3] = {static_cast<uint64_t>(rs > 0 ? rs : hd) * 2, static_cast<uint64_t>(hs > 0 ? hs : hd) * 2,
                          static_cast<uint64_t>(bs > 0 ? bs : hd) * 2};
  int e = encode_tmap_4d(main_map, base, dims, st, 64, FM_BM, 128);
  if (e) return e;
  if (hd % 64) e = encode_tmap_4d(tail_map, base, dims, st, hd % 64, FM_BM, 32);
  else *tail_map = *main_map;
  return e;
}

template <typename T, int HD, int RP, int EMU, bool ALT>
static int fmha_launch_emu(const FmhaMaps& maps, const FmhaParams& p, int batch, int heads, cudaStream_t stream) {
  using C = FmhaCfg<HD>;
  const int smem = C::smem_bytes(RP ? FM_BM * (2 * p.S + 1) : 0);
  auto kern = fmha_tcgen05_kernel<T, HD, RP, EMU, ALT>;
  static SmemOptIn opt_in;   // per device (common.cuh)
  { const int _st = ensure_dynamic_smem(kern, smem, opt_in); if (_st != OK) return _st; }
  // bias-free: one CTA per SM works through the items; rel-pos: one item per CTA (its table prologue runs once)
  const int grid = (RP == 0 && p.n_items > p.sms) ? p.sms : p.n_items;
  (void)batch; (void)heads;
  kern<<<grid, fm_threads(HD), smem, stream>>>(maps, p);
  return check_cuda(cudaGetLastError(), "fmha_tcgen05 launch");
}

// Of every 8 pairs of exponentials, FM_EX2_FMA run on the FMA pipe (ex2_fma2, common.cuh).  Swept with -DFM_EX2_FMA=n
// (tools/bench_attn.py, B200): 0 / 2 / 3 / 4 -> 4096 x hd 64: 723 / 735 / 741 / 718 TFLOP/s, SAM global 574 / 600 / 617 /
// 590, CLIP 295 / 304 / 312 / 300.
#ifndef FM_EX2_FMA
#define FM_EX2_FMA 3
#endif

template <typename T, int HD, int RP>
static int fmha_launch(const FmhaMaps& maps, const FmhaParams& p, int batch, int heads, cudaStream_t stream) {
  // one thread per row and alternating tiles (ALT) where the scores are all a thread holds: 740 vs 620 TFLOP/s at hd 64,
  // 1180 vs 1110 at hd 128 (4096 keys).  With the rel-pos bias registers beside them it spills; the two-threads-per-row
  // form stays for those (SAM global attention: 615 TFLOP/s against 557)
  return fmha_launch_emu<T, HD, RP, FM_EX2_FMA, RP == 0>(maps, p, batch, heads, stream);
}

bool fmha_supported(const AttnArgs& a, bool relpos, int S) {
  if (a.head_dim != 64 && a.head_dim != 80 && a.head_dim != 128) return false;
  if (a.dtype != DT_BF16 && a.dtype != DT_F16) return false;
  if (a.batch > 65535 || a.heads > 65535) return false;
  if (relpos && (S < 1 || S > 64 || a.causal)) return false;
  if (relpos && FmhaCfg<80>::smem_bytes(FM_BM * (2 * S + 1)) > 227 * 1024) return false;
  return true;
}

// rel_h / rel_w == nullptr: plain (optionally causal) attention.
int fmha_run(Context* ctx, const AttnArgs& a, const void* rel_h, const void* rel_w, int S, const int32_t* o_row_map,
             cudaStream_t stream) {
  const bool relpos = rel_h != nullptr;
  FmhaMaps maps;
  int e;
  if ((e = fmha_map(&maps.q, &maps.qt, a.q, a.head_dim, a.seq_q, a.heads, a.batch, a.q_rs, a.q_hs, a.q_bs))) return e;
  if ((e = fmha_map(&maps.k, &maps.kt, a.k, a.head_dim, a.seq_k, a.heads, a.batch, a.k_rs, a.k_hs, a.k_bs))) return e;
  if ((e = fmha_map(&maps.v, &maps.vt, a.v, a.head_dim, a.seq_k, a.heads, a.batch, a.v_rs, a.v_hs, a.v_bs))) return e;
  if (relpos) {
    if ((e = fmha_map(&maps.rh, &maps.rht, rel_h, a.head_dim, 2 * S - 1, 1, 1, a.head_dim, 0, 0))) return e;
    if ((e = fmha_map(&maps.rw, &maps.rwt, rel_w, a.head_dim, 2 * S - 1, 1, 1, a.head_dim, 0, 0))) return e;
  } else {
    maps.rh = maps.rht = maps.rw = maps.rwt = maps.q;
  }
  FmhaParams p;
  p.o = a.o; p.o_bs = a.o_bs; p.o_rs = a.o_rs; p.o_hs = a.o_hs;
  p.o_row_map = o_row_map;
  p.seq_q = a.seq_q; p.seq_k = a.seq_k; p.causal = a.causal; p.q_pos0 = a.q_pos0;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.S = S;
  p.n_qt = (a.seq_q + FM_BM - 1) / FM_BM;
  p.heads = a.heads;
  p.n_items = p.n_qt * a.heads * a.batch;
  p.sms = ctx->sm_count;
  p.trace = static_cast<long long*>(ctx->fmha_trace);
  int st = ERR_UNSUPPORTED;
#define ULLAVA_FMHA(TT)                                                                                   \
  if (relpos) {                                                                                            \
    if (a.head_dim == 80 && S == 64) st = fmha_launch<TT, 80, 2>(maps, p, a.batch, a.heads, stream);       \
    else if (a.head_dim == 80) st = fmha_launch<TT, 80, 1>(maps, p, a.batch, a.heads, stream);             \
    else if (a.head_dim == 64) st = fmha_launch<TT, 64, 1>(maps, p, a.batch, a.heads, stream);             \
  } else {                                                                                                 \
    if (a.head_dim == 64) st = fmha_launch<TT, 64, 0>(maps, p, a.batch, a.heads, stream);                  \
    else if (a.head_dim == 80) st = fmha_launch<TT, 80, 0>(maps, p, a.batch, a.heads, stream);             \
    else if (a.head_dim == 128) st = fmha_launch<TT, 128, 0>(maps, p, a.batch, a.heads, stream);           \
  }
  if (a.dtype == DT_BF16) { ULLAVA_FMHA(__nv_bfloat16) }
  else { ULLAVA_FMHA(__half) }
#undef ULLAVA_FMHA
  if (st == ERR_UNSUPPORTED) set_last_error("fmha: head_dim %d / rel-pos combination not compiled", a.head_dim);
  if (st == OK) ctx->launches++;
  return st;
}

}  // namespace ullava
