// Stage-level entry points: the layer loops of the CLIP ViT, the LLaMA decoder stack and the SAM
// mask decoder run here in C++ (one ctypes call per stage, ~2 us per kernel launch instead of a
// Python round trip per op).  Every kernel they enqueue is one of the hand-written sm_100a kernels
// of this library; nothing here calls cuBLAS/cuDNN/torch.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct Arena {
  uint8_t* base;
  size_t cap, off = 0;
  bool ok = true;
  Arena(void* p, size_t bytes) : base(static_cast<uint8_t*>(p)), cap(bytes) {}
  void* take(size_t bytes) {
    const size_t o = align_up(off);
    if (o + bytes > cap) { ok = false; return nullptr; }
    off = o + bytes;
    return base + o;
  }
};

static int gemm(Context* ctx, cudaStream_t s, int dtype, const void* A, int64_t lda, const void* B, int64_t ldb, void* D,
                int64_t ldd, int M, int N, int K, const void* bias = nullptr, int epi = EPI_NONE,
                const void* resid = nullptr, int64_t ldr = 0, int out_f32 = 0) {
  GemmArgs a{};
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.D = D; a.ldd = ldd;
  a.bias = bias; a.residual = resid; a.ldr = ldr;
  a.M = M; a.N = N; a.K = K; a.dtype = dtype; a.out_f32 = out_f32; a.epilogue = epi;
  return gemm_run(ctx, a, s);
}

static inline void next_weight(Context* ctx, const void* w, int n, int k, int64_t ldb) {
  ctx->next_w = w; ctx->next_n = n; ctx->next_k = k; ctx->next_ldb = ldb;
}

#define RUN(expr)              \
  do {                         \
    int _st = (expr);          \
    if (_st != OK) return _st; \
  } while (0)

// =================================================================================================
// CLIP ViT  (UllavaCoreForCausalLM.encode_image, models/ullava_core.py:146-158;
//            CLIPVisionTransformer hf:models/clip/modeling_clip.py:138-218,282-385,647-690)
// weights: 0 patch_w[hidden,k_pad] 1 cls[hidden] 2 pos[np+1,hidden] 3 pre_ln_w 4 pre_ln_b, then per layer
//          ln1_w ln1_b wqkv[3h,h] bqkv[3h] wo[h,h] bo ln2_w ln2_b w1[ffn,h] b1 w2[h,ffn] b2   (12 per layer)
// =================================================================================================
size_t vit_scratch(int batch, int img, int patch, int hidden, int ffn, int k_pad) {
  const size_t g = img / patch, np = g * g, rows = static_cast<size_t>(batch) * (np + 1);
  size_t t = 0;
  t += align_up(static_cast<size_t>(batch) * np * k_pad * 2);      // im2col
  t += align_up(static_cast<size_t>(batch) * np * hidden * 2);     // patch embeds
  t += 2 * align_up(rows * hidden * 2);                            // x, xn
  t += align_up(rows * 3 * hidden * 2);                            // qkv
  t += align_up(rows * hidden * 2);                                // attn
  t += align_up(rows * ffn * 2);                                   // mlp act
  return t + 4096;
}

int vit_forward_run(Context* ctx, const ullava_vit_args& a, cudaStream_t s) {
  ULLAVA_REQUIRE(a.weights && a.pixels && a.out && a.scratch, "vit_forward: null pointer");
  ULLAVA_REQUIRE(a.n_weights == 5 + 12 * a.layers_used, "vit_forward: expected %d weights, got %d",
                 5 + 12 * a.layers_used, a.n_weights);
  ULLAVA_REQUIRE(a.hidden % a.heads == 0, "vit_forward: hidden %% heads != 0");
  const int g = a.img / a.patch, np = g * g, T = np + 1;
  const int rows = a.batch * T, H = a.hidden, hd = H / a.heads;
  if (a.batch == 0) return OK;
  Arena ar(a.scratch, a.scratch_bytes);
  void* col = ar.take(static_cast<size_t>(a.batch) * np * a.k_pad * 2);
  void* pe = ar.take(static_cast<size_t>(a.batch) * np * H * 2);
  void* x = ar.take(static_cast<size_t>(rows) * H * 2);
  void* xn = ar.take(static_cast<size_t>(rows) * H * 2);
  void* qkv = ar.take(static_cast<size_t>(rows) * 3 * H * 2);
  void* att = ar.take(static_cast<size_t>(rows) * H * 2);
  void* act = ar.take(static_cast<size_t>(rows) * a.ffn * 2);
  if (!ar.ok) { set_last_error("vit_forward: scratch too small (%zu bytes)", a.scratch_bytes); return ERR_WORKSPACE; }
  const void* const* W = a.weights;
  const int dt = a.dtype;

  RUN(vit_im2col_run(ctx, a.pixels, col, a.batch, a.img, a.patch, a.k_pad, dt, s));
  RUN(gemm(ctx, s, dt, col, a.k_pad, W[0], a.k_pad, pe, H, a.batch * np, H, a.k_pad));
  RUN(vit_assemble_run(ctx, pe, W[1], W[2], x, a.batch, np, H, dt, s));
  RUN(layernorm_run(ctx, x, H, W[3], W[4], x, H, rows, H, a.eps, EPI_NONE, dt, s));
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  for (int l = 0; l < a.layers_used; ++l) {
    const void* const* L = W + 5 + 12 * l;
    RUN(layernorm_run(ctx, x, H, L[0], L[1], xn, H, rows, H, a.eps, EPI_NONE, dt, s));
    RUN(gemm(ctx, s, dt, xn, H, L[2], H, qkv, 3 * H, rows, 3 * H, H, L[3]));
    AttnArgs at{};
    const uint16_t* q16 = static_cast<const uint16_t*>(qkv);
    at.q = q16; at.k = q16 + H; at.v = q16 + 2 * H; at.o = att;
    at.q_bs = at.k_bs = at.v_bs = static_cast<int64_t>(T) * 3 * H;
    at.q_rs = at.k_rs = at.v_rs = 3 * H;
    at.q_hs = at.k_hs = at.v_hs = hd;
    at.o_bs = static_cast<int64_t>(T) * H; at.o_rs = H; at.o_hs = hd;
    at.batch = a.batch; at.heads = a.heads; at.seq_q = T; at.seq_k = T; at.head_dim = hd;
    at.causal = 0; at.q_pos0 = 0; at.scale = scale; at.dtype = dt;
    RUN(attention_run(ctx, at, s));
    RUN(gemm(ctx, s, dt, att, H, L[4], H, x, H, rows, H, H, L[5], EPI_NONE, x, H));
    RUN(layernorm_run(ctx, x, H, L[6], L[7], xn, H, rows, H, a.eps, EPI_NONE, dt, s));
    RUN(gemm(ctx, s, dt, xn, H, L[8], H, act, a.ffn, rows, a.ffn, H, L[9], a.act));
    RUN(gemm(ctx, s, dt, act, a.ffn, L[10], a.ffn, x, H, rows, H, a.ffn, L[11], EPI_NONE, x, H));
  }
  // drop CLS: out[b, p] = x[b, 1 + p]
  RUN(copy_rows_run(ctx, static_cast<const uint16_t*>(x) + H, static_cast<int64_t>(T) * H, H, a.out,
                    static_cast<int64_t>(np) * H, H, a.batch, np, H, dt, s));
  return OK;
}

// =================================================================================================
// LLaMA decoder stack (LlamaModel.forward hf:models/llama/modeling_llama.py:355-424; layer :292-333)
// weights: per layer  ln1[H] wqkv[3H,H] wo[H,H] ln2[H] wgu[2F,H] (gate/up interleaved by 16) wdown[H,F]; then final norm[H]
// =================================================================================================
size_t llama_scratch(int rows, int hidden, int ffn) {
  size_t t = 0;
  t += align_up(static_cast<size_t>(rows) * hidden * 2);       // xn
  t += align_up(static_cast<size_t>(rows) * 3 * hidden * 2);   // qkv
  t += align_up(static_cast<size_t>(rows) * hidden * 2);       // attn
  t += align_up(static_cast<size_t>(rows) * ffn * 2);          // act
  return t + 4096;
}

// ---- decode-layer chains (gemm_chain_sm100.cu) -----------------------------------------------------------------
// Program of one decode step: chain 0 = [RMSNorm(ln1_0) -> qkv_0]; after the attention of layer l, chain l + 1 =
// [o_l (+residual) | RMSNorm(ln2_l) -> gate/up_l (SiLU*mul) | down_l (+residual) | RMSNorm(ln1_{l+1}) -> qkv_{l+1}], the
// last one ending in [final RMSNorm -> lm_head (fp32 logits)] instead.  Device layout of the program buffer:
// [2 ints of grid-barrier state per step + one arrival counter per weight tile of every step, padded to 128 B]
// [ChainStep x (1 + 4 * layers)].  The first part is zeroed with ONE memset node at the start of every decode step.
static inline int chain_total_steps(int layers) { return 1 + 4 * layers; }
static inline int chain_tiles(int n) { return (n + 127) / 128; }
static inline size_t chain_state_ints(int layers, int H, int F, int V) {
  return static_cast<size_t>(2) * chain_total_steps(layers) +
         static_cast<size_t>(layers) * (chain_tiles(3 * H) + chain_tiles(H) + chain_tiles(2 * F) + chain_tiles(H)) + chain_tiles(V);
}
static inline size_t chain_state_bytes(int layers, int H, int F, int V) {
  return align_up(chain_state_ints(layers, H, F, V) * sizeof(int), 128);
}
size_t llama_chain_bytes(int layers, int hidden, int ffn, int vocab) {
  return chain_state_bytes(layers, hidden, ffn, vocab) + chain_total_steps(layers) * chain_step_bytes() + 256;
}

static inline uint8_t* chain_base(const ullava_llama_args& a) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a.chain_program) + 127) & ~uintptr_t(127));
}

int llama_chain_prepare_run(Context* ctx, const ullava_decode_args& d) {
  const ullava_llama_args& a = d.llama;
  ULLAVA_REQUIRE(a.seq == 1 && a.final_out && a.chain_program && d.lm_head && d.logits, "chain_prepare: decode-step arguments expected");
  ULLAVA_REQUIRE(a.batch >= 1 && a.batch <= 32, "chain_prepare: batch %d not in 1..32 (larger decode batches take the GEMM-per-kernel path)", a.batch);
  ULLAVA_REQUIRE(a.chain_bytes >= llama_chain_bytes(a.layers, a.hidden_size, a.ffn, d.vocab),
                 "chain_prepare: program buffer too small (%zu < %zu)", a.chain_bytes,
                 llama_chain_bytes(a.layers, a.hidden_size, a.ffn, d.vocab));
  ULLAVA_REQUIRE(a.n_weights == 6 * a.layers + 1 && a.layers >= 1, "chain_prepare: expected %d weights", 6 * a.layers + 1);
  const int rows = a.batch, H = a.hidden_size, F = a.ffn, V = d.vocab;
  const int bn = rows <= 16 ? 16 : 32;
  Arena ar(a.scratch, a.scratch_bytes);
  void* xn = ar.take(static_cast<size_t>(rows) * H * 2);
  void* qkv = ar.take(static_cast<size_t>(rows) * 3 * H * 2);
  void* att = ar.take(static_cast<size_t>(rows) * H * 2);
  void* act = ar.take(static_cast<size_t>(rows) * F * 2);
  if (!ar.ok) { set_last_error("chain_prepare: scratch too small (%zu bytes)", a.scratch_bytes); return ERR_WORKSPACE; }
  const int total = chain_total_steps(a.layers);
  const size_t sb = chain_step_bytes();
  std::vector<uint8_t> host(static_cast<size_t>(total) * sb + 128);
  uint8_t* hs = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(host.data()) + 127) & ~uintptr_t(127));
  int k = 0;
  uint8_t* base = chain_base(a);
  int* tile_counters = reinterpret_cast<int*>(base) + 2 * total;   // behind the grid-barrier state
  auto step = [&](const void* W, int64_t ldb, const void* X, int64_t lda, int N, int K, void* D, int64_t ldd,
                  const void* resid, int ek, int out_f32, const void* nsrc, const void* nw, void* ndst) {
    int* cnt = tile_counters;
    tile_counters += chain_tiles(N);
    return chain_encode_step(hs + static_cast<size_t>(k++) * sb, bn, W, ldb, X, lda, rows, N, K, D, ldd, resid, H, ek,
                             out_f32, nsrc, nw, ndst, H, a.eps, cnt);
  };
  const void* const* W = a.weights;
  RUN(step(W[1], H, xn, H, 3 * H, H, qkv, 3 * H, nullptr, 0, 0, a.hidden, W[0], xn));
  for (int l = 0; l < a.layers; ++l) {
    const void* const* L = W + 6 * l;
    RUN(step(L[2], H, att, H, H, H, a.hidden, H, a.hidden, 0, 0, nullptr, nullptr, nullptr));
    RUN(step(L[4], H, xn, H, 2 * F, H, act, F, nullptr, 1, 0, a.hidden, L[3], xn));
    RUN(step(L[5], F, act, F, H, F, a.hidden, H, a.hidden, 0, 0, nullptr, nullptr, nullptr));
    if (l + 1 < a.layers) {
      RUN(step(W[6 * (l + 1) + 1], H, xn, H, 3 * H, H, qkv, 3 * H, nullptr, 0, 0, a.hidden, W[6 * (l + 1)], xn));
    } else {
      RUN(step(d.lm_head, H, a.final_out, H, V, H, d.logits, V, nullptr, 0, 1, a.hidden, W[6 * a.layers], a.final_out));
    }
  }
  const size_t state = chain_state_bytes(a.layers, H, F, V);
  ULLAVA_REQUIRE(static_cast<size_t>(tile_counters - reinterpret_cast<int*>(base)) * sizeof(int) <= state, "chain_prepare: internal");
  ULLAVA_CHECK_CUDA(cudaMemcpy(base + state, hs, static_cast<size_t>(total) * sb, cudaMemcpyHostToDevice));
  ULLAVA_CHECK_CUDA(cudaMemset(base, 0, state));
  return OK;
}

// One decode step through the chains (a.chain_program prepared for exactly these arguments and this context).
static int llama_decode_chained(Context* ctx, const ullava_llama_args& a, cudaStream_t s, const int32_t* pos_dev, int vocab) {
  const int H = a.hidden_size, hd = a.head_dim;
  Arena ar(a.scratch, a.scratch_bytes);
  ar.take(static_cast<size_t>(a.batch) * H * 2);                      // xn
  void* qkv = ar.take(static_cast<size_t>(a.batch) * 3 * H * 2);
  void* att = ar.take(static_cast<size_t>(a.batch) * H * 2);
  const int dt = a.dtype;
  const int64_t cache_hs = static_cast<int64_t>(a.max_seq) * hd;
  const int64_t cache_bs = cache_hs * a.heads;
  const int64_t layer_stride = cache_bs * a.batch;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  uint16_t* kc0 = static_cast<uint16_t*>(a.k_cache);
  uint16_t* vc0 = static_cast<uint16_t*>(a.v_cache);
  uint8_t* base = chain_base(a);
  int* sync = reinterpret_cast<int*>(base);
  const size_t state = chain_state_bytes(a.layers, H, a.ffn, vocab);
  const uint8_t* steps = base + state;
  const size_t sb = chain_step_bytes();
  // grid-barrier state + tile arrival counters of every chain of this step: one memset node in front of the
  // ~2 * layers kernels
  ULLAVA_CHECK_CUDA(cudaMemsetAsync(sync, 0, state, s));
  const double w_bytes = 2.0 * (4.0 * H * H + 3.0 * H * a.ffn);
  {
    ProfScope _ps(ctx, s, ULLAVA_PROF_GEMM_STREAM, 2.0 * a.batch * 3.0 * H * H, 2.0 * 3.0 * H * H);
    RUN(gemm_chain_run(ctx, steps, 1, a.batch, dt, sync, s));
  }
  for (int l = 0; l < a.layers; ++l) {
    RUN(attention_decode_run(ctx, qkv, 3 * H, kc0 + l * layer_stride, vc0 + l * layer_stride, cache_bs, cache_hs, att, H,
                             a.batch, a.heads, hd, a.pos0 + 1, scale, dt, s, pos_dev, a.max_seq, a.rope_cos, a.rope_sin,
                             a.pos_offset));
    const int first = 1 + 4 * l;
    ProfScope _ps(ctx, s, ULLAVA_PROF_GEMM_STREAM, a.batch * w_bytes, w_bytes);
    RUN(gemm_chain_run(ctx, steps + static_cast<size_t>(first) * sb, 4, a.batch, dt, sync + 2 * first, s));
  }
  return OK;
}

int llama_forward_run(Context* ctx, const ullava_llama_args& a, cudaStream_t s, const int32_t* pos_dev,
                      const void* tail_w, int tail_n) {
  ULLAVA_REQUIRE(pos_dev == nullptr || a.seq == 1, "llama_forward: a device-side position needs seq == 1");
  ULLAVA_REQUIRE(a.weights && a.hidden && a.k_cache && a.v_cache && a.scratch && a.rope_cos && a.rope_sin,
                 "llama_forward: null pointer");
  ULLAVA_REQUIRE(a.n_weights == 6 * a.layers + 1, "llama_forward: expected %d weights, got %d", 6 * a.layers + 1,
                 a.n_weights);
  ULLAVA_REQUIRE(a.hidden_size == a.heads * a.head_dim, "llama_forward: hidden != heads*head_dim");
  ULLAVA_REQUIRE(a.pos0 >= 0 && a.pos0 + a.seq <= a.max_seq, "llama_forward: positions %d..%d exceed the KV cache (%d)",
                 a.pos0, a.pos0 + a.seq, a.max_seq);
  ULLAVA_REQUIRE(a.ffn % 16 == 0, "llama_forward: ffn must be a multiple of 16");
  ULLAVA_REQUIRE(a.pos_offset == nullptr || a.seq == 1, "llama_forward: pos_offset applies to the decode step (seq == 1)");
  const int rows = a.batch * a.seq, H = a.hidden_size, F = a.ffn, hd = a.head_dim;
  if (rows == 0) return OK;
  if (a.chain_program != nullptr) {
    ULLAVA_REQUIRE(a.seq == 1 && a.final_out && !a.all_hidden && a.batch <= 32 && tail_w != nullptr,
                   "llama_forward: a chain program serves the decode step (ullava_llama_decode_step) only");
    return llama_decode_chained(ctx, a, s, pos_dev, tail_n);
  }
  Arena ar(a.scratch, a.scratch_bytes);
  void* xn = ar.take(static_cast<size_t>(rows) * H * 2);
  void* qkv = ar.take(static_cast<size_t>(rows) * 3 * H * 2);
  void* att = ar.take(static_cast<size_t>(rows) * H * 2);
  void* act = ar.take(static_cast<size_t>(rows) * F * 2);
  if (!ar.ok) { set_last_error("llama_forward: scratch too small (%zu bytes)", a.scratch_bytes); return ERR_WORKSPACE; }
  const int dt = a.dtype;
  const int64_t cache_hs = static_cast<int64_t>(a.max_seq) * hd;
  const int64_t cache_bs = cache_hs * a.heads;
  const int64_t layer_stride = cache_bs * a.batch;  // elements
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  uint16_t* kc0 = static_cast<uint16_t*>(a.k_cache);
  uint16_t* vc0 = static_cast<uint16_t*>(a.v_cache);

  for (int l = 0; l < a.layers; ++l) {
    const void* const* L = a.weights + 6 * l;
    uint16_t* kc = kc0 + l * layer_stride;
    uint16_t* vc = vc0 + l * layer_stride;
    if (a.all_hidden) {
      RUN(copy_rows_run(ctx, a.hidden, 0, H, static_cast<uint16_t*>(a.all_hidden) + static_cast<int64_t>(l) * rows * H,
                        0, H, 1, rows, H, dt, s));
    }
    RUN(rmsnorm_run(ctx, a.hidden, H, L[0], xn, H, rows, H, a.eps, dt, s));
    // Prefill at LLaMA-7B geometry: RoPE and the KV-cache write ride in the epilogue of the qkv GEMM (one pass less
    // over q / k / v: ~0.6 GB of HBM traffic per layer at 32 x 608 tokens); k and v then never reach the qkv buffer.
    const bool fuse_rope = a.seq > 1 && hd == 128 && pos_dev == nullptr && gemm_pair_eligible(ctx, rows, 3 * H);
    RopeFuse rf{a.rope_cos, a.rope_sin, kc, vc, cache_bs, cache_hs, a.seq, a.pos0, a.heads};
    if (fuse_rope) ctx->rope_fuse = &rf;
    RUN(gemm(ctx, s, dt, xn, H, L[1], H, qkv, 3 * H, rows, 3 * H, H));
    if (a.seq == 1) {
      // decode step: RoPE of the new q / k and the KV-cache write are fused into the single-query attention kernel
      RUN(attention_decode_run(ctx, qkv, 3 * H, kc, vc, cache_bs, cache_hs, att, H, a.batch, a.heads, hd, a.pos0 + 1,
                               scale, dt, s, pos_dev, a.max_seq, a.rope_cos, a.rope_sin, a.pos_offset));
    } else {
      if (!fuse_rope) {
        RUN(rope_kvcache_run(ctx, qkv, 3 * H, kc, vc, cache_bs, cache_hs, a.batch, a.seq, a.heads, hd, a.pos0,
                             a.rope_cos, a.rope_sin, dt, s, pos_dev));
      }
      AttnArgs at{};
      at.q = qkv; at.q_bs = static_cast<int64_t>(a.seq) * 3 * H; at.q_rs = 3 * H; at.q_hs = hd;
      at.k = kc; at.k_bs = cache_bs; at.k_rs = hd; at.k_hs = cache_hs;
      at.v = vc; at.v_bs = cache_bs; at.v_rs = hd; at.v_hs = cache_hs;
      at.o = att; at.o_bs = static_cast<int64_t>(a.seq) * H; at.o_rs = H; at.o_hs = hd;
      at.batch = a.batch; at.heads = a.heads; at.seq_q = a.seq; at.seq_k = a.pos0 + a.seq; at.head_dim = hd;
      at.causal = 1; at.q_pos0 = a.pos0; at.scale = scale; at.dtype = dt;
      RUN(attention_run(ctx, at, s));
    }
    // Decode (rows <= 32): each weight-streaming GEMM pulls the head of the NEXT weight matrix into L2 under its own
    // tail.  (Not across the attention kernel: its KV stream would evict the lines before W_o is read.)
    const bool hint = rows <= 32;
    if (hint) next_weight(ctx, L[4], 2 * F, H, H);
    RUN(gemm(ctx, s, dt, att, H, L[2], H, a.hidden, H, rows, H, H, nullptr, EPI_NONE, a.hidden, H));
    RUN(rmsnorm_run(ctx, a.hidden, H, L[3], xn, H, rows, H, a.eps, dt, s));
    if (hint) next_weight(ctx, L[5], H, F, F);
    RUN(gemm(ctx, s, dt, xn, H, L[4], H, act, F, rows, 2 * F, H, nullptr, EPI_SILU_MUL));
    if (hint) {
      if (l + 1 < a.layers) next_weight(ctx, a.weights[6 * (l + 1) + 1], 3 * H, H, H);
      else if (tail_w) next_weight(ctx, tail_w, tail_n, H, H);
    }
    RUN(gemm(ctx, s, dt, act, F, L[5], F, a.hidden, H, rows, H, F, nullptr, EPI_NONE, a.hidden, H));
  }
  if (a.final_out) {
    RUN(rmsnorm_run(ctx, a.hidden, H, a.weights[6 * a.layers], a.final_out, H, rows, H, a.eps, dt, s));
  }
  return OK;
}

// =================================================================================================
// One greedy decode step with the position in device memory (CUDA-graph replayable):
// embed(cur_ids) -> decoder stack (seq 1) -> final norm -> lm_head -> argmax -> bookkeeping.
// Replaces one iteration of GenerationMixin.generate's loop as driven by UllavaForCausalLM.evaluate
// (models/ullava.py:350-362 -> models/ullava_core.py:357-395,279-355).
// =================================================================================================
int llama_decode_step_run(Context* ctx, const ullava_decode_args& a, cudaStream_t s) {
  const ullava_llama_args& L = a.llama;
  ULLAVA_REQUIRE(L.seq == 1 && L.final_out, "decode_step: llama.seq must be 1 and final_out set");
  ULLAVA_REQUIRE(a.pos_dev && a.embed_table && a.lm_head && a.cur_ids && a.logits, "decode_step: null pointer");
  RUN(embed_gather_run(ctx, a.cur_ids, a.embed_table, L.hidden, L.batch, L.hidden_size, a.vocab, L.dtype, s));
  RUN(llama_forward_run(ctx, L, s, a.pos_dev, a.lm_head, a.vocab));
  if (L.chain_program == nullptr) {   // (the last chain of the program ends with final norm -> lm_head)
    GemmArgs g{};
    g.A = L.final_out; g.lda = L.hidden_size; g.B = a.lm_head; g.ldb = L.hidden_size; g.D = a.logits; g.ldd = a.vocab;
    g.M = L.batch; g.N = a.vocab; g.K = L.hidden_size; g.dtype = L.dtype; g.out_f32 = 1; g.epilogue = EPI_NONE;
    RUN(gemm_run(ctx, g, s));
  }
  if (a.uniforms) {
    RUN(sample_step_run(ctx, a.logits, a.vocab, L.batch, a.vocab, a.temperature, a.top_p, a.top_k, a.uniforms, a.uniforms_ld,
                        a.cur_ids, a.seqs, a.seqs_ld, L.final_out, a.hid_buf, a.hid_bs, L.hidden_size, a.finished,
                        a.eos_id, a.pad_id, a.pos_dev, nullptr, s));
  } else {
    RUN(greedy_step_run(ctx, a.logits, a.vocab, L.batch, a.vocab, a.cur_ids, a.seqs, a.seqs_ld, L.final_out, a.hid_buf,
                        a.hid_bs, L.hidden_size, a.finished, a.eos_id, a.pad_id, a.pos_dev, s));
  }
  return OK;
}

}  // namespace ullava
