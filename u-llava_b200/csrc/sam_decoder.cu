// SAM prompt encoder (text prompt) + two-way-attention mask decoder + mask post-processing.
//
// Reference: models/segment_anything/modeling/prompt_encoder.py:140-186 (text_embeds path),
//            mask_decoder.py:75-164, transformer.py:16-242, common.py:13-43, sam.py:137-172,
//            call sites models/ullava.py:231-253,403-427.
//
// All prompts of all images of a batch are decoded together (the reference loops image by image
// and repeat_interleaves the 2 MB image embedding per prompt).  Tensors stay token-major
// ([prompt, 64*64, C], C contiguous) from the first kernel to the last, so every dense contraction
// (k/v/q projections of the 4096 image tokens, out_proj, both ConvTranspose2d 2x2/s2 written as
// GEMM + pixel shuffle) is a call of the tcgen05 GEMM; the 6-token side runs through the same GEMM
// (swap-AB small-M path).  Custom kernels here: embedding transpose+dense add, PE adds, token init,
// hyper-network dot product with pixel shuffle, fused double-bilinear post-process.
#include <algorithm>

#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

static constexpr int C = 256;      // transformer dim
static constexpr int NT = 6;       // 1 iou + 4 mask + 1 text token
static constexpr int HW = 64 * 64; // image tokens
static constexpr int NH = 8;       // heads
static constexpr int MLP = 2048;

// ---- keys[p][t][c] = emb[img[p]][c][t] + add[c]   (32x32 smem transpose) ------------------------
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_tokens_kernel(const T* __restrict__ emb, const int32_t* __restrict__ img_of, const T* __restrict__ add,
                      T* __restrict__ out, int channels, int tokens) {
  __shared__ float tile[32][33];
  const int p = blockIdx.z;
  const int im = img_of ? img_of[p] : 0;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const T* src = emb + static_cast<int64_t>(im) * channels * tokens;
  for (int i = ty; i < 32; i += 8) tile[i][tx] = T16<T>::to_f(src[static_cast<int64_t>(c0 + i) * tokens + t0 + tx]);
  __syncthreads();
  T* dst = out + static_cast<int64_t>(p) * tokens * channels;
  for (int i = ty; i < 32; i += 8) {
    float v = tile[tx][i];
    if (add) v = T16<T>::to_f(T16<T>::from_f(v + T16<T>::to_f(add[c0 + tx])));
    dst[static_cast<int64_t>(t0 + i) * channels + c0 + tx] = T16<T>::from_f(v);
  }
}

// ---- out[i] = a[i] + b[i % period]  (16-byte vectors) ---------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
add_bcast_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t nvec, int64_t period_vec) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 x = reinterpret_cast<const uint4*>(a)[i];
    const uint4 y = reinterpret_cast<const uint4*>(b)[i % period_vec];
    const uint32_t xu[4] = {x.x, x.y, x.z, x.w}, yu[4] = {y.x, y.y, y.z, y.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2<T>(xu[e]), g = unpack2<T>(yu[e]);
      o[e] = pack2<T>(f.x + g.x, f.y + g.y);
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---- tokens[p] = cat(iou_token, mask_tokens, text_embeds[p]) ----------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
tokens_init_kernel(const T* __restrict__ iou_tok, const T* __restrict__ mask_tok, const T* __restrict__ text,
                   T* __restrict__ out) {
  const int p = blockIdx.x, c = threadIdx.x;
  T* o = out + static_cast<int64_t>(p) * NT * C;
  o[c] = iou_tok[c];
#pragma unroll
  for (int m = 0; m < 4; ++m) o[(1 + m) * C + c] = mask_tok[m * C + c];
  o[5 * C + c] = text[static_cast<int64_t>(p) * C + c];
}

// ---- masks[p][m][4y+2dy+dy2][4x+2dx+dx2] = sum_c hyper[p][m][c] * up[p][(y,x)][dy][dx][dy2][dx2][c] ----
// up: [n*4096*4, 4*32] (second transposed conv output, (dy2,dx2,c) along the row).  One thread per
// 128-grid pixel (row of `up`): 4 sub-pixels x 32 channels = 128 contiguous 16-bit values.
template <typename T>
__global__ void __launch_bounds__(256)
hyper_mask_kernel(const T* __restrict__ up, const T* __restrict__ hyper, T* __restrict__ masks, int n_prompts) {
  __shared__ float hy[4][32];
  const int p = blockIdx.y;
  if (threadIdx.x < 128) hy[threadIdx.x >> 5][threadIdx.x & 31] = T16<T>::to_f(hyper[p * 128 + threadIdx.x]);
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. 4096*4-1 : (token, dy, dx)
  if (r >= HW * 4) return;
  const int tok = r >> 2, dy = (r >> 1) & 1, dx = r & 1;
  const int y = tok >> 6, x = tok & 63;
  const T* row = up + (static_cast<int64_t>(p) * HW * 4 + r) * 128;
  float acc[4][4];  // [sub][mask]
#pragma unroll
  for (int sp = 0; sp < 4; ++sp) {
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[sp][m] = 0.f;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint4 q = *reinterpret_cast<const uint4*>(row + sp * 32 + v * 8);
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack2<T>(u[e]);
        const int c = v * 8 + e * 2;
#pragma unroll
        for (int m = 0; m < 4; ++m) acc[sp][m] += f.x * hy[m][c] + f.y * hy[m][c + 1];
      }
    }
  }
#pragma unroll
  for (int sp = 0; sp < 4; ++sp) {
    const int Y = 4 * y + 2 * dy + (sp >> 1), X = 4 * x + 2 * dx + (sp & 1);
#pragma unroll
    for (int m = 0; m < 4; ++m)
      masks[((static_cast<int64_t>(p) * 4 + m) * 256 + Y) * 256 + X] = T16<T>::from_f(acc[sp][m]);
  }
}

// ---- post-process: bilinear low_res -> img_size, crop, bilinear -> (out_h, out_w); fp32 -------------
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * (dst + 0.5f) - 0.5f;     // area_pixel_compute_source_index, align_corners=False
  if (s < 0.f) s = 0.f;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - i0;
}

template <typename T>
__global__ void __launch_bounds__(256)
sam_postprocess_kernel(const T* __restrict__ masks, int64_t mask_stride, float* __restrict__ out,
                       uint32_t* __restrict__ bits, int low, int img, int in_h, int in_w, int out_h, int out_w,
                       int words_per_mask) {
  const int p = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = idx < out_h * out_w;
  float val = 0.f;
  if (valid) {
    const int oy = idx / out_w, ox = idx - oy * out_w;
    const T* m = masks + p * mask_stride;
    const float s1 = static_cast<float>(low) / img;            // stage 1: low -> img
    const float s2h = static_cast<float>(in_h) / out_h;         // stage 2: crop -> out
    const float s2w = static_cast<float>(in_w) / out_w;
    int Y0, Y1, X0, X1;
    float ly, lx;
    src_index(s2h, oy, in_h, Y0, Y1, ly);
    src_index(s2w, ox, in_w, X0, X1, lx);
    const int Ys[2] = {Y0, Y1}, Xs[2] = {X0, X1};
    float v[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      int y0, y1;
      float wy;
      src_index(s1, Ys[a], low, y0, y1, wy);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        int x0, x1;
        float wx;
        src_index(s1, Xs[b], low, x0, x1, wx);
        const float m00 = T16<T>::to_f(m[y0 * low + x0]), m01 = T16<T>::to_f(m[y0 * low + x1]);
        const float m10 = T16<T>::to_f(m[y1 * low + x0]), m11 = T16<T>::to_f(m[y1 * low + x1]);
        v[a][b] = (1.f - wy) * ((1.f - wx) * m00 + wx * m01) + wy * ((1.f - wx) * m10 + wx * m11);
      }
    }
    val = (1.f - ly) * ((1.f - lx) * v[0][0] + lx * v[0][1]) + ly * ((1.f - lx) * v[1][0] + lx * v[1][1]);
    out[static_cast<int64_t>(p) * out_h * out_w + idx] = val;
  }
  if (bits) {
    const uint32_t b = __ballot_sync(0xffffffffu, valid && val > 0.f);
    if ((threadIdx.x & 31) == 0 && (idx >> 5) < words_per_mask) bits[static_cast<int64_t>(p) * words_per_mask + (idx >> 5)] = b;
  }
}

int sam_postprocess_run(Context* ctx, const void* masks, int64_t mask_stride, float* out, uint32_t* bits, int n,
                        int low, int img, int in_h, int in_w, int out_h, int out_w, int dtype, cudaStream_t s) {
  ProfScope _ps(ctx, s, ULLAVA_PROF_SAM, 0.0, n * (2.0 * low * low + 4.0 * out_h * out_w));
  ULLAVA_REQUIRE(masks && out, "sam_postprocess: null pointer");
  ULLAVA_REQUIRE(low > 0 && img > 0 && in_h > 0 && in_w > 0 && in_h <= img && in_w <= img && out_h > 0 && out_w > 0,
                 "sam_postprocess: bad geometry");
  if (n == 0) return OK;
  const int total = out_h * out_w;
  const int words = (total + 31) / 32;
  dim3 grid((total + 255) / 256, n);
  if (dtype == DT_BF16)
    sam_postprocess_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(masks), mask_stride,
                                                               out, bits, low, img, in_h, in_w, out_h, out_w, words);
  else if (dtype == DT_F16)
    sam_postprocess_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(masks), mask_stride, out, bits, low,
                                                        img, in_h, in_w, out_h, out_w, words);
  else { set_last_error("sam_postprocess: unsupported dtype"); return ERR_UNSUPPORTED; }
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "sam_postprocess launch");
}

// =================================================================================================
// decoder composition
// =================================================================================================
enum SamW {
  W_IOU_TOKEN = 0, W_MASK_TOKENS, W_NO_MASK,
  W_LAYER0,                 // 36 entries per layer
  // per-layer offsets
  L_SA = 0,                 // q_w q_b k_w k_b v_w v_b o_w o_b
  L_N1 = 8, L_T2I = 10, L_N2 = 18, L_MLP = 20, L_N3 = 24, L_N4 = 26, L_I2T = 28, L_SIZE = 36,
  W_FINAL = W_LAYER0 + 2 * L_SIZE,   // 8 attn + 2 norm
  W_CT1 = W_FINAL + 10,              // ct1_w[(dy,dx,co)=256, 256] ct1_b[256] ln_w[64] ln_b[64] ct2_w[(dy,dx,co)=128, 64] ct2_b[128]
  W_HYPER = W_CT1 + 6,               // 4 x (w0 b0 w1 b1 w2 b2)
  W_IOU = W_HYPER + 24,              // w0 b0 w1 b1 w2 b2
  W_COUNT = W_IOU + 6
};

size_t sam_mask_decoder_scratch(int n) {
  const size_t N = static_cast<size_t>(n);
  size_t t = 0;
  t += 8 * align_up(N * NT * C * 2);         // Q, P, tq, qs, ks, vs, ao, spare
  t += align_up(N * NT * MLP * 2);           // mlp hidden
  t += 2 * align_up(N * HW * C * 2);         // K, kk
  t += 2 * align_up(N * HW * 128 * 2);       // kproj / vproj (reused by i2t)
  t += align_up(static_cast<size_t>(HW) * C * 2);  // key_pe tokens
  t += align_up(N * HW * C * 2);             // ct1 out
  t += align_up(N * HW * 4 * 128 * 2);       // ct2 out
  t += 4 * align_up(N * 4 * C * 2);          // hyper temps
  return t + 8192;
}

namespace {
struct Dec {
  Context* ctx;
  cudaStream_t s;
  int dt;
  const void* const* W;
  int gemm(const void* A, int64_t lda, int wi, void* D, int64_t ldd, int M, int N, int K, int epi = EPI_NONE,
           const void* resid = nullptr, int64_t ldr = 0) {
    GemmArgs a{};
    a.A = A; a.lda = lda; a.B = W[wi]; a.ldb = K; a.D = D; a.ldd = ldd; a.bias = W[wi + 1];
    a.residual = resid; a.ldr = ldr; a.M = M; a.N = N; a.K = K; a.dtype = dt; a.epilogue = epi;
    return gemm_run(ctx, a, s);
  }
  int ln(void* x, int rows, int cols, int wi, float eps, int act = EPI_NONE) {
    return layernorm_run(ctx, x, cols, W[wi], W[wi + 1], x, cols, rows, cols, eps, act, dt, s);
  }
  int add(const void* a, const void* b, void* out, int64_t n_elems, int64_t period) {
    ProfScope _ps(ctx, s, ULLAVA_PROF_SAM, 0.0, 4.0 * n_elems);
    const int64_t nvec = n_elems / 8, pv = period / 8;
    const int grid = static_cast<int>(std::min<int64_t>((nvec + 255) / 256, 148 * 8));
    if (dt == DT_BF16)
      add_bcast_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(a),
                                                          static_cast<const __nv_bfloat16*>(b),
                                                          static_cast<__nv_bfloat16*>(out), nvec, pv);
    else
      add_bcast_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(a), static_cast<const __half*>(b),
                                                   static_cast<__half*>(out), nvec, pv);
    ctx->launches++;
    return check_cuda(cudaGetLastError(), "add_bcast launch");
  }
  // multi-head attention on already projected q/k/v; internal dim `dim`, NH heads
  int attn(const void* q, int sq, const void* k, const void* v, int sk, void* o, int n, int dim) {
    AttnArgs at{};
    const int hd = dim / NH;
    at.q = q; at.q_bs = static_cast<int64_t>(sq) * dim; at.q_rs = dim; at.q_hs = hd;
    at.k = k; at.k_bs = static_cast<int64_t>(sk) * dim; at.k_rs = dim; at.k_hs = hd;
    at.v = v; at.v_bs = static_cast<int64_t>(sk) * dim; at.v_rs = dim; at.v_hs = hd;
    at.o = o; at.o_bs = static_cast<int64_t>(sq) * dim; at.o_rs = dim; at.o_hs = hd;
    at.batch = n; at.heads = NH; at.seq_q = sq; at.seq_k = sk; at.head_dim = hd;
    at.causal = 0; at.q_pos0 = 0; at.scale = 1.0f / sqrtf(static_cast<float>(hd)); at.dtype = dt;
    return attention_run(ctx, at, s);
  }
};
}  // namespace

#define RUN(expr)              \
  do {                         \
    int _st = (expr);          \
    if (_st != OK) return _st; \
  } while (0)

int sam_mask_decoder_run(Context* ctx, const ullava_sam_decoder_args& a, cudaStream_t s) {
  ULLAVA_REQUIRE(a.weights && a.image_embeddings && a.image_pe && a.text_embeds && a.low_res_masks && a.scratch,
                 "sam_mask_decoder: null pointer");
  ULLAVA_REQUIRE(a.n_weights == W_COUNT, "sam_mask_decoder: expected %d weights, got %d", (int)W_COUNT, a.n_weights);
  ULLAVA_REQUIRE(a.dtype == DT_BF16 || a.dtype == DT_F16, "sam_mask_decoder: 16-bit dtypes only");
  const int n = a.n_prompts;
  if (n == 0) return OK;
  ULLAVA_REQUIRE(a.scratch_bytes >= sam_mask_decoder_scratch(n), "sam_mask_decoder: scratch too small (%zu < %zu)",
                 a.scratch_bytes, sam_mask_decoder_scratch(n));
  uint8_t* base = static_cast<uint8_t*>(a.scratch);
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = base + off; off += align_up(bytes); return p; };
  const size_t tokB = static_cast<size_t>(n) * NT * C * 2;
  void* Q = take(tokB);  void* P = take(tokB);  void* tq = take(tokB);
  void* qs = take(tokB); void* ks = take(tokB); void* vs = take(tokB); void* ao = take(tokB); take(tokB);
  void* mh = take(static_cast<size_t>(n) * NT * MLP * 2);
  void* K = take(static_cast<size_t>(n) * HW * C * 2);
  void* kk = take(static_cast<size_t>(n) * HW * C * 2);
  void* kp = take(static_cast<size_t>(n) * HW * 128 * 2);
  void* vp = take(static_cast<size_t>(n) * HW * 128 * 2);
  void* KP = take(static_cast<size_t>(HW) * C * 2);
  void* up1 = take(static_cast<size_t>(n) * HW * C * 2);
  void* up2 = take(static_cast<size_t>(n) * HW * 4 * 128 * 2);
  void* h0 = take(static_cast<size_t>(n) * 4 * C * 2);
  void* h1 = take(static_cast<size_t>(n) * 4 * C * 2);
  void* hyper = take(static_cast<size_t>(n) * 4 * C * 2);  // uses [n][4][32]
  void* iou_t = take(static_cast<size_t>(n) * 4 * C * 2);

  Dec d{ctx, s, a.dtype, a.weights};
  const void* const* W = a.weights;
  const int M6 = n * NT, MK = n * HW;

  // ---- inputs: tokens, image tokens (+ dense no-mask embedding), positional encoding ----
  {
    ProfScope _ps(ctx, s, ULLAVA_PROF_SAM, 0.0, 4.0 * (n + 1.0) * HW * C);
    dim3 g(HW / 32, C / 32, n), g1(HW / 32, C / 32, 1);
    if (a.dtype == DT_BF16) {
      using T = __nv_bfloat16;
      tokens_init_kernel<T><<<n, C, 0, s>>>(static_cast<const T*>(W[W_IOU_TOKEN]), static_cast<const T*>(W[W_MASK_TOKENS]),
                                            static_cast<const T*>(a.text_embeds), static_cast<T*>(Q));
      nchw_to_tokens_kernel<T><<<g, 256, 0, s>>>(static_cast<const T*>(a.image_embeddings), a.prompt_image,
                                                 static_cast<const T*>(W[W_NO_MASK]), static_cast<T*>(K), C, HW);
      nchw_to_tokens_kernel<T><<<g1, 256, 0, s>>>(static_cast<const T*>(a.image_pe), nullptr, nullptr,
                                                  static_cast<T*>(KP), C, HW);
    } else {
      using T = __half;
      tokens_init_kernel<T><<<n, C, 0, s>>>(static_cast<const T*>(W[W_IOU_TOKEN]), static_cast<const T*>(W[W_MASK_TOKENS]),
                                            static_cast<const T*>(a.text_embeds), static_cast<T*>(Q));
      nchw_to_tokens_kernel<T><<<g, 256, 0, s>>>(static_cast<const T*>(a.image_embeddings), a.prompt_image,
                                                 static_cast<const T*>(W[W_NO_MASK]), static_cast<T*>(K), C, HW);
      nchw_to_tokens_kernel<T><<<g1, 256, 0, s>>>(static_cast<const T*>(a.image_pe), nullptr, nullptr,
                                                  static_cast<T*>(KP), C, HW);
    }
    ctx->launches += 3;
    RUN(check_cuda(cudaGetLastError(), "sam decoder init launch"));
    RUN(check_cuda(cudaMemcpyAsync(P, Q, tokB, cudaMemcpyDeviceToDevice, s), "copy point embedding"));
  }

  // ---- two-way transformer (transformer.py:62-106) ----
  for (int l = 0; l < 2; ++l) {
    const int w = W_LAYER0 + l * L_SIZE;
    // (1) token self attention
    const void* qin = Q;
    if (l > 0) { RUN(d.add(Q, P, tq, static_cast<int64_t>(M6) * C, static_cast<int64_t>(M6) * C)); qin = tq; }
    RUN(d.gemm(qin, C, w + L_SA + 0, qs, C, M6, C, C));
    RUN(d.gemm(qin, C, w + L_SA + 2, ks, C, M6, C, C));
    RUN(d.gemm(Q, C, w + L_SA + 4, vs, C, M6, C, C));
    RUN(d.attn(qs, NT, ks, vs, NT, ao, n, C));
    if (l == 0) RUN(d.gemm(ao, C, w + L_SA + 6, Q, C, M6, C, C));
    else RUN(d.gemm(ao, C, w + L_SA + 6, Q, C, M6, C, C, EPI_NONE, Q, C));
    RUN(d.ln(Q, M6, C, w + L_N1, 1e-5f));
    // (2) tokens -> image cross attention (internal dim 128)
    RUN(d.add(Q, P, tq, static_cast<int64_t>(M6) * C, static_cast<int64_t>(M6) * C));
    RUN(d.add(K, KP, kk, static_cast<int64_t>(MK) * C, static_cast<int64_t>(HW) * C));
    RUN(d.gemm(tq, C, w + L_T2I + 0, qs, 128, M6, 128, C));
    RUN(d.gemm(kk, C, w + L_T2I + 2, kp, 128, MK, 128, C));
    RUN(d.gemm(K, C, w + L_T2I + 4, vp, 128, MK, 128, C));
    RUN(d.attn(qs, NT, kp, vp, HW, ao, n, 128));
    RUN(d.gemm(ao, 128, w + L_T2I + 6, Q, C, M6, C, 128, EPI_NONE, Q, C));
    RUN(d.ln(Q, M6, C, w + L_N2, 1e-5f));
    // (3) MLP on tokens
    RUN(d.gemm(Q, C, w + L_MLP + 0, mh, MLP, M6, MLP, C, EPI_RELU));
    RUN(d.gemm(mh, MLP, w + L_MLP + 2, Q, C, M6, C, MLP, EPI_NONE, Q, C));
    RUN(d.ln(Q, M6, C, w + L_N3, 1e-5f));
    // (4) image -> tokens cross attention (kk = K + KP is still valid: K unchanged since step 2)
    RUN(d.add(Q, P, tq, static_cast<int64_t>(M6) * C, static_cast<int64_t>(M6) * C));
    RUN(d.gemm(kk, C, w + L_I2T + 0, kp, 128, MK, 128, C));       // q of the image tokens
    RUN(d.gemm(tq, C, w + L_I2T + 2, ks, 128, M6, 128, C));
    RUN(d.gemm(Q, C, w + L_I2T + 4, vs, 128, M6, 128, C));
    RUN(d.attn(kp, HW, ks, vs, NT, vp, n, 128));
    RUN(d.gemm(vp, 128, w + L_I2T + 6, K, C, MK, C, 128, EPI_NONE, K, C));
    RUN(d.ln(K, MK, C, w + L_N4, 1e-5f));
  }
  // final token -> image attention
  RUN(d.add(Q, P, tq, static_cast<int64_t>(M6) * C, static_cast<int64_t>(M6) * C));
  RUN(d.add(K, KP, kk, static_cast<int64_t>(MK) * C, static_cast<int64_t>(HW) * C));
  RUN(d.gemm(tq, C, W_FINAL + 0, qs, 128, M6, 128, C));
  RUN(d.gemm(kk, C, W_FINAL + 2, kp, 128, MK, 128, C));
  RUN(d.gemm(K, C, W_FINAL + 4, vp, 128, MK, 128, C));
  RUN(d.attn(qs, NT, kp, vp, HW, ao, n, 128));
  RUN(d.gemm(ao, 128, W_FINAL + 6, Q, C, M6, C, 128, EPI_NONE, Q, C));
  RUN(d.ln(Q, M6, C, W_FINAL + 8, 1e-5f));

  // ---- mask head (mask_decoder.py:148-162) ----
  // ConvTranspose2d(256->64,k2,s2) as GEMM with N = (dy,dx,co); LayerNorm2d over co + GELU on 64-wide rows
  RUN(d.gemm(K, C, W_CT1 + 0, up1, 256, MK, 256, C));
  RUN(d.ln(up1, MK * 4, 64, W_CT1 + 2, 1e-6f, EPI_GELU));
  // ConvTranspose2d(64->32,k2,s2) + GELU: rows = 128x128 pixels, N = (dy2,dx2,co2)
  RUN(d.gemm(up1, 64, W_CT1 + 4, up2, 128, MK * 4, 128, 64, EPI_GELU));
  // hyper-network MLPs on the 4 mask tokens (rows of Q with stride NT*C)
  const uint16_t* Q16 = static_cast<const uint16_t*>(Q);
  for (int m = 0; m < 4; ++m) {
    const int w = W_HYPER + 6 * m;
    RUN(d.gemm(Q16 + (1 + m) * C, NT * C, w + 0, h0, C, n, C, C, EPI_RELU));
    RUN(d.gemm(h0, C, w + 2, h1, C, n, C, C, EPI_RELU));
    RUN(d.gemm(h1, C, w + 4, static_cast<uint16_t*>(hyper) + m * 32, 128, n, 32, C));
  }
  {
    ProfScope _ps(ctx, s, ULLAVA_PROF_SAM, 2.0 * n * 4.0 * 32 * 65536, 2.0 * n * (HW * 4.0 * 128 + 4.0 * 65536));
    dim3 g((HW * 4 + 255) / 256, n);
    if (a.dtype == DT_BF16)
      hyper_mask_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(up2),
                                                         static_cast<const __nv_bfloat16*>(hyper),
                                                         static_cast<__nv_bfloat16*>(a.low_res_masks), n);
    else
      hyper_mask_kernel<__half><<<g, 256, 0, s>>>(static_cast<const __half*>(up2), static_cast<const __half*>(hyper),
                                                  static_cast<__half*>(a.low_res_masks), n);
    ctx->launches++;
    RUN(check_cuda(cudaGetLastError(), "hyper_mask launch"));
  }
  if (a.iou_pred) {
    RUN(d.gemm(Q16, NT * C, W_IOU + 0, h0, C, n, C, C, EPI_RELU));
    RUN(d.gemm(h0, C, W_IOU + 2, h1, C, n, C, C, EPI_RELU));
    RUN(d.gemm(h1, C, W_IOU + 4, iou_t, 8, n, 4, C));
    RUN(copy_rows_run(ctx, iou_t, 0, 8, a.iou_pred, 0, 4, 1, n, 4, a.dtype, s));
  }
  return OK;
}

}  // namespace ullava
