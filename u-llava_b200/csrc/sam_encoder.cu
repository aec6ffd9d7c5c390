// SAM ViT image encoder (ImageEncoderViT.forward, segment_anything/modeling/image_encoder.py:110-125;
// Block :128-193, Attention :196-260, window partition :263-318, rel-pos :321-392, PatchEmbed :395-426;
// call site models/ullava.py:139-150) on a batch of images, tokens kept channels-last from the patch
// embedding to the neck.
//
//   patch embed   : im2col (16x16, stride 16) + tcgen05 GEMM, bias and abs. position embedding fused in the epilogue
//   block x depth : LayerNorm (window blocks: output rows scattered straight into the zero-padded 14x14
//                   window layout) -> qkv GEMM (+bias) -> flash attention with the decomposed rel-pos bias
//                   (window blocks: output rows gathered back to the token grid, pads dropped) ->
//                   proj GEMM (+bias, +residual) -> LayerNorm -> lin1 GEMM (+bias, erf-GELU) -> lin2 GEMM (+bias, +residual)
//   neck          : 1x1 conv as GEMM -> LayerNorm2d -> 3x3 conv as channels-last im2col + GEMM -> LayerNorm2d -> NCHW
//
// weights: 0 patch_w[D, 3*p*p] 1 patch_b[D] 2 pos_embed[g*g, D]; per block (14): n1_w n1_b qkv_w[3D,D] qkv_b relh[2S-1,hd]
//          relw[2S-1,hd] proj_w proj_b n2_w n2_b lin1_w lin1_b lin2_w lin2_b; neck (6): conv1_w[C,D] ln1_w ln1_b
//          conv2_w[C, 9*C] (ky,kx,ci order) ln2_w ln2_b
#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// ---- channels-last 3x3 im2col with zero padding: in [B, g, g, C] -> out [B*g*g, 9*C], column = (ky*3+kx)*C + c ----
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int g, int C, int64_t total_vec) {
  const int vpc = C >> 3;  // 16-byte vectors per pixel
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpc);
    int64_t r = i / vpc;
    const int tap = static_cast<int>(r % 9);
    r /= 9;                                   // output row = b*g*g + y*g + x
    const int x = static_cast<int>(r % g);
    const int y = static_cast<int>((r / g) % g);
    const int64_t b = r / (static_cast<int64_t>(g) * g);
    const int sy = y + tap / 3 - 1, sx = x + tap % 3 - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (sy >= 0 && sy < g && sx >= 0 && sx < g)
      val = *reinterpret_cast<const uint4*>(in + ((b * g + sy) * g + sx) * C + v * 8);
    reinterpret_cast<uint4*>(out)[i] = val;
  }
}

// ---- tokens [B, N, C] -> NCHW [B, C, N] (32x32 smem transpose) ----
__global__ void __launch_bounds__(256)
tokens_to_nchw_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int N, int C) {
  __shared__ uint16_t tile[32][34];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const uint16_t* src = in + static_cast<int64_t>(b) * N * C;
  for (int i = ty; i < 32; i += 8) tile[i][tx] = src[static_cast<int64_t>(t0 + i) * C + c0 + tx];
  __syncthreads();
  uint16_t* dst = out + static_cast<int64_t>(b) * N * C;
  for (int i = ty; i < 32; i += 8) dst[static_cast<int64_t>(c0 + i) * N + t0 + tx] = tile[tx][i];
}

size_t sam_encoder_scratch(int batch, int img, int patch, int D, int window, int C) {
  const size_t g = img / patch, N = g * g, rows = static_cast<size_t>(batch) * N;
  const size_t nw = (g + window - 1) / window, wrows = static_cast<size_t>(batch) * nw * nw * window * window;
  const size_t big = std::max(rows, wrows);
  size_t t = 0;
  t += 2 * align_up(rows * D * 2);                                  // x, xn
  t += align_up(wrows * D * 2);                                     // xw
  t += align_up(big * 3 * D * 2);                                   // qkv
  t += align_up(rows * D * 2);                                      // att
  t += align_up(std::max(rows * 4 * D, rows * 3 * (size_t)patch * patch) * 2);  // mlp act / patch im2col
  t += 2 * align_up(rows * C * 2);                                  // neck 1, neck 2
  t += align_up(rows * 9 * C * 2);                                  // neck im2col
  return t + 8192;
}

#define RUN(expr)              \
  do {                         \
    int _st = (expr);          \
    if (_st != OK) return _st; \
  } while (0)

static int gemm(Context* ctx, cudaStream_t s, int dt, const void* A, int64_t lda, const void* B, int64_t ldb, void* D,
                int64_t ldd, int M, int N, int K, const void* bias, int epi, const void* resid, int64_t ldr) {
  GemmArgs a{};
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.D = D; a.ldd = ldd; a.bias = bias; a.residual = resid; a.ldr = ldr;
  a.M = M; a.N = N; a.K = K; a.dtype = dt; a.epilogue = epi;
  return gemm_run(ctx, a, s);
}

int sam_encoder_run(Context* ctx, const ullava_sam_encoder_args& a, cudaStream_t s) {
  ULLAVA_REQUIRE(a.weights && a.pixels && a.out && a.scratch, "sam_encoder: null pointer");
  ULLAVA_REQUIRE(a.n_weights == 3 + 14 * a.depth + 6, "sam_encoder: expected %d weights, got %d", 3 + 14 * a.depth + 6,
                 a.n_weights);
  ULLAVA_REQUIRE(a.dtype == DT_BF16 || a.dtype == DT_F16, "sam_encoder: 16-bit dtypes only");
  ULLAVA_REQUIRE(a.img % a.patch == 0 && a.embed_dim % a.heads == 0 && a.embed_dim % 8 == 0 && a.out_chans % 32 == 0,
                 "sam_encoder: bad geometry");
  ULLAVA_REQUIRE(a.depth <= 64, "sam_encoder: depth > 64");
  const int B = a.batch, g = a.img / a.patch, N = g * g, D = a.embed_dim, hd = D / a.heads, C = a.out_chans;
  ULLAVA_REQUIRE(N % 32 == 0 && g <= 64, "sam_encoder: token grid must be <= 64x64 with g*g %% 32 == 0");
  if (B == 0) return OK;
  const int ws = a.window, nw = (g + ws - 1) / ws, wtok = ws * ws;
  const int rows = B * N, wrows = B * nw * nw * wtok;
  const bool any_window = ws > 0 && (a.global_mask != ~0ull);
  ULLAVA_REQUIRE(!any_window || (a.win_rows && a.unwin_rows), "sam_encoder: window row maps missing");
  ULLAVA_REQUIRE(a.scratch_bytes >= sam_encoder_scratch(B, a.img, a.patch, D, ws, C), "sam_encoder: scratch too small");
  const int kp = 3 * a.patch * a.patch;
  ULLAVA_REQUIRE(kp % 8 == 0, "sam_encoder: 3*patch*patch must be a multiple of 8");
  const int blk0 = a.block_begin, blk1 = a.block_end <= 0 ? a.depth : a.block_end;
  ULLAVA_REQUIRE(blk0 >= 0 && blk0 <= blk1 && blk1 <= a.depth, "sam_encoder: block range [%d, %d) outside [0, %d]", blk0,
                 blk1, a.depth);

  uint8_t* base = static_cast<uint8_t*>(a.scratch);
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = base + off; off += align_up(bytes); return p; };
  const size_t big = std::max(rows, wrows);
  void* x = take(static_cast<size_t>(rows) * D * 2);
  void* xn = take(static_cast<size_t>(rows) * D * 2);
  void* xw = take(static_cast<size_t>(wrows) * D * 2);
  void* qkv = take(big * 3 * D * 2);
  void* att = take(static_cast<size_t>(rows) * D * 2);
  void* act = take(std::max(static_cast<size_t>(rows) * 4 * D, static_cast<size_t>(rows) * kp) * 2);
  void* n1 = take(static_cast<size_t>(rows) * C * 2);
  void* n2 = take(static_cast<size_t>(rows) * C * 2);
  void* col3 = take(static_cast<size_t>(rows) * 9 * C * 2);
  const void* const* W = a.weights;
  const int dt = a.dtype;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));

  // ---- patch embedding (+bias, +pos_embed per image) ----
  if (blk0 == 0) {
    RUN(vit_im2col_run(ctx, a.pixels, act, B, a.img, a.patch, kp, dt, s));
    for (int b = 0; b < B; ++b) {
      const uint16_t* A = static_cast<const uint16_t*>(act) + static_cast<size_t>(b) * N * kp;
      uint16_t* Dst = static_cast<uint16_t*>(x) + static_cast<size_t>(b) * N * D;
      RUN(gemm(ctx, s, dt, A, kp, W[0], kp, Dst, D, N, D, kp, W[1], EPI_NONE, W[2], D));
    }
    // pad rows of the window layout stay zero from here on (only valid rows are ever rewritten)
    if (any_window) RUN(check_cuda(cudaMemsetAsync(xw, 0, static_cast<size_t>(wrows) * D * 2, s), "zero window pads"));
  }

  for (int l = blk0; l < blk1; ++l) {
    const void* const* L = W + 3 + 14 * l;
    const bool global = ws <= 0 || ((a.global_mask >> l) & 1ull);
    AttnArgs at{};
    const uint16_t* q16 = static_cast<const uint16_t*>(qkv);
    at.q = q16; at.k = q16 + D; at.v = q16 + 2 * D; at.o = att;
    at.q_rs = at.k_rs = at.v_rs = 3 * D;
    at.q_hs = at.k_hs = at.v_hs = hd;
    at.o_rs = D; at.o_hs = hd;
    at.heads = a.heads; at.head_dim = hd; at.scale = scale; at.dtype = dt;
    if (global) {
      RUN(layernorm_run(ctx, x, D, L[0], L[1], xn, D, rows, D, a.eps, EPI_NONE, dt, s));
      RUN(gemm(ctx, s, dt, xn, D, L[2], D, qkv, 3 * D, rows, 3 * D, D, L[3], EPI_NONE, nullptr, 0));
      at.batch = B; at.seq_q = at.seq_k = N;
      at.q_bs = at.k_bs = at.v_bs = static_cast<int64_t>(N) * 3 * D;
      at.o_bs = static_cast<int64_t>(N) * D;
      RUN(attention_relpos_run(ctx, at, L[4], L[5], g, nullptr, s));
    } else {
      RUN(layernorm_run(ctx, x, D, L[0], L[1], xw, D, rows, D, a.eps, EPI_NONE, dt, s, a.win_rows));
      RUN(gemm(ctx, s, dt, xw, D, L[2], D, qkv, 3 * D, wrows, 3 * D, D, L[3], EPI_NONE, nullptr, 0));
      at.batch = B * nw * nw; at.seq_q = at.seq_k = wtok;
      at.q_bs = at.k_bs = at.v_bs = static_cast<int64_t>(wtok) * 3 * D;
      at.o_bs = 0;
      RUN(attention_relpos_run(ctx, at, L[4], L[5], ws, a.unwin_rows, s));
    }
    RUN(gemm(ctx, s, dt, att, D, L[6], D, x, D, rows, D, D, L[7], EPI_NONE, x, D));
    RUN(layernorm_run(ctx, x, D, L[8], L[9], xn, D, rows, D, a.eps, EPI_NONE, dt, s));
    RUN(gemm(ctx, s, dt, xn, D, L[10], D, act, 4 * D, rows, 4 * D, D, L[11], EPI_GELU, nullptr, 0));
    RUN(gemm(ctx, s, dt, act, 4 * D, L[12], 4 * D, x, D, rows, D, 4 * D, L[13], EPI_NONE, x, D));
  }

  if (blk1 < a.depth) return OK;   // the remaining blocks and the neck come with a later call

  // ---- neck ----
  const void* const* Nk = W + 3 + 14 * a.depth;
  RUN(gemm(ctx, s, dt, x, D, Nk[0], D, n1, C, rows, C, D, nullptr, EPI_NONE, nullptr, 0));
  RUN(layernorm_run(ctx, n1, C, Nk[1], Nk[2], n1, C, rows, C, 1e-6f, EPI_NONE, dt, s));
  {
    ProfScope _ps(ctx, s, ULLAVA_PROF_GLUE, 0.0, 2.0 * rows * 10.0 * C);
    const int64_t total_vec = static_cast<int64_t>(rows) * 9 * (C / 8);
    const int grid = static_cast<int>(std::min<int64_t>((total_vec + 255) / 256, 148 * 16));
    im2col3x3_kernel<<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(n1), static_cast<uint16_t*>(col3), g, C, total_vec);
    ctx->launches++;
    RUN(check_cuda(cudaGetLastError(), "im2col3x3 launch"));
  }
  RUN(gemm(ctx, s, dt, col3, 9 * C, Nk[3], 9 * C, n2, C, rows, C, 9 * C, nullptr, EPI_NONE, nullptr, 0));
  RUN(layernorm_run(ctx, n2, C, Nk[4], Nk[5], n2, C, rows, C, 1e-6f, EPI_NONE, dt, s));
  {
    ProfScope _ps(ctx, s, ULLAVA_PROF_GLUE, 0.0, 4.0 * rows * C);
    dim3 grid(N / 32, C / 32, B);
    tokens_to_nchw_kernel<<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(n2), static_cast<uint16_t*>(a.out), N, C);
    ctx->launches++;
    RUN(check_cuda(cudaGetLastError(), "tokens_to_nchw launch"));
  }
  return OK;
}

}  // namespace ullava
