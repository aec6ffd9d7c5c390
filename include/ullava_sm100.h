/*
 * ullava_sm100.h -- C ABI of libullava_sm100.so, the B200 (sm_100a) implementation of the
 * u-LLaVA data-parallel inference hot path.
 *
 * The reference (OPPOMKLab/u-LLaVA) has no FFI: its boundary is the Python nn.Module API of
 * models/ullava_core.py, models/ullava.py and models/segment_anything.
 * The host-side mirror of that API lives in u-llava_b200/models/ and calls ONLY the entry
 * points declared here (through ctypes, see u-llava_b200/native.py and INTEGRATION.md).
 * Each entry point cites the reference code whose arithmetic it replaces.
 *
 * Conventions
 *   - plain C: raw device pointers, integer shapes, a cudaStream_t passed as void*;
 *     no torch / C++ types cross the boundary, no exception crosses it;
 *   - every function returns 0 on success and a negative ullava_status otherwise; the message
 *     is available from ullava_last_error() (thread local);
 *   - the library never allocates or frees caller-visible memory: outputs, KV caches and the
 *     scratch workspace are allocated by the caller (PyTorch caching allocator) and passed in;
 *   - all kernels are enqueued on the given stream and are CUDA-graph capturable
 *     (no synchronisation, no allocation inside);
 *   - "dtype" is the 16-bit storage type of activations and weights: ULLAVA_BF16 or ULLAVA_F16;
 *     accumulation, softmax, LayerNorm/RMSNorm statistics and RoPE are fp32.
 */
#ifndef ULLAVA_SM100_H_
#define ULLAVA_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ULLAVA_ABI_VERSION 4

#if defined(__GNUC__)
#define ULLAVA_API __attribute__((visibility("default")))
#else
#define ULLAVA_API
#endif

typedef struct ullava_ctx ullava_ctx;

enum ullava_status {
  ULLAVA_OK = 0,
  ULLAVA_ERR_BAD_ARG = -1,
  ULLAVA_ERR_UNSUPPORTED = -2,
  ULLAVA_ERR_CUDA = -3,
  ULLAVA_ERR_ARCH = -4,
  ULLAVA_ERR_WORKSPACE = -5
};

enum ullava_dtype { ULLAVA_BF16 = 0, ULLAVA_F16 = 1, ULLAVA_F32 = 2 };

/* GEMM epilogues (applied as  D = act(A*B^T + bias) + residual) */
enum ullava_epilogue {
  ULLAVA_EPI_NONE = 0,
  ULLAVA_EPI_RELU = 1,       /* models/ullava.py:86-101, segment_anything/modeling/transformer.py (MLPBlock, ReLU) */
  ULLAVA_EPI_GELU = 2,       /* erf GELU: models/ullava_core.py:121-126 (mlp2x projector) */
  ULLAVA_EPI_QUICK_GELU = 3, /* x*sigmoid(1.702x): CLIP MLP, hf:models/clip/modeling_clip.py:339-351 */
  ULLAVA_EPI_SILU_MUL = 4    /* silu(gate)*up on gate/up rows interleaved in blocks of 16:
                                LLaMA MLP, hf:models/llama/modeling_llama.py:171-184 */
};

/* ---- life cycle ---------------------------------------------------------------------- */
ULLAVA_API int ullava_abi_version(void);
ULLAVA_API const char* ullava_last_error(void);
/* Creates a context on `device` (compute capability 10.x).  A context carries per-launch-sequence state (scratch
 * workspace with the stream-K arrival counters, the next-weight prefetch hint, launch counters): it serves ONE stream
 * at a time.  Stages that run concurrently on two streams (see ullava_partition below) use one context each. */
ULLAVA_API int ullava_create(int device, ullava_ctx** out);
ULLAVA_API int ullava_destroy(ullava_ctx* ctx);
/* Registers caller-owned scratch memory (split-K partials, attention scratch). 256 B aligned. */
ULLAVA_API int ullava_set_workspace(ullava_ctx* ctx, void* ptr, size_t bytes);
/* Number of kernels this context has enqueued so far (bench.py's gpu_launches). */
ULLAVA_API int64_t ullava_launch_count(ullava_ctx* ctx);

/* ---- SM partition (csrc/partition.cu) ------------------------------------------------------------------------
 * Two streams whose kernels run on disjoint SM sets of one device (CUDA green contexts): lane A gets `sms_a` SMs
 * (rounded up to the architecture's granularity, 8 on sm_100), lane B the rest.  priority_*: CUDA stream priorities
 * (0 = default, negative = higher).  UllavaForCausalLM.evaluate (reference models/ullava.py:335-434, which runs
 * generate and get_visual_embs back to back) uses it to run the HBM-bound decode steps on lane A while the
 * tensor-bound SAM ViT-H encoder runs on lane B.  ullava_set_sm_limit sizes a context's persistent grids and
 * stream-K splits for `sms` SMs (0 = the whole device) -- set it to the lane's SM count on the lane's context. */
typedef struct ullava_partition ullava_partition;
ULLAVA_API int ullava_partition_create(int device, int32_t sms_a, int32_t priority_a, int32_t priority_b,
                                       ullava_partition** out);
ULLAVA_API int ullava_partition_info(const ullava_partition* p, int32_t* sms_a, int32_t* sms_b, void** stream_a,
                                     void** stream_b);
ULLAVA_API int ullava_partition_destroy(ullava_partition* p);
ULLAVA_API int ullava_set_sm_limit(ullava_ctx* ctx, int32_t sms);
ULLAVA_API int ullava_sm_count(ullava_ctx* ctx);

/* CUDA-event profiler by kernel class (bench.py's live roofline measurement).  Between begin and end every
 * kernel-launching entry point brackets its launches with events on the caller's stream; end synchronises
 * and fills out[ULLAVA_PROF_CLASSES][4] = {milliseconds, algorithmic FLOPs, algorithmic bytes, launches}.
 * Not usable during CUDA-graph capture. */
enum ullava_prof_class {
  ULLAVA_PROF_GEMM_TENSOR = 0,  /* large-M tcgen05 GEMM (tensor-pipe bound) */
  ULLAVA_PROF_GEMM_STREAM = 1,  /* small-M swap-AB weight-streaming GEMM + split-K reduce (HBM bound) */
  ULLAVA_PROF_ATTN_PREFILL = 2,
  ULLAVA_PROF_ATTN_DECODE = 3,
  ULLAVA_PROF_NORM = 4,
  ULLAVA_PROF_GLUE = 5,         /* im2col, gather/splice/copy, RoPE + KV scatter, argmax */
  ULLAVA_PROF_SAM = 6           /* SAM decoder-specific kernels (token transposes, hyper-mask, post-process) */
};
#define ULLAVA_PROF_CLASSES 8
ULLAVA_API int ullava_profile_begin(ullava_ctx* ctx);
ULLAVA_API int ullava_profile_end(ullava_ctx* ctx, double* out);

/* ---- GEMM: D[M,N] = act(A[M,K] * B[N,K]^T + bias[N]) + residual[M,N] ------------------
 * Replaces torch.nn.Linear at every call site of the path (CLIP / LLaMA / projector / lm_head /
 * seg & det projectors).  A, B: 16-bit row-major, lda/ldb in elements (multiples of 8).
 * D: 16-bit (out_f32 = 0) or fp32 (out_f32 = 1).  With ULLAVA_EPI_SILU_MUL, D has N/2 columns. */
typedef struct ullava_gemm_args {
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* D; int64_t ldd;
  const void* bias;              /* [N] 16-bit or NULL */
  const void* residual; int64_t ldr; /* 16-bit, layout of D, or NULL */
  int32_t M, N, K;
  int32_t dtype;                 /* ullava_dtype of A, B, bias, residual */
  int32_t out_f32;
  int32_t epilogue;              /* ullava_epilogue */
  int32_t force_bn;              /* 0 = auto; else tile N in {16,32,64,128,256} (tests/tuning) */
  int32_t force_splits;          /* 0 = auto; else split-K factor */
  int32_t no_swap;               /* 1 = never use the small-M swap-AB path */
} ullava_gemm_args;
ULLAVA_API int ullava_gemm(ullava_ctx* ctx, const ullava_gemm_args* args, void* stream);

/* Products with M <= 32 (the decode step's nn.Linear calls) run a weight-streaming kernel that prefetches B (the
 * weight matrix) before it synchronises with the preceding kernel of the stream (programmatic dependent launch).
 * B must therefore not be produced by the immediately preceding kernel when that kernel is one of this library's
 * norm / decode-attention / weight-streaming kernels -- true for every nn.Linear weight.  ullava_set_pdl(ctx, 0)
 * turns the early launch off (plain stream order). */
ULLAVA_API int ullava_set_pdl(ullava_ctx* ctx, int32_t enabled);

/* Next-weight hint for a chain of M <= 32 products (the decode step is nn.Linear after nn.Linear,
 * hf:models/llama/modeling_llama.py:171-333): the NEXT ullava_gemm call on this context that takes the
 * weight-streaming path pulls the head of `next_weight` [n, k] (row stride ldb elements) into L2 while its own split
 * reduction and epilogue run, so HBM stays busy across the kernel boundary.  A pure hint: results never depend on it;
 * it is dropped by the next ullava_gemm call whatever path that takes.  ullava_llama_forward / _decode_step set it
 * themselves.  ullava_set_weight_prefetch: 16 KB tiles per SM pulled ahead (0 = off, default 12). */
ULLAVA_API int ullava_gemm_next_weight(ullava_ctx* ctx, const void* next_weight, int32_t n, int32_t k, int64_t ldb);
ULLAVA_API int ullava_set_weight_prefetch(ullava_ctx* ctx, int32_t tiles_per_sm);

/* ---- normalisation ---------------------------------------------------------------------
 * LayerNorm over the last dim (fp32 statistics): CLIP pre_layrnorm / layer_norm1/2
 * (hf:models/clip/modeling_clip.py:354-385,677), SAM TwoWayTransformer norms
 * (segment_anything/modeling/transformer.py:151-182), LayerNorm2d of the mask head (common.py:31-43,
 * channels-last rows).  act: ULLAVA_EPI_NONE or ULLAVA_EPI_GELU (fused activation).  y may alias x. */
ULLAVA_API int ullava_layernorm(ullava_ctx* ctx, const void* x, int64_t ldx, const void* weight, const void* bias, void* y,
                     int64_t ldy, int32_t rows, int32_t cols, float eps, int32_t act, int32_t dtype, void* stream);
/* LlamaRMSNorm (hf:models/llama/modeling_llama.py:52-69): y = w * (x * rsqrt(mean(x^2) + eps)),
 * x up-cast to fp32, normalised value rounded to the 16-bit dtype BEFORE the multiply by w. */
ULLAVA_API int ullava_rmsnorm(ullava_ctx* ctx, const void* x, int64_t ldx, const void* weight, void* y, int64_t ldy, int32_t rows,
                   int32_t cols, float eps, int32_t dtype, void* stream);

/* ---- attention -------------------------------------------------------------------------
 * Flash-style softmax(Q K^T * scale) V with fp32 online softmax.  Q/K/V are strided views
 * [batch, seq, heads, head_dim] given by (batch stride, row stride, head stride) in elements, so
 * the packed QKV GEMM output or a KV cache can be consumed in place.
 * Replaces CLIPAttention (hf:models/clip/modeling_clip.py:282-336, non causal, hd 64) and
 * LlamaAttention eager (hf:models/llama/modeling_llama.py:199-291, causal, hd 128).
 * causal: query i (absolute position q_pos0 + i) sees keys j <= q_pos0 + i. */
typedef struct ullava_attn_args {
  const void* q; int64_t q_bs, q_rs, q_hs;
  const void* k; int64_t k_bs, k_rs, k_hs;
  const void* v; int64_t v_bs, v_rs, v_hs;
  void* o; int64_t o_bs, o_rs, o_hs;
  int32_t batch, heads, seq_q, seq_k, head_dim;
  int32_t causal, q_pos0;
  float scale;
  int32_t dtype;
} ullava_attn_args;
ULLAVA_API int ullava_attention(ullava_ctx* ctx, const ullava_attn_args* args, void* stream);

/* Attention of the SAM ViT image encoder with decomposed relative-position bias
 * (segment_anything/modeling/image_encoder.py:196-260, add_decomposed_rel_pos :355-392):
 *   softmax(scale * q k^T + q . rel_h[qh - kh + S - 1] + q . rel_w[qw - kw + S - 1]) v
 * over a grid_side x grid_side token grid (seq_q == seq_k == grid_side^2), rel_h / rel_w: [2*grid_side-1, head_dim].
 * o_row_map (optional, int32 [batch*seq_q]): output row of each query row in units of o_rs (window_unpartition;
 * -1 drops the row), in which case o_bs is ignored. */
ULLAVA_API int ullava_attention_relpos(ullava_ctx* ctx, const ullava_attn_args* args, const void* rel_h, const void* rel_w,
                                       int32_t grid_side, const int32_t* o_row_map, void* stream);

/* Kernel selection for ullava_attention / ullava_attention_relpos and the model-level entry points:
 * 0 (default) = per shape (tcgen05/TMEM flash attention for head_dim 64 / 80 / 128 with >= 32 rows and keys and for
 * the 64 x 64 rel-pos grid; warp-level mma.sync kernels otherwise), 1 = warp-level kernels only, 2 = tcgen05/TMEM
 * wherever it is compiled.  1 and 2 exist for A/B measurements and tests; all of it is sm_100a code. */
ULLAVA_API int ullava_set_attention_impl(ullava_ctx* ctx, int32_t impl);

/* Single-query (decode) attention against a KV cache [batch, heads, max_seq, head_dim]:
 * LlamaAttention with past_key_values, one new token per sample.  ctx_len keys are attended. */
ULLAVA_API int ullava_attention_decode(ullava_ctx* ctx, const void* q, int64_t q_bs, const void* k_cache, const void* v_cache,
                            int64_t cache_bs, int64_t cache_hs, void* o, int64_t o_bs, int32_t batch, int32_t heads,
                            int32_t head_dim, int32_t ctx_len, float scale, int32_t dtype, void* stream);

/* Decode step, fused: RoPE of the new token's q and k, the KV-cache write of its k and v (row ctx_len - 1) and the
 * single-query attention over rows 0 .. ctx_len - 1, in one kernel (what ullava_rope_kvcache followed by
 * ullava_attention_decode compute for seq == 1; ullava_llama_forward / _decode_step use this form).
 * qkv: packed [B, 3 * heads * head_dim] rows (row stride ld_qkv); rope_cos / rope_sin fp32 [max_seq, head_dim / 2];
 * pos_dev (optional): device int32, ctx_len = *pos_dev + 1 (CUDA-graph replay; max_seq then bounds the score buffer);
 * pos_offset (optional): device int32 [B], added to ctx_len per sample (see ullava_llama_args.pos_offset). */
ULLAVA_API int ullava_attention_decode_rope(ullava_ctx* ctx, const void* qkv, int64_t ld_qkv, void* k_cache, void* v_cache,
                                 int64_t cache_bs, int64_t cache_hs, void* o, int64_t o_bs, int32_t batch, int32_t heads,
                                 int32_t head_dim, int32_t ctx_len, const int32_t* pos_dev, int32_t max_seq,
                                 const float* rope_cos, const float* rope_sin, const int32_t* pos_offset, float scale,
                                 int32_t dtype, void* stream);

/* RoPE (rotate-half, hf:models/llama/modeling_llama.py:74-168) applied to the q and k thirds of a
 * packed QKV buffer [rows, 3*heads*head_dim]; rotated k and the untouched v are scattered into the
 * KV cache [batch, heads, max_seq, head_dim] at position pos0 + (row % seq); rotated q is written
 * back in place.  cos_table / sin_table: fp32 [max_pos, head_dim/2] (LlamaRotaryEmbedding, fp32). */
ULLAVA_API int ullava_rope_kvcache(ullava_ctx* ctx, void* qkv, int64_t ld_qkv, void* k_cache, void* v_cache, int64_t cache_bs,
                        int64_t cache_hs, int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, int32_t pos0,
                        const float* cos_table, const float* sin_table, int32_t dtype, void* stream);

/* ---- glue kernels ------------------------------------------------------------------------ */
/* CLIP patchify as im2col (hf:models/clip/modeling_clip.py:148-154,209-210): pixels [B,3,H,W] ->
 * patches [B*gh*gw, k_pad] with column order (c, ky, kx) matching Conv2d.weight.view(out, -1);
 * columns >= 3*patch*patch are zero. */
ULLAVA_API int ullava_vit_im2col(ullava_ctx* ctx, const void* pixels, void* out, int32_t batch, int32_t img, int32_t patch,
                      int32_t k_pad, int32_t dtype, void* stream);
/* CLIP embeddings (hf:...modeling_clip.py:212-218): out[b,0]=cls+pos[0]; out[b,1+p]=patch[b,p]+pos[1+p]. */
ULLAVA_API int ullava_vit_assemble(ullava_ctx* ctx, const void* patch_embeds, const void* cls, const void* pos, void* out,
                        int32_t batch, int32_t n_patches, int32_t dim, int32_t dtype, void* stream);
/* Token-embedding gather (models/ullava_core.py:191): out[r] = table[ids[r]]. */
ULLAVA_API int ullava_embed_gather(ullava_ctx* ctx, const int64_t* ids, const void* table, void* out, int32_t rows, int32_t dim,
                        int32_t vocab, int32_t dtype, void* stream);
/* Image-feature splice (models/ullava_core.py:222-246): for each sample b overwrite rows
 * [start[b]+1, start[b]+1+n_patch) of embeds [B,L,dim] with feats [B,n_patch,dim]; start[b] < 0 = skip. */
ULLAVA_API int ullava_splice_rows(ullava_ctx* ctx, void* embeds, const void* feats, const int32_t* start, int32_t batch,
                       int32_t seq, int32_t n_patch, int32_t dim, int32_t dtype, void* stream);
/* Strided row copy dst[r] = src[r] for 16-bit matrices (drop-CLS, hidden-state accumulation). */
ULLAVA_API int ullava_copy_rows(ullava_ctx* ctx, const void* src, int64_t src_bs, int64_t src_rs, void* dst, int64_t dst_bs,
                     int64_t dst_rs, int32_t batch, int32_t rows, int32_t cols, int32_t dtype, void* stream);
/* Greedy token: out[r] = argmax_c logits[r, c] over fp32 logits (first index on ties, like torch.argmax). */
/* encode_video pooling (models/ullava_core.py:160-180): feats [batch, frames, patches, dim] 16-bit ->
 * out [batch, frames + patches, dim]: temporal means (over patches) first, then spatial means (over frames). */
ULLAVA_API int ullava_video_pool(ullava_ctx* ctx, const void* feats, void* out, int32_t batch, int32_t frames, int32_t patches,
                      int32_t dim, int32_t dtype, void* stream);
ULLAVA_API int ullava_argmax(ullava_ctx* ctx, const float* logits, int64_t ld, int64_t* out, int32_t rows, int32_t cols,
                  void* stream);

/* ---- SAM mask decoder ----------------------------------------------------------------------
 * Everything after the image encoder for n prompts of one image:
 * PromptEncoder(text_embeds) (segment_anything/modeling/prompt_encoder.py:140-186),
 * MaskDecoder.predict_masks + TwoWayTransformer (mask_decoder.py:116-164, transformer.py:62-242),
 * mask slice 0 (multimask_output=False, mask_decoder.py:108-111).
 * Weights are passed as a table of device pointers in the order documented in
 * u-llava_b200/models/segment_anything/native_decoder.py (SAM_DECODER_WEIGHT_ORDER). */
typedef struct ullava_sam_decoder_args {
  const void* const* weights;   /* host array of device pointers, 16-bit, order fixed (ULLAVA_SAM_N_WEIGHTS) */
  int32_t n_weights;
  const void* image_embeddings; /* [n_images, 256, 64, 64] 16-bit (SAM image encoder output, NCHW) */
  const int32_t* prompt_image;  /* device int32 [n_prompts]: image index of every prompt */
  const void* image_pe;         /* [256, 64, 64] 16-bit dense positional encoding (get_dense_pe) */
  const void* text_embeds;      /* [n_prompts, 256] 16-bit ([SEG] embeddings after seg_projector) */
  int32_t n_prompts;
  void* low_res_masks;          /* out [n_prompts, 4, 256, 256] 16-bit (all mask tokens; caller slices 0:1) */
  void* iou_pred;               /* out [n_prompts, 4] 16-bit or NULL */
  void* scratch; size_t scratch_bytes; /* caller-owned, >= ullava_sam_mask_decoder_scratch_bytes(n_prompts) */
  int32_t dtype;
} ullava_sam_decoder_args;
#define ULLAVA_SAM_N_WEIGHTS 121
ULLAVA_API int ullava_sam_mask_decoder(ullava_ctx* ctx, const ullava_sam_decoder_args* args, void* stream);
ULLAVA_API size_t ullava_sam_mask_decoder_scratch_bytes(int32_t n_prompts);

/* Sam.postprocess_masks (segment_anything/modeling/sam.py:137-172): fp32 bilinear low_res->img_size
 * (align_corners=False), crop [:in_h,:in_w], bilinear -> (out_h,out_w); fused, no img_size^2 buffer.
 * masks: n masks of [low_res,low_res] 16-bit, mask_stride elements apart; out: [n,out_h,out_w] fp32.
 * packed_bits (optional): [n, ceil(out_h*out_w/32)] uint32, bit = (logit > 0), for the eval gather. */
ULLAVA_API int ullava_sam_postprocess(ullava_ctx* ctx, const void* masks, int64_t mask_stride, float* out, uint32_t* packed_bits,
                           int32_t n, int32_t low_res, int32_t img_size, int32_t in_h, int32_t in_w, int32_t out_h,
                           int32_t out_w, int32_t dtype, void* stream);

/* ---- fused model-level entry points (layer loops run in C++, one call per stage) ----------- */
/* CLIP ViT (encode_image, models/ullava_core.py:146-158): pixel_values -> hidden_states[n_layers_used]
 * without CLS.  Weight table order: see u-llava_b200/models/native_layout.py (VIT_WEIGHT_ORDER). */
typedef struct ullava_vit_args {
  const void* const* weights; int32_t n_weights;
  const void* pixels;      /* [B,3,img,img] 16-bit */
  void* out;               /* [B, n_patches, hidden] 16-bit (CLS dropped) */
  void* scratch; size_t scratch_bytes;
  int32_t batch, img, patch, hidden, heads, ffn, layers_used, k_pad;
  int32_t act;             /* ullava_epilogue of the MLP: QUICK_GELU (openai CLIP) or GELU */
  float eps;
  int32_t dtype;
} ullava_vit_args;
ULLAVA_API int ullava_vit_forward(ullava_ctx* ctx, const ullava_vit_args* args, void* stream);
ULLAVA_API size_t ullava_vit_scratch_bytes(int32_t batch, int32_t img, int32_t patch, int32_t hidden, int32_t ffn, int32_t k_pad);

/* SAM ViT image encoder (ImageEncoderViT.forward, segment_anything/modeling/image_encoder.py:110-426; called per
 * image by UllavaForCausalLM.get_visual_embs, models/ullava.py:139-150) on a batch: pixels -> [B, out_chans, g, g].
 * Window blocks attend inside zero-padded window x window tiles (pads are real tokens holding the qkv bias, as in
 * the reference); blocks whose bit is set in global_mask attend over the whole g x g grid.  Both use the decomposed
 * relative-position bias.  win_rows / unwin_rows are the window_partition / window_unpartition row maps
 * (u-llava_b200/models/segment_anything/modeling/image_encoder.py builds them).  Weight order: csrc/sam_encoder.cu. */
typedef struct ullava_sam_encoder_args {
  const void* const* weights; int32_t n_weights;   /* 3 + 14*depth + 6 */
  const void* pixels;       /* [B,3,img,img] 16-bit */
  void* out;                /* [B,out_chans,g,g] 16-bit, NCHW */
  void* scratch; size_t scratch_bytes;
  const int32_t* win_rows;   /* device [B*g*g]: token row -> row in the padded window layout */
  const int32_t* unwin_rows; /* device [B*nw*nw*window*window]: window-layout row -> token row, -1 for pads */
  int32_t batch, img, patch, embed_dim, depth, heads, window, out_chans;
  uint64_t global_mask;     /* bit l set: block l uses global attention */
  float eps;                /* LayerNorm eps of the blocks (1e-6) */
  int32_t dtype;
  /* Block range [block_begin, block_end) of this call (block_end <= 0: depth).  The patch embedding runs when
   * block_begin == 0, the neck + NCHW store when the range ends at `depth`; the token state lives in `scratch`
   * between two calls (same scratch, same batch), so the encoder can be cut into a part that runs on an SM-partition
   * lane beside the decode steps and a remainder that runs on the whole machine afterwards. */
  int32_t block_begin, block_end;
} ullava_sam_encoder_args;
ULLAVA_API int ullava_sam_encoder_forward(ullava_ctx* ctx, const ullava_sam_encoder_args* args, void* stream);
ULLAVA_API size_t ullava_sam_encoder_scratch_bytes(int32_t batch, int32_t img, int32_t patch, int32_t embed_dim, int32_t window,
                                        int32_t out_chans);

/* LLaMA decoder stack (LlamaModel.forward, hf:models/llama/modeling_llama.py:355-424) on a batch of
 * equal-length sequences: hidden [B*S, H] (in place) at absolute positions pos0..pos0+S-1, appending
 * K/V to the caches.  S == 1 is the decode step (swap-AB GEMMs + single-query attention).
 * After the call `hidden` holds the output of the LAST layer (pre final norm) and, if final_out is
 * not NULL, final_out holds final RMSNorm(hidden).  all_hidden (optional) receives the input of every
 * layer ([layers][B*S][H]) for output_hidden_states=True. */
typedef struct ullava_llama_args {
  const void* const* weights; int32_t n_weights;   /* per layer: ln1, wqkv, wo, ln2, wgu(packed), wdown; then final norm */
  void* hidden;            /* [B*S, H] 16-bit, in/out */
  void* final_out;         /* [B*S, H] 16-bit or NULL */
  void* all_hidden;        /* [layers, B*S, H] 16-bit or NULL */
  void* k_cache; void* v_cache;  /* [layers, B, heads, max_seq, hd] 16-bit */
  void* scratch; size_t scratch_bytes;
  int32_t batch, seq, pos0, max_seq;
  int32_t layers, hidden_size, heads, head_dim, ffn;
  float eps;
  const float* rope_cos; const float* rope_sin; /* fp32 [max_seq, head_dim/2] */
  int32_t dtype;
  /* decode step only (ullava_llama_decode_step, batch <= 32), optional: device buffer of ullava_llama_chain_bytes(...)
   * bytes holding the decode-layer chain program that ullava_llama_chain_prepare built for EXACTLY these arguments
   * (pointers, shapes) and this context.  With it the step runs 2 kernels per layer (single-query attention + one
   * persistent chain kernel: o_proj, RMSNorm, gate/up, down, RMSNorm, next q/k/v, csrc/gemm_chain_sm100.cu) instead of 7;
   * results are bit-identical. */
  void* chain_program; size_t chain_bytes;
  /* decode only (seq == 1), optional: device int32 [B]; sample b's token sits at KV / RoPE position
   * pos + pos_offset[b] (<= 0).  Lets one batch hold right-padded prompts of different lengths: the prefill runs on
   * the padded [B, P] block (causal attention keeps valid positions exact), then every sample continues right after
   * ITS last valid token (the reference generates one prompt at a time, models/ullava.py:350-362). */
  const int32_t* pos_offset;
} ullava_llama_args;
ULLAVA_API int ullava_llama_forward(ullava_ctx* ctx, const ullava_llama_args* args, void* stream);
ULLAVA_API size_t ullava_llama_scratch_bytes(int32_t rows, int32_t hidden_size, int32_t ffn);

/* One greedy decode step, CUDA-graph replayable: the position of the token being processed is read from
 * device memory (*pos_dev = number of tokens already in the KV cache) and incremented by the last kernel.
 *   hidden <- embed_table[cur_ids]; decoder stack (seq 1, llama.pos0 ignored); final norm; logits = lm_head(final);
 *   next = argmax(logits) (pad once finished; eos sets finished); cur_ids <- next; seqs[b][pos+1] <- next;
 *   hid_buf[b][pos] <- final (post-norm hidden state, what UllavaForCausalLM.evaluate gathers for [SEG]).
 * One iteration of the generate loop of models/ullava.py:350-362 / models/ullava_core.py:357-395. */
typedef struct ullava_decode_args {
  ullava_llama_args llama;      /* seq == 1, final_out != NULL */
  int32_t* pos_dev;             /* device int32 scalar */
  const void* embed_table; int32_t vocab;   /* [vocab, H] 16-bit */
  const void* lm_head;          /* [vocab, H] 16-bit */
  int64_t* cur_ids;             /* device [B]: in = token to process, out = next token */
  float* logits;                /* device [B, vocab] fp32 */
  int64_t* seqs; int64_t seqs_ld;           /* device [B, seqs_ld] or NULL */
  void* hid_buf; int64_t hid_bs;            /* device [B, hid_bs / H, H] 16-bit or NULL (hid_bs in elements) */
  uint8_t* finished;            /* device [B] or NULL */
  int32_t eos_id, pad_id;       /* eos_id < 0: never stop */
  /* sampling (do_sample = temperature > 0, models/ullava.py:350-362): uniforms != NULL selects ullava_sample_step
   * instead of the argmax; uniforms[pos * uniforms_ld + b] in [0, 1) is the draw of row b at position pos */
  const float* uniforms; int64_t uniforms_ld;
  float temperature, top_p;
  int32_t top_k;                /* 0 = off; HF's GenerationConfig default (50) is applied by the Python caller */
} ullava_decode_args;
ULLAVA_API int ullava_llama_decode_step(ullava_ctx* ctx, const ullava_decode_args* args, void* stream);
/* Builds the decode-layer chain program into args->llama.chain_program (synchronous: host-side TMA descriptor encoding +
 * one cudaMemcpy; call it once per decode session, outside any stream capture). */
/* Debug aid (tools/bench_chain.py): buf != NULL makes every chain kernel launched through ctx write globaltimer
 * stamps [grid][steps][8] (W issued, X ready, first MMA, last MMA, step done, W first, norm begin, norm done) to buf. */
ULLAVA_API int ullava_debug_chain_trace(ullava_ctx* ctx, void* buf);
/* Debug aid (tools/bench_attn.py TRACE=1): buf != NULL makes CTA (0, 0, 0) of every fmha_tcgen05 launch write clock64
 * stamps [tile < 64][48] (slot names in tools/bench_attn.py). */
ULLAVA_API int ullava_debug_fmha_trace(ullava_ctx* ctx, void* buf);
ULLAVA_API size_t ullava_llama_chain_bytes(int32_t layers, int32_t hidden, int32_t ffn, int32_t vocab);
ULLAVA_API int ullava_llama_chain_prepare(ullava_ctx* ctx, const ullava_decode_args* args);
/* The bookkeeping tail of a step on its own (used once after the prefill, with *pos_dev = P - 1):
 * argmax + eos/pad handling, cur_ids / seqs[b][pos+1] / hid_buf[b][pos] updates, ++*pos_dev. */
ULLAVA_API int ullava_greedy_step(ullava_ctx* ctx, const float* logits, int64_t ld, int32_t rows, int32_t cols, int64_t* cur_ids,
                       int64_t* seqs, int64_t seqs_ld, const void* final_h, void* hid_buf, int64_t hid_bs, int32_t hdim,
                       uint8_t* finished, int32_t eos_id, int32_t pad_id, int32_t* pos_dev, void* stream);

/* Sampling flavour of the same tail (temperature > 0; top_k > 0 keeps the k largest scores like TopKLogitsWarper,
 * 0 = off; top_p in [0, 1) then filters the survivors like TopPLogitsWarper, 1 = off, 0 = top-1 only):
 * scores = logits / temperature, top-k filter, nucleus filter, then ONE draw per row by inverse CDF in vocabulary order with the
 * caller's uniform number uniforms[pos * uniforms_ld + b] (pos = *pos_dev, 0 if pos_dev is NULL).  torch.multinomial's
 * random stream is not reproducible outside torch: parity is on the filtered distribution (probs_out, optional
 * [rows, cols] fp32, receives it) and on the inverse-CDF draw.  Replaces the sampling branch of
 * GenerationMixin.generate as called by models/ullava.py:350-362. */
ULLAVA_API int ullava_sample_step(ullava_ctx* ctx, const float* logits, int64_t ld, int32_t rows, int32_t cols, float temperature,
                       float top_p, int32_t top_k, const float* uniforms, int64_t uniforms_ld, int64_t* cur_ids, int64_t* seqs,
                       int64_t seqs_ld, const void* final_h, void* hid_buf, int64_t hid_bs, int32_t hdim,
                       uint8_t* finished, int32_t eos_id, int32_t pad_id, int32_t* pos_dev, float* probs_out,
                       void* stream);

/* ---- evaluation metrics on the device (evaluation/tools.py, evaluation/eval_ullava.py:41-102) -----------------
 * ullava_mask_iou_counts: intersectionAndUnionGPU (evaluation/tools.py:29-41) with K = 2, for n masks of hw pixels
 * in one launch.  pred_kind: 0 = fp32 mask logits (label = logit > 0, the callers' threshold, eval_ullava.py:65),
 * 1 = int32 labels, 2 = uint8 labels; target_kind: 1 = int32, 2 = uint8; pixels whose target == ignore_index are
 * dropped.  counts [n, 6] int32 = area_intersection[0..1], area_union[0..1], area_target[0..1] (exact integers). */
ULLAVA_API int ullava_mask_iou_counts(ullava_ctx* ctx, const void* pred, int32_t pred_kind, const void* target, int32_t target_kind,
                           int32_t n, int64_t hw, int32_t ignore_index, int32_t* counts, void* stream);
/* The meter arithmetic of validate() (evaluation/eval_ullava.py:66-86) for a batch of images: image i owns the
 * masks offsets[i] .. offsets[i+1]-1 of `counts`.  state: 8 doubles on the device = intersection sum[2], union
 * sum[2], acc_iou sum[2], images, masks (ciou = state[1] / (state[3] + 1e-10), giou = state[5] / state[7]). */
ULLAVA_API int ullava_seg_meter_update(ullava_ctx* ctx, const int32_t* counts, const int32_t* offsets, int32_t n_images,
                            double* state, void* stream);
/* bbox_iou (evaluation/tools.py:13-26): iou[i] = torchvision box_iou(pred[i] * 1000, gt[i] * 1000), xyxy, dtype
 * ULLAVA_BF16 / ULLAVA_F16 / 2 (fp32) with torchvision's rounding points.  meter (optional, 3 doubles on the device:
 * hits, boxes, scratch): hits += #{iou > 0.5}, boxes += n. */
ULLAVA_API int ullava_box_iou_diag(ullava_ctx* ctx, const void* pred, const void* gt, int32_t n, int32_t dtype, float* iou,
                        double* meter, void* stream);

/* Shifted token cross-entropy of UllavaCoreForCausalLM.forward(labels=...) (models/ullava_core.py:327-338):
 * mean over t < T - 1, labels[b, t + 1] != ignore_index, of logsumexp(logits[b, t, :]) - logits[b, t, labels[b, t+1]].
 * logits: [batch, T, cols] fp32 (logits_f32 = 1) or 16-bit `dtype`, row stride ld_row, batch stride ld_batch
 * (elements); labels int64 [batch, >= T] with row stride labels_ld.  out: device float[2] = {loss, valid tokens}
 * (loss is NaN when every label is ignored, like torch).  scratch: ullava_cross_entropy_scratch_bytes. */
ULLAVA_API size_t ullava_cross_entropy_scratch_bytes(int32_t batch, int32_t seq);
ULLAVA_API int ullava_cross_entropy(ullava_ctx* ctx, const void* logits, int32_t logits_f32, int32_t dtype, int64_t ld_row,
                         int64_t ld_batch, const int64_t* labels, int64_t labels_ld, int32_t batch, int32_t seq,
                         int32_t cols, int32_t ignore_index, float* out, void* scratch, size_t scratch_bytes,
                         void* stream);

/* ---- image preprocessing on the device (dataset/processors/clip_processor.py, dataset/tools/mask_toolbox.py) ---
 * ullava_resize_u8: PIL.Image.resize of an RGB uint8 HWC image, bit-exact with Pillow's 8-bit resampler (separable,
 * antialiased, 22-bit fixed-point taps); filter 0 = BILINEAR (ResizeLongestSide.apply_image,
 * models/segment_anything/utils/transforms.py:29-37), 1 = BICUBIC (CLIPImageProcessor.resize). */
/* Host-only (no GPU work, no context): the fixed-point tap table of one resampling pass, exactly Pillow's
 * precompute_coeffs + normalize_coeffs_8bpc.  Returns ksize (taps per output sample) or -1 on bad arguments; when
 * bounds ([out_size][2] = first input index, tap count) and taps ([out_size][ksize]) are given and `capacity` (int32
 * entries available in taps) suffices they are filled in. */
ULLAVA_API int ullava_resample_coeffs(int32_t in_size, int32_t out_size, int32_t filter, int32_t* bounds, int32_t* taps,
                           size_t capacity);
ULLAVA_API size_t ullava_resize_u8_scratch_bytes(int32_t h, int32_t w, int32_t out_h, int32_t out_w);
ULLAVA_API int ullava_resize_u8(ullava_ctx* ctx, const uint8_t* src, int32_t h, int32_t w, uint8_t* dst, int32_t out_h, int32_t out_w,
                     int32_t filter, void* scratch, size_t scratch_bytes, void* stream);
/* CLIPImageProcessor.preprocess after the resize: center crop (top, left, size) -> x * rescale (double product
 * rounded to fp32) -> (x - mean[c]) / std[c] in fp32 -> 16-bit [3, size, size].  mean / std: host float[3]. */
ULLAVA_API int ullava_clip_preprocess(ullava_ctx* ctx, const uint8_t* src, int32_t h, int32_t w, int32_t top, int32_t left,
                           int32_t size, const float* mean, const float* std, double rescale, void* out, int32_t dtype,
                           void* stream);
/* SegToolBox.preprocess (dataset/tools/mask_toolbox.py:15-25): (x - mean) / std in fp32, zero padding to
 * [3, sam_size, sam_size], 16-bit. */
ULLAVA_API int ullava_sam_preprocess(ullava_ctx* ctx, const uint8_t* src, int32_t h, int32_t w, int32_t sam_size, const float* mean,
                          const float* std, void* out, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ULLAVA_SM100_H_ */
