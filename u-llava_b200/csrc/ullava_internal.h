// Internal C++ declarations shared by the .cu translation units of libullava_sm100.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <vector>
#include "../../include/ullava_sm100.h"

struct ullava_prof_rec {
  cudaEvent_t a, b;
  int cls;
  double flops, bytes;
  int launches;
};

namespace ullava { struct RopeFuse; }

struct ullava_ctx {
  int device = 0;
  int device_sm_count = 148;  // SMs of the device
  int sm_count = 148;         // SMs this context sizes persistent grids / stream-K splits for: the device's, or the SM
                              // partition's when the context serves one lane of an ullava_partition (ullava_set_sm_limit)
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
  int64_t launches = 0;
  int pdl = 1;        // decode GEMMs are launched with programmatic stream serialization (weight prefetch under the
                      // previous kernel's tail); 0 = plain launches
  // Next-weight hint for the weight-streaming GEMM (ullava_gemm_next_weight): the M <= 32 GEMM launched next on this
  // context pulls the first tiles of THIS matrix into L2 while its own split reduction / epilogue tail runs, so HBM
  // stays busy across the kernel boundary.  Consumed (cleared) by the next ullava_gemm call.
  const void* next_w = nullptr;
  int next_n = 0, next_k = 0;
  int64_t next_ldb = 0;
  int prefetch_units = 12;  // 16 KB tiles per SM pulled into L2 (0 = off); ULLAVA_PREFETCH_UNITS overrides at create
  int gemm_pair = 1;  // large-M GEMMs on CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles); ULLAVA_GEMM_PAIR=0 turns it off
  int gemm_tma_store = 1;  // CTA-pair GEMM: 16-bit D leaves through shared memory + cp.async.bulk.tensor (ULLAVA_GEMM_TMA_STORE=0: direct stores)
  int gemm_hints = 0; // L2 eviction-priority hints on the large-M GEMM operand loads (ULLAVA_GEMM_HINTS=1): measured
                      // counter-productive on B200, see gemm_sm100.cu
  int group_m = 0;    // 0 = default rasterisation group of the large-M GEMM; ULLAVA_GROUP_M overrides at create (tuning)
  const ullava::RopeFuse* rope_fuse = nullptr;   // see RopeFuse
  void* fmha_trace = nullptr;   // debug: clock64 stamps of CTA (0,0,0) of the next fmha_tcgen05 launches
  void* chain_trace = nullptr;  // debug: globaltimer stamps of the next chain kernels (ullava_debug_chain_trace)
  int attn_impl = 0;  // 0 = pick per shape, 1 = warp-level mma.sync kernels only, 2 = tcgen05/TMEM wherever compiled
  // per-kernel-class CUDA-event profiling (ullava_profile_begin/end); off on the normal path
  bool prof_on = false;
  std::vector<ullava_prof_rec> prof;
};

namespace ullava {

using Context = ::ullava_ctx;
using GemmArgs = ::ullava_gemm_args;

// RoPE + KV-cache scatter fused into the epilogue of the prefill qkv GEMM (gemm_sm100.cu, CTA-pair kernel): set on the
// context right before that ullava_gemm-shaped call, consumed (and cleared) by gemm_run.  head_dim 128 only.
struct RopeFuse {
  const float* cos_t;   // [max_pos][64] fp32
  const float* sin_t;
  void* k_cache;        // [batch][heads][max_seq][128], strides in elements
  void* v_cache;
  int64_t cache_bs, cache_hs;
  int seq, pos0, heads;
};
constexpr int EPI_QKV_ROPE = 100;   // internal epilogue code of that GEMM
bool gemm_pair_eligible(const struct ::ullava_ctx* ctx, int M, int N);
using AttnArgs = ::ullava_attn_args;

enum Epilogue : int {
  EPI_NONE = ULLAVA_EPI_NONE,
  EPI_RELU = ULLAVA_EPI_RELU,
  EPI_GELU = ULLAVA_EPI_GELU,
  EPI_QUICK_GELU = ULLAVA_EPI_QUICK_GELU,
  EPI_SILU_MUL = ULLAVA_EPI_SILU_MUL,
};

// runtime.cu -- kernel classes of the CUDA-event profiler (ULLAVA_PROF_* in the public header)
struct ProfScope {
  Context* ctx;
  cudaStream_t stream;
  int idx;
  int64_t launches0;
  ProfScope(Context* c, cudaStream_t s, int cls, double flops, double bytes);
  ~ProfScope();
};

// gemm_sm100.cu
int gemm_run(Context* ctx, const GemmArgs& a, cudaStream_t stream);

// gemm_stream_sm100.cu -- M <= 32 weight-streaming GEMM (stream-K, in-kernel split reduction, PDL weight prefetch).
// The first kStreamCounterBytes of the context workspace hold its per-tile arrival counters (zero between launches);
// every other workspace user starts behind them.
static constexpr size_t kStreamCounterBytes = 64 * 1024;
int gemm_stream_run(Context* ctx, const GemmArgs& a, cudaStream_t stream);
size_t gemm_stream_workspace_bytes(int sm_count);

// gemm_chain_sm100.cu -- the GEMMs + RMSNorms of a decode layer as one persistent kernel (program in device memory)
size_t chain_step_bytes();
int chain_encode_step(void* host_step, int bn, const void* W, int64_t ldb, const void* X, int64_t lda, int M, int N, int K,
                      void* D, int64_t ldd, const void* residual, int64_t ldr, int ek, int out_f32, const void* norm_src,
                      const void* norm_w, void* norm_dst, int norm_cols, float norm_eps, int* counters_dev);
int gemm_chain_run(Context* ctx, const void* steps_dev, int n_steps, int M, int dtype, int* sync_dev, cudaStream_t stream);

// norm.cu
int layernorm_run(Context* ctx, const void* x, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy,
                  int rows, int cols, float eps, int act, int dtype, cudaStream_t stream,
                  const int32_t* dst_rows = nullptr);
int rmsnorm_run(Context* ctx, const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int rows, int cols,
                float eps, int dtype, cudaStream_t stream);

// attention.cu
int attention_run(Context* ctx, const AttnArgs& a, cudaStream_t stream);
int attention_relpos_run(Context* ctx, const AttnArgs& a, const void* rel_h, const void* rel_w, int S,
                         const int32_t* o_row_map, cudaStream_t stream);
int attention_decode_run(Context* ctx, const void* q, int64_t q_bs, const void* kc, const void* vc, int64_t cache_bs,
                         int64_t cache_hs, void* o, int64_t o_bs, int batch, int heads, int head_dim, int ctx_len,
                         float scale, int dtype, cudaStream_t stream, const int32_t* ctx_dev = nullptr,
                         int max_ctx = 0, const float* rope_cos = nullptr, const float* rope_sin = nullptr,
                         const int32_t* ctx_offset = nullptr);

// fmha_sm100.cu -- tcgen05/TMEM flash attention (hd 64 / 80 / 128); rel_h == nullptr: no rel-pos bias
bool fmha_supported(const AttnArgs& a, bool relpos, int S);
int fmha_run(Context* ctx, const AttnArgs& a, const void* rel_h, const void* rel_w, int S, const int32_t* o_row_map,
             cudaStream_t stream);

// fmha_window_sm100.cu -- SAM 14 x 14 windowed attention (hd 80), single pass, 2 CTAs per SM; needs [Rh; Rw] contiguous
bool fmha_window_supported(const AttnArgs& a, const void* rel_h, const void* rel_w, int S);
int fmha_window_run(Context* ctx, const AttnArgs& a, const void* rel_h, const int32_t* o_row_map, cudaStream_t stream);

// elementwise.cu
int rope_kvcache_run(Context* ctx, void* qkv, int64_t ld_qkv, void* kc, void* vc, int64_t cache_bs, int64_t cache_hs,
                     int batch, int seq, int heads, int head_dim, int pos0, const float* cos_t, const float* sin_t,
                     int dtype, cudaStream_t stream, const int32_t* pos_dev = nullptr);
int greedy_step_run(Context* ctx, const float* logits, int64_t ld, int rows, int cols, int64_t* cur_ids, int64_t* seqs,
                    int64_t seqs_ld, const void* final_h, void* hid_buf, int64_t hid_bs, int hdim, uint8_t* finished,
                    int eos_id, int pad_id, int32_t* pos_dev, cudaStream_t stream);
int vit_im2col_run(Context* ctx, const void* pixels, void* out, int batch, int img, int patch, int k_pad, int dtype,
                   cudaStream_t stream);
int vit_assemble_run(Context* ctx, const void* pe, const void* cls, const void* pos, void* out, int batch, int np,
                     int dim, int dtype, cudaStream_t stream);
int embed_gather_run(Context* ctx, const int64_t* ids, const void* table, void* out, int rows, int dim, int vocab,
                     int dtype, cudaStream_t stream);
int splice_rows_run(Context* ctx, void* embeds, const void* feats, const int32_t* start, int batch, int seq,
                    int n_patch, int dim, int dtype, cudaStream_t stream);
int copy_rows_run(Context* ctx, const void* src, int64_t sbs, int64_t srs, void* dst, int64_t dbs, int64_t drs,
                  int batch, int rows, int cols, int dtype, cudaStream_t stream);
int video_pool_run(Context* ctx, const void* feats, void* out, int batch, int frames, int patches, int dim, int dtype,
                   cudaStream_t stream);
int argmax_run(Context* ctx, const float* logits, int64_t ld, int64_t* out, int rows, int cols, cudaStream_t stream);

// sampling.cu -- temperature / top-p step (inverse CDF over the filtered distribution, one uniform per row)
int sample_step_run(Context* ctx, const float* logits, int64_t ld, int rows, int cols, float temperature, float top_p,
                    int top_k, const float* uniforms, int64_t uni_ld, int64_t* cur_ids, int64_t* seqs, int64_t seqs_ld,
                    const void* final_h, void* hid_buf, int64_t hid_bs, int hdim, uint8_t* finished, int eos_id,
                    int pad_id, int32_t* pos_dev, float* probs_out, cudaStream_t stream);

// eval_ops.cu -- metrics of evaluation/eval_ullava.py:validate on the device
int mask_iou_counts_run(Context* ctx, const void* pred, int pred_kind, const void* target, int target_kind, int n,
                        int64_t hw, int ignore, int32_t* counts, cudaStream_t s);
int seg_meter_update_run(Context* ctx, const int32_t* counts, const int32_t* offsets, int n_images, double* state,
                         cudaStream_t s);
int box_iou_diag_run(Context* ctx, const void* pred, const void* gt, int n, int dtype, float* iou, double* meter,
                     cudaStream_t s);

size_t cross_entropy_scratch(int batch, int T_len);
int cross_entropy_run(Context* ctx, const void* logits, int logits_f32, int dtype, int64_t ld_row, int64_t ld_batch,
                      const int64_t* labels, int64_t labels_ld, int batch, int T_len, int cols, int ignore_index,
                      float* out, void* scratch, size_t scratch_bytes, cudaStream_t s);

// preprocess.cu -- Pillow-exact uint8 resize, CLIP / SAM normalisation
size_t resize_u8_scratch(int h, int w, int oh, int ow);
int resample_coeffs_host(int in_size, int out_size, int filter, int32_t* bounds_out, int32_t* kk_out, size_t capacity);
int resize_u8_run(Context* ctx, const uint8_t* src, int h, int w, uint8_t* dst, int oh, int ow, int filter,
                  void* scratch, size_t scratch_bytes, cudaStream_t s);
int clip_preprocess_run(Context* ctx, const uint8_t* src, int h, int w, int top, int left, int size,
                        const float* mean, const float* stdv, double rescale, void* out, int dtype, cudaStream_t s);
int sam_preprocess_run(Context* ctx, const uint8_t* src, int h, int w, int S, const float* mean, const float* stdv,
                       void* out, int dtype, cudaStream_t s);

// sam_decoder.cu
int sam_mask_decoder_run(Context* ctx, const ullava_sam_decoder_args& a, cudaStream_t stream);
size_t sam_mask_decoder_scratch(int n);
int sam_postprocess_run(Context* ctx, const void* masks, int64_t mask_stride, float* out, uint32_t* bits, int n,
                        int low_res, int img_size,
                        int in_h, int in_w, int out_h, int out_w, int dtype, cudaStream_t stream);

// sam_encoder.cu
int sam_encoder_run(Context* ctx, const ullava_sam_encoder_args& a, cudaStream_t s);
size_t sam_encoder_scratch(int batch, int img, int patch, int embed_dim, int window, int out_chans);

// models.cu
int vit_forward_run(Context* ctx, const ullava_vit_args& a, cudaStream_t stream);
size_t vit_scratch(int batch, int img, int patch, int hidden, int ffn, int k_pad);
// tail_w/tail_n: weight [tail_n, H] of the GEMM the caller runs right after the stack (lm_head), prefetch hint only
int llama_forward_run(Context* ctx, const ullava_llama_args& a, cudaStream_t s, const int32_t* pos_dev = nullptr,
                      const void* tail_w = nullptr, int tail_n = 0);
int llama_decode_step_run(Context* ctx, const ullava_decode_args& a, cudaStream_t s);
size_t llama_scratch(int rows, int hidden, int ffn);
size_t llama_chain_bytes(int layers, int hidden, int ffn, int vocab);
int llama_chain_prepare_run(Context* ctx, const ullava_decode_args& a);

}  // namespace ullava
