// extern "C" surface of libullava_sm100.so (declared in include/ullava_sm100.h).
// Thin argument checks + dispatch; no exception crosses this boundary.
#include "common.cuh"
#include "ullava_internal.h"

using namespace ullava;

#define CTX_CHECK(name)                              \
  if (!ctx) {                                        \
    set_last_error(name ": ctx is NULL");            \
    return ERR_BAD_ARG;                              \
  }

extern "C" {

int ullava_layernorm(ullava_ctx* ctx, const void* x, int64_t ldx, const void* weight, const void* bias, void* y,
                     int64_t ldy, int32_t rows, int32_t cols, float eps, int32_t act, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_layernorm");
  return layernorm_run(ctx, x, ldx, weight, bias, y, ldy, rows, cols, eps, act, dtype, static_cast<cudaStream_t>(stream));
}

int ullava_rmsnorm(ullava_ctx* ctx, const void* x, int64_t ldx, const void* weight, void* y, int64_t ldy, int32_t rows,
                   int32_t cols, float eps, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_rmsnorm");
  return rmsnorm_run(ctx, x, ldx, weight, y, ldy, rows, cols, eps, dtype, static_cast<cudaStream_t>(stream));
}

int ullava_attention(ullava_ctx* ctx, const ullava_attn_args* args, void* stream) {
  CTX_CHECK("ullava_attention");
  if (!args) { set_last_error("ullava_attention: args is NULL"); return ERR_BAD_ARG; }
  return attention_run(ctx, *args, static_cast<cudaStream_t>(stream));
}

int ullava_attention_relpos(ullava_ctx* ctx, const ullava_attn_args* args, const void* rel_h, const void* rel_w,
                             int32_t grid_side, const int32_t* o_row_map, void* stream) {
  CTX_CHECK("ullava_attention_relpos");
  if (!args) { set_last_error("ullava_attention_relpos: args is NULL"); return ERR_BAD_ARG; }
  return attention_relpos_run(ctx, *args, rel_h, rel_w, grid_side, o_row_map, static_cast<cudaStream_t>(stream));
}

int ullava_set_attention_impl(ullava_ctx* ctx, int32_t impl) {
  CTX_CHECK("ullava_set_attention_impl");
  if (impl < 0 || impl > 2) { set_last_error("ullava_set_attention_impl: impl must be 0, 1 or 2"); return ERR_BAD_ARG; }
  ctx->attn_impl = impl;
  return OK;
}

int ullava_set_pdl(ullava_ctx* ctx, int32_t enabled) {
  CTX_CHECK("ullava_set_pdl");
  ctx->pdl = enabled ? 1 : 0;
  return OK;
}

int ullava_attention_decode_rope(ullava_ctx* ctx, const void* qkv, int64_t ld_qkv, void* k_cache, void* v_cache,
                                 int64_t cache_bs, int64_t cache_hs, void* o, int64_t o_bs, int32_t batch, int32_t heads,
                                 int32_t head_dim, int32_t ctx_len, const int32_t* pos_dev, int32_t max_seq,
                                 const float* rope_cos, const float* rope_sin, const int32_t* pos_offset, float scale,
                                 int32_t dtype, void* stream) {
  CTX_CHECK("ullava_attention_decode_rope");
  if (!rope_cos || !rope_sin) { set_last_error("ullava_attention_decode_rope: rope tables are NULL"); return ERR_BAD_ARG; }
  if (pos_dev == nullptr && ctx_len > max_seq) { set_last_error("ullava_attention_decode_rope: ctx_len > max_seq"); return ERR_BAD_ARG; }
  return attention_decode_run(ctx, qkv, ld_qkv, k_cache, v_cache, cache_bs, cache_hs, o, o_bs, batch, heads, head_dim,
                              ctx_len, scale, dtype, static_cast<cudaStream_t>(stream), pos_dev, max_seq, rope_cos,
                              rope_sin, pos_offset);
}

int ullava_gemm_next_weight(ullava_ctx* ctx, const void* next_weight, int32_t n, int32_t k, int64_t ldb) {
  CTX_CHECK("ullava_gemm_next_weight");
  if (next_weight != nullptr && (n <= 0 || k <= 0 || ldb < k || (ldb % 8) != 0 ||
                                 (reinterpret_cast<uintptr_t>(next_weight) & 15) != 0)) {
    set_last_error("ullava_gemm_next_weight: bad shape / stride / alignment (n=%d k=%d ldb=%lld)", n, k,
                   static_cast<long long>(ldb));
    return ERR_BAD_ARG;
  }
  ctx->next_w = next_weight; ctx->next_n = n; ctx->next_k = k; ctx->next_ldb = ldb;
  return OK;
}

int ullava_set_weight_prefetch(ullava_ctx* ctx, int32_t tiles_per_sm) {
  CTX_CHECK("ullava_set_weight_prefetch");
  if (tiles_per_sm < 0 || tiles_per_sm > 256) {
    set_last_error("ullava_set_weight_prefetch: tiles_per_sm must be in [0, 256]");
    return ERR_BAD_ARG;
  }
  ctx->prefetch_units = tiles_per_sm;
  return OK;
}

int ullava_attention_decode(ullava_ctx* ctx, const void* q, int64_t q_bs, const void* k_cache, const void* v_cache,
                            int64_t cache_bs, int64_t cache_hs, void* o, int64_t o_bs, int32_t batch, int32_t heads,
                            int32_t head_dim, int32_t ctx_len, float scale, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_attention_decode");
  return attention_decode_run(ctx, q, q_bs, k_cache, v_cache, cache_bs, cache_hs, o, o_bs, batch, heads, head_dim,
                              ctx_len, scale, dtype, static_cast<cudaStream_t>(stream));
}

int ullava_rope_kvcache(ullava_ctx* ctx, void* qkv, int64_t ld_qkv, void* k_cache, void* v_cache, int64_t cache_bs,
                        int64_t cache_hs, int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, int32_t pos0,
                        const float* cos_table, const float* sin_table, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_rope_kvcache");
  return rope_kvcache_run(ctx, qkv, ld_qkv, k_cache, v_cache, cache_bs, cache_hs, batch, seq, heads, head_dim, pos0,
                          cos_table, sin_table, dtype, static_cast<cudaStream_t>(stream));
}

int ullava_vit_im2col(ullava_ctx* ctx, const void* pixels, void* out, int32_t batch, int32_t img, int32_t patch,
                      int32_t k_pad, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_vit_im2col");
  return vit_im2col_run(ctx, pixels, out, batch, img, patch, k_pad, dtype, static_cast<cudaStream_t>(stream));
}

int ullava_vit_assemble(ullava_ctx* ctx, const void* patch_embeds, const void* cls, const void* pos, void* out,
                        int32_t batch, int32_t n_patches, int32_t dim, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_vit_assemble");
  return vit_assemble_run(ctx, patch_embeds, cls, pos, out, batch, n_patches, dim, dtype,
                          static_cast<cudaStream_t>(stream));
}

int ullava_embed_gather(ullava_ctx* ctx, const int64_t* ids, const void* table, void* out, int32_t rows, int32_t dim,
                        int32_t vocab, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_embed_gather");
  return embed_gather_run(ctx, ids, table, out, rows, dim, vocab, dtype, static_cast<cudaStream_t>(stream));
}

int ullava_splice_rows(ullava_ctx* ctx, void* embeds, const void* feats, const int32_t* start, int32_t batch,
                       int32_t seq, int32_t n_patch, int32_t dim, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_splice_rows");
  return splice_rows_run(ctx, embeds, feats, start, batch, seq, n_patch, dim, dtype, static_cast<cudaStream_t>(stream));
}

int ullava_copy_rows(ullava_ctx* ctx, const void* src, int64_t src_bs, int64_t src_rs, void* dst, int64_t dst_bs,
                     int64_t dst_rs, int32_t batch, int32_t rows, int32_t cols, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_copy_rows");
  return copy_rows_run(ctx, src, src_bs, src_rs, dst, dst_bs, dst_rs, batch, rows, cols, dtype,
                       static_cast<cudaStream_t>(stream));
}

int ullava_argmax(ullava_ctx* ctx, const float* logits, int64_t ld, int64_t* out, int32_t rows, int32_t cols,
                  void* stream) {
  CTX_CHECK("ullava_argmax");
  return argmax_run(ctx, logits, ld, out, rows, cols, static_cast<cudaStream_t>(stream));
}

int ullava_sam_mask_decoder(ullava_ctx* ctx, const ullava_sam_decoder_args* args, void* stream) {
  CTX_CHECK("ullava_sam_mask_decoder");
  if (!args) { set_last_error("ullava_sam_mask_decoder: args is NULL"); return ERR_BAD_ARG; }
  return sam_mask_decoder_run(ctx, *args, static_cast<cudaStream_t>(stream));
}

size_t ullava_sam_mask_decoder_scratch_bytes(int32_t n_prompts) { return sam_mask_decoder_scratch(n_prompts); }

int ullava_sam_postprocess(ullava_ctx* ctx, const void* masks, int64_t mask_stride, float* out, uint32_t* packed_bits,
                           int32_t n, int32_t low_res, int32_t img_size, int32_t in_h, int32_t in_w, int32_t out_h,
                           int32_t out_w, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_sam_postprocess");
  return sam_postprocess_run(ctx, masks, mask_stride, out, packed_bits, n, low_res, img_size, in_h, in_w, out_h, out_w,
                             dtype, static_cast<cudaStream_t>(stream));
}

int ullava_vit_forward(ullava_ctx* ctx, const ullava_vit_args* args, void* stream) {
  CTX_CHECK("ullava_vit_forward");
  if (!args) { set_last_error("ullava_vit_forward: args is NULL"); return ERR_BAD_ARG; }
  return vit_forward_run(ctx, *args, static_cast<cudaStream_t>(stream));
}

size_t ullava_vit_scratch_bytes(int32_t batch, int32_t img, int32_t patch, int32_t hidden, int32_t ffn, int32_t k_pad) {
  return vit_scratch(batch, img, patch, hidden, ffn, k_pad);
}

int ullava_sam_encoder_forward(ullava_ctx* ctx, const ullava_sam_encoder_args* args, void* stream) {
  CTX_CHECK("ullava_sam_encoder_forward");
  if (!args) { set_last_error("ullava_sam_encoder_forward: args is NULL"); return ERR_BAD_ARG; }
  return sam_encoder_run(ctx, *args, static_cast<cudaStream_t>(stream));
}

size_t ullava_sam_encoder_scratch_bytes(int32_t batch, int32_t img, int32_t patch, int32_t embed_dim, int32_t window,
                                        int32_t out_chans) {
  return sam_encoder_scratch(batch, img, patch, embed_dim, window, out_chans);
}

int ullava_llama_forward(ullava_ctx* ctx, const ullava_llama_args* args, void* stream) {
  CTX_CHECK("ullava_llama_forward");
  if (!args) { set_last_error("ullava_llama_forward: args is NULL"); return ERR_BAD_ARG; }
  return llama_forward_run(ctx, *args, static_cast<cudaStream_t>(stream));
}

int ullava_llama_decode_step(ullava_ctx* ctx, const ullava_decode_args* args, void* stream) {
  CTX_CHECK("ullava_llama_decode_step");
  if (!args) { set_last_error("ullava_llama_decode_step: args is NULL"); return ERR_BAD_ARG; }
  return llama_decode_step_run(ctx, *args, static_cast<cudaStream_t>(stream));
}

int ullava_debug_fmha_trace(ullava_ctx* ctx, void* buf) {
  CTX_CHECK("ullava_debug_fmha_trace");
  ctx->fmha_trace = buf;
  return OK;
}

int ullava_debug_chain_trace(ullava_ctx* ctx, void* buf) {
  CTX_CHECK("ullava_debug_chain_trace");
  ctx->chain_trace = buf;
  return OK;
}

size_t ullava_llama_chain_bytes(int32_t layers, int32_t hidden, int32_t ffn, int32_t vocab) {
  return llama_chain_bytes(layers, hidden, ffn, vocab);
}

int ullava_llama_chain_prepare(ullava_ctx* ctx, const ullava_decode_args* args) {
  CTX_CHECK("ullava_llama_chain_prepare");
  if (!args) { set_last_error("ullava_llama_chain_prepare: args is NULL"); return ERR_BAD_ARG; }
  return llama_chain_prepare_run(ctx, *args);
}

int ullava_greedy_step(ullava_ctx* ctx, const float* logits, int64_t ld, int32_t rows, int32_t cols, int64_t* cur_ids,
                       int64_t* seqs, int64_t seqs_ld, const void* final_h, void* hid_buf, int64_t hid_bs, int32_t hdim,
                       uint8_t* finished, int32_t eos_id, int32_t pad_id, int32_t* pos_dev, void* stream) {
  CTX_CHECK("ullava_greedy_step");
  return greedy_step_run(ctx, logits, ld, rows, cols, cur_ids, seqs, seqs_ld, final_h, hid_buf, hid_bs, hdim, finished,
                         eos_id, pad_id, pos_dev, static_cast<cudaStream_t>(stream));
}

int ullava_video_pool(ullava_ctx* ctx, const void* feats, void* out, int32_t batch, int32_t frames, int32_t patches,
                      int32_t dim, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_video_pool");
  return video_pool_run(ctx, feats, out, batch, frames, patches, dim, dtype, static_cast<cudaStream_t>(stream));
}

int ullava_sample_step(ullava_ctx* ctx, const float* logits, int64_t ld, int32_t rows, int32_t cols, float temperature,
                       float top_p, int32_t top_k, const float* uniforms, int64_t uniforms_ld, int64_t* cur_ids, int64_t* seqs,
                       int64_t seqs_ld, const void* final_h, void* hid_buf, int64_t hid_bs, int32_t hdim,
                       uint8_t* finished, int32_t eos_id, int32_t pad_id, int32_t* pos_dev, float* probs_out,
                       void* stream) {
  CTX_CHECK("ullava_sample_step");
  return sample_step_run(ctx, logits, ld, rows, cols, temperature, top_p, top_k, uniforms, uniforms_ld, cur_ids, seqs, seqs_ld,
                         final_h, hid_buf, hid_bs, hdim, finished, eos_id, pad_id, pos_dev, probs_out,
                         static_cast<cudaStream_t>(stream));
}

int ullava_mask_iou_counts(ullava_ctx* ctx, const void* pred, int32_t pred_kind, const void* target, int32_t target_kind,
                           int32_t n, int64_t hw, int32_t ignore_index, int32_t* counts, void* stream) {
  CTX_CHECK("ullava_mask_iou_counts");
  return mask_iou_counts_run(ctx, pred, pred_kind, target, target_kind, n, hw, ignore_index, counts,
                             static_cast<cudaStream_t>(stream));
}

int ullava_seg_meter_update(ullava_ctx* ctx, const int32_t* counts, const int32_t* offsets, int32_t n_images,
                            double* state, void* stream) {
  CTX_CHECK("ullava_seg_meter_update");
  return seg_meter_update_run(ctx, counts, offsets, n_images, state, static_cast<cudaStream_t>(stream));
}

int ullava_box_iou_diag(ullava_ctx* ctx, const void* pred, const void* gt, int32_t n, int32_t dtype, float* iou,
                        double* meter, void* stream) {
  CTX_CHECK("ullava_box_iou_diag");
  return box_iou_diag_run(ctx, pred, gt, n, dtype, iou, meter, static_cast<cudaStream_t>(stream));
}

size_t ullava_cross_entropy_scratch_bytes(int32_t batch, int32_t seq) {
  if (batch <= 0 || seq <= 1) return 256;
  return cross_entropy_scratch(batch, seq);
}

int ullava_cross_entropy(ullava_ctx* ctx, const void* logits, int32_t logits_f32, int32_t dtype, int64_t ld_row,
                         int64_t ld_batch, const int64_t* labels, int64_t labels_ld, int32_t batch, int32_t seq,
                         int32_t cols, int32_t ignore_index, float* out, void* scratch, size_t scratch_bytes,
                         void* stream) {
  CTX_CHECK("ullava_cross_entropy");
  return cross_entropy_run(ctx, logits, logits_f32, dtype, ld_row, ld_batch, labels, labels_ld, batch, seq, cols,
                           ignore_index, out, scratch, scratch_bytes, static_cast<cudaStream_t>(stream));
}

int ullava_resample_coeffs(int32_t in_size, int32_t out_size, int32_t filter, int32_t* bounds, int32_t* taps,
                           size_t capacity) {
  return resample_coeffs_host(in_size, out_size, filter, bounds, taps, capacity);
}

size_t ullava_resize_u8_scratch_bytes(int32_t h, int32_t w, int32_t out_h, int32_t out_w) {
  if (h <= 0 || w <= 0 || out_h <= 0 || out_w <= 0) return 0;
  return resize_u8_scratch(h, w, out_h, out_w);
}

int ullava_resize_u8(ullava_ctx* ctx, const uint8_t* src, int32_t h, int32_t w, uint8_t* dst, int32_t out_h, int32_t out_w,
                     int32_t filter, void* scratch, size_t scratch_bytes, void* stream) {
  CTX_CHECK("ullava_resize_u8");
  return resize_u8_run(ctx, src, h, w, dst, out_h, out_w, filter, scratch, scratch_bytes,
                       static_cast<cudaStream_t>(stream));
}

int ullava_clip_preprocess(ullava_ctx* ctx, const uint8_t* src, int32_t h, int32_t w, int32_t top, int32_t left,
                           int32_t size, const float* mean, const float* std, double rescale, void* out, int32_t dtype,
                           void* stream) {
  CTX_CHECK("ullava_clip_preprocess");
  return clip_preprocess_run(ctx, src, h, w, top, left, size, mean, std, rescale, out, dtype,
                             static_cast<cudaStream_t>(stream));
}

int ullava_sam_preprocess(ullava_ctx* ctx, const uint8_t* src, int32_t h, int32_t w, int32_t sam_size, const float* mean,
                          const float* std, void* out, int32_t dtype, void* stream) {
  CTX_CHECK("ullava_sam_preprocess");
  return sam_preprocess_run(ctx, src, h, w, sam_size, mean, std, out, dtype, static_cast<cudaStream_t>(stream));
}

size_t ullava_llama_scratch_bytes(int32_t rows, int32_t hidden_size, int32_t ffn) {
  return llama_scratch(rows, hidden_size, ffn);
}

}  // extern "C"
