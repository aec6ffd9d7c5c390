// HBM-bound glue kernels of the u-LLaVA hot path (all coalesced 16-byte accesses where the
// layout allows): RoPE + KV-cache scatter, CLIP im2col / embedding assembly, token-embedding
// gather, image-feature splice, strided row copy, greedy argmax.
#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

// ---------------------------------------------------------------------------------------------
// RoPE (rotate_half, hf:models/llama/modeling_llama.py:137-168) on the q,k parts of packed QKV
// + scatter of k (rotated) and v into the KV cache.  One CTA per token row.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
rope_kvcache_kernel(T* __restrict__ qkv, int64_t ld, T* __restrict__ kc, T* __restrict__ vc, int64_t cache_bs,
                    int64_t cache_hs, int seq, int heads, int hd, int pos0, const float* __restrict__ cos_t,
                    const float* __restrict__ sin_t, const int32_t* __restrict__ pos_dev) {
  // Launched with programmatic serialization in the decode step: the CTA is resident (launch latency paid) while the
  // QKV GEMM still runs; nothing is read before griddepcontrol.wait.
  pdl_launch_dependents();
  pdl_wait();
  if (pos_dev) pos0 = *pos_dev;  // CUDA-graph decode: the position lives in device memory
  const int row = blockIdx.x;
  const int b = row / seq, s = row - b * seq;
  const int pos = pos0 + s;
  const int half = hd >> 1;
  const int vph = half >> 3;               // 8-element vectors per half head
  T* base = qkv + static_cast<int64_t>(row) * ld;
  const float* cs = cos_t + static_cast<int64_t>(pos) * half;
  const float* sn = sin_t + static_cast<int64_t>(pos) * half;
  const int hdim = heads * hd;
  const int n_rot = heads * vph;           // work items for q (and for k)
  for (int it = threadIdx.x; it < 2 * n_rot; it += blockDim.x) {
    const bool is_k = it >= n_rot;
    const int w = is_k ? it - n_rot : it;
    const int h = w / vph, j = (w - h * vph) * 8;
    T* p = base + (is_k ? hdim : 0) + h * hd + j;
    const uint4 lo = *reinterpret_cast<const uint4*>(p);
    const uint4 hi = *reinterpret_cast<const uint4*>(p + half);
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
    if (is_k) {
      T* d = kc + b * cache_bs + h * cache_hs + static_cast<int64_t>(pos) * hd + j;
      *reinterpret_cast<uint4*>(d) = vlo;
      *reinterpret_cast<uint4*>(d + half) = vhi;
    } else {
      *reinterpret_cast<uint4*>(p) = vlo;
      *reinterpret_cast<uint4*>(p + half) = vhi;
    }
  }
  const int n_v = hdim >> 3;
  for (int it = threadIdx.x; it < n_v; it += blockDim.x) {
    const int e = it * 8;
    const int h = e / hd, j = e - h * hd;
    const uint4 v = *reinterpret_cast<const uint4*>(base + 2 * hdim + e);
    *reinterpret_cast<uint4*>(vc + b * cache_bs + h * cache_hs + static_cast<int64_t>(pos) * hd + j) = v;
  }
}

int rope_kvcache_run(Context* ctx, void* qkv, int64_t ld, void* kc, void* vc, int64_t cache_bs, int64_t cache_hs,
                     int batch, int seq, int heads, int hd, int pos0, const float* cos_t, const float* sin_t,
                     int dtype, cudaStream_t stream, const int32_t* pos_dev) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_GLUE, 0.0, 2.0 * batch * seq * 3.0 * heads * hd * 2.0);
  ULLAVA_REQUIRE(qkv && kc && vc && cos_t && sin_t, "rope: null pointer");
  ULLAVA_REQUIRE(hd % 16 == 0 && ld % 8 == 0 && cache_bs % 8 == 0 && cache_hs % 8 == 0, "rope: bad alignment");
  const int rows = batch * seq;
  if (rows == 0) return OK;
  const bool pdl = ctx->pdl != 0 && seq == 1;
  cudaError_t e;
  if (dtype == DT_BF16)
    e = launch_pdl(rope_kvcache_kernel<__nv_bfloat16>, dim3(rows), dim3(256), 0, stream, pdl,
                   static_cast<__nv_bfloat16*>(qkv), ld, static_cast<__nv_bfloat16*>(kc),
                   static_cast<__nv_bfloat16*>(vc), cache_bs, cache_hs, seq, heads, hd, pos0, cos_t, sin_t, pos_dev);
  else if (dtype == DT_F16)
    e = launch_pdl(rope_kvcache_kernel<__half>, dim3(rows), dim3(256), 0, stream, pdl, static_cast<__half*>(qkv), ld,
                   static_cast<__half*>(kc), static_cast<__half*>(vc), cache_bs, cache_hs, seq, heads, hd, pos0, cos_t,
                   sin_t, pos_dev);
  else { set_last_error("rope: unsupported dtype"); return ERR_UNSUPPORTED; }
  ctx->launches++;
  return check_cuda(e, "rope_kvcache launch");
}

// ---------------------------------------------------------------------------------------------
// CLIP patchify im2col: [B,3,img,img] -> [B*g*g, k_pad], column = c*P*P + ky*P + kx.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
im2col_kernel(const T* __restrict__ px, T* __restrict__ out, int img, int patch, int k_pad, int64_t total) {
  const int g = img / patch;
  const int kk = 3 * patch * patch;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % k_pad);
    const int64_t r = i / k_pad;
    T v = T16<T>::from_f(0.f);
    if (col < kk) {
      const int c = col / (patch * patch);
      const int rem = col - c * patch * patch;
      const int ky = rem / patch, kx = rem - ky * patch;
      const int b = static_cast<int>(r / (g * g));
      const int pr = static_cast<int>(r - static_cast<int64_t>(b) * g * g);
      const int py = pr / g, pxx = pr - py * g;
      v = px[((static_cast<int64_t>(b) * 3 + c) * img + (py * patch + ky)) * img + pxx * patch + kx];
    }
    out[i] = v;
  }
}

int vit_im2col_run(Context* ctx, const void* pixels, void* out, int batch, int img, int patch, int k_pad, int dtype,
                   cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_GLUE, 0.0, 2.0 * batch * (3.0 * img * img + (double)(img / patch) * (img / patch) * k_pad));
  ULLAVA_REQUIRE(pixels && out, "im2col: null pointer");
  ULLAVA_REQUIRE(img % patch == 0 && k_pad >= 3 * patch * patch, "im2col: bad geometry");
  const int g = img / patch;
  const int64_t total = static_cast<int64_t>(batch) * g * g * k_pad;
  if (total == 0) return OK;
  const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 16));
  if (dtype == DT_BF16)
    im2col_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(pixels),
                                                           static_cast<__nv_bfloat16*>(out), img, patch, k_pad, total);
  else if (dtype == DT_F16)
    im2col_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(pixels), static_cast<__half*>(out),
                                                    img, patch, k_pad, total);
  else { set_last_error("im2col: unsupported dtype"); return ERR_UNSUPPORTED; }
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "im2col launch");
}

// ---------------------------------------------------------------------------------------------
// CLIP embeddings: out[b,0] = cls + pos[0]; out[b,1+p] = patch[b,p] + pos[1+p]   (dim % 8 == 0)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
vit_assemble_kernel(const T* __restrict__ pe, const T* __restrict__ cls, const T* __restrict__ pos,
                    T* __restrict__ out, int np, int dim) {
  const int tok = blockIdx.x;             // 0..np
  const int b = blockIdx.y;
  const T* src = (tok == 0) ? cls : pe + (static_cast<int64_t>(b) * np + (tok - 1)) * dim;
  const T* ps = pos + static_cast<int64_t>(tok) * dim;
  T* dst = out + (static_cast<int64_t>(b) * (np + 1) + tok) * dim;
  for (int v = threadIdx.x; v < dim / 8; v += blockDim.x) {
    const uint4 a = *reinterpret_cast<const uint4*>(src + v * 8);
    const uint4 p = *reinterpret_cast<const uint4*>(ps + v * 8);
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, pu[4] = {p.x, p.y, p.z, p.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = unpack2<T>(au[e]), y = unpack2<T>(pu[e]);
      o[e] = pack2<T>(x.x + y.x, x.y + y.y);
    }
    *reinterpret_cast<uint4*>(dst + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

int vit_assemble_run(Context* ctx, const void* pe, const void* cls, const void* pos, void* out, int batch, int np,
                     int dim, int dtype, cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_GLUE, 0.0, 4.0 * batch * (np + 1.0) * dim);
  ULLAVA_REQUIRE(pe && cls && pos && out, "vit_assemble: null pointer");
  ULLAVA_REQUIRE(dim % 8 == 0, "vit_assemble: dim must be a multiple of 8");
  if (batch == 0) return OK;
  dim3 grid(np + 1, batch);
  if (dtype == DT_BF16)
    vit_assemble_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(pe), static_cast<const __nv_bfloat16*>(cls),
        static_cast<const __nv_bfloat16*>(pos), static_cast<__nv_bfloat16*>(out), np, dim);
  else if (dtype == DT_F16)
    vit_assemble_kernel<__half><<<grid, 128, 0, stream>>>(static_cast<const __half*>(pe),
                                                          static_cast<const __half*>(cls),
                                                          static_cast<const __half*>(pos), static_cast<__half*>(out),
                                                          np, dim);
  else { set_last_error("vit_assemble: unsupported dtype"); return ERR_UNSUPPORTED; }
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "vit_assemble launch");
}

// ---------------------------------------------------------------------------------------------
// Generic strided row copy (16-bit): dst[b][r][0:cols] = src[b][r][0:cols]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
copy_rows_kernel(const uint16_t* __restrict__ src, int64_t sbs, int64_t srs, uint16_t* __restrict__ dst, int64_t dbs,
                 int64_t drs, int rows, int cols, int vec_ok) {
  const int r = blockIdx.x, b = blockIdx.y;
  const uint16_t* s = src + b * sbs + r * srs;
  uint16_t* d = dst + b * dbs + r * drs;
  if (vec_ok) {
    for (int v = threadIdx.x; v < cols / 8; v += blockDim.x)
      reinterpret_cast<uint4*>(d)[v] = reinterpret_cast<const uint4*>(s)[v];
  } else {
    for (int c = threadIdx.x; c < cols; c += blockDim.x) d[c] = s[c];
  }
}

int copy_rows_run(Context* ctx, const void* src, int64_t sbs, int64_t srs, void* dst, int64_t dbs, int64_t drs,
                  int batch, int rows, int cols, int dtype, cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_GLUE, 0.0, 4.0 * batch * rows * cols);
  ULLAVA_REQUIRE(src && dst, "copy_rows: null pointer");
  ULLAVA_REQUIRE(dtype == DT_BF16 || dtype == DT_F16, "copy_rows: 16-bit dtypes only");
  if (batch == 0 || rows == 0 || cols == 0) return OK;
  const int vec_ok = (cols % 8 == 0) && (sbs % 8 == 0) && (srs % 8 == 0) && (dbs % 8 == 0) && (drs % 8 == 0) &&
                     ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  dim3 grid(rows, batch);
  copy_rows_kernel<<<grid, 128, 0, stream>>>(static_cast<const uint16_t*>(src), sbs, srs,
                                             static_cast<uint16_t*>(dst), dbs, drs, rows, cols, vec_ok);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "copy_rows launch");
}

// ---------------------------------------------------------------------------------------------
// Token embedding gather and image-feature splice
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
embed_gather_kernel(const int64_t* __restrict__ ids, const uint16_t* __restrict__ table, uint16_t* __restrict__ out,
                    int dim, int vocab) {
  const int r = blockIdx.x;
  int64_t id = ids[r];
  if (id < 0 || id >= vocab) id = 0;  // defensive: torch would raise; callers validate on the host
  const uint4* s = reinterpret_cast<const uint4*>(table + id * dim);
  uint4* d = reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * dim);
  for (int v = threadIdx.x; v < dim / 8; v += blockDim.x) d[v] = s[v];
}

int embed_gather_run(Context* ctx, const int64_t* ids, const void* table, void* out, int rows, int dim, int vocab,
                     int dtype, cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_GLUE, 0.0, 4.0 * rows * dim);
  ULLAVA_REQUIRE(ids && table && out, "embed_gather: null pointer");
  ULLAVA_REQUIRE(dim % 8 == 0, "embed_gather: dim must be a multiple of 8");
  ULLAVA_REQUIRE(dtype == DT_BF16 || dtype == DT_F16, "embed_gather: 16-bit dtypes only");
  if (rows == 0) return OK;
  embed_gather_kernel<<<rows, 128, 0, stream>>>(ids, static_cast<const uint16_t*>(table),
                                                static_cast<uint16_t*>(out), dim, vocab);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "embed_gather launch");
}

__global__ void __launch_bounds__(128)
splice_rows_kernel(uint16_t* __restrict__ embeds, const uint16_t* __restrict__ feats, const int32_t* __restrict__ start,
                   int seq, int n_patch, int dim) {
  const int p = blockIdx.x, b = blockIdx.y;
  const int st = start[b];
  if (st < 0 || st + 1 + p >= seq) return;
  const uint4* s = reinterpret_cast<const uint4*>(feats + (static_cast<int64_t>(b) * n_patch + p) * dim);
  uint4* d = reinterpret_cast<uint4*>(embeds + (static_cast<int64_t>(b) * seq + st + 1 + p) * dim);
  for (int v = threadIdx.x; v < dim / 8; v += blockDim.x) d[v] = s[v];
}

int splice_rows_run(Context* ctx, void* embeds, const void* feats, const int32_t* start, int batch, int seq,
                    int n_patch, int dim, int dtype, cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_GLUE, 0.0, 4.0 * batch * n_patch * dim);
  ULLAVA_REQUIRE(embeds && feats && start, "splice_rows: null pointer");
  ULLAVA_REQUIRE(dim % 8 == 0, "splice_rows: dim must be a multiple of 8");
  ULLAVA_REQUIRE(dtype == DT_BF16 || dtype == DT_F16, "splice_rows: 16-bit dtypes only");
  if (batch == 0 || n_patch == 0) return OK;
  dim3 grid(n_patch, batch);
  splice_rows_kernel<<<grid, 128, 0, stream>>>(static_cast<uint16_t*>(embeds), static_cast<const uint16_t*>(feats),
                                               start, seq, n_patch, dim);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "splice_rows launch");
}

// ---------------------------------------------------------------------------------------------
// Video pooling (UllavaCoreForCausalLM.encode_video, models/ullava_core.py:160-180): per-frame patch features
// feats [bs, T, N, D] -> out [bs, T + N, D]; rows 0..T-1 = mean over the N patches of frame t (temporal features),
// rows T..T+N-1 = mean over the T frames of patch n (spatial features).  fp32 accumulation, one rounding.
// One CTA per output row, 8 columns per thread, 16-byte loads that are contiguous across the CTA.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
video_pool_kernel(const T* __restrict__ feats, T* __restrict__ out, int frames, int patches, int dim) {
  const int b = blockIdx.y, r = blockIdx.x;
  const bool temporal = r < frames;
  const int count = temporal ? patches : frames;
  const int64_t step = temporal ? dim : static_cast<int64_t>(patches) * dim;
  const T* base = feats + static_cast<int64_t>(b) * frames * patches * dim +
                  (temporal ? static_cast<int64_t>(r) * patches * dim : static_cast<int64_t>(r - frames) * dim);
  const float inv = 1.0f / static_cast<float>(count);
  for (int c = threadIdx.x * 8; c < dim; c += blockDim.x * 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < count; ++i) {
      const uint4 v = *reinterpret_cast<const uint4*>(base + i * step + c);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack2<T>(u[e]);
        acc[2 * e] += f.x;
        acc[2 * e + 1] += f.y;
      }
    }
    uint4 o;
    o.x = pack2<T>(acc[0] * inv, acc[1] * inv); o.y = pack2<T>(acc[2] * inv, acc[3] * inv);
    o.z = pack2<T>(acc[4] * inv, acc[5] * inv); o.w = pack2<T>(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(out + (static_cast<int64_t>(b) * (frames + patches) + r) * dim + c) = o;
  }
}

int video_pool_run(Context* ctx, const void* feats, void* out, int batch, int frames, int patches, int dim, int dtype,
                   cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_GLUE, 0.0, 2.0 * batch * (2.0 * frames * patches + frames + patches) * dim);
  ULLAVA_REQUIRE(feats && out, "video_pool: null pointer");
  ULLAVA_REQUIRE(frames > 0 && patches > 0 && dim > 0 && dim % 8 == 0, "video_pool: bad shape");
  if (batch == 0) return OK;
  dim3 grid(frames + patches, batch);
  if (dtype == DT_BF16)
    video_pool_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(feats),
                                                             static_cast<__nv_bfloat16*>(out), frames, patches, dim);
  else if (dtype == DT_F16)
    video_pool_kernel<__half><<<grid, 128, 0, stream>>>(static_cast<const __half*>(feats), static_cast<__half*>(out),
                                                      frames, patches, dim);
  else { set_last_error("video_pool: unsupported dtype"); return ERR_UNSUPPORTED; }
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "video_pool launch");
}

// ---------------------------------------------------------------------------------------------
// Greedy argmax over fp32 logits, first index wins ties.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
argmax_kernel(const float* __restrict__ logits, int64_t ld, int64_t* __restrict__ out, int cols) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const float* row = logits + blockIdx.x * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = row[c];
    if (v > best || (v == best && c < bi)) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sv[w] = best; si[w] = bi; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    best = l < nw ? sv[l] : -INFINITY;
    bi = l < nw ? si[l] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (l == 0) out[blockIdx.x] = (bi == 0x7fffffff) ? 0 : bi;
  }
}

// ---------------------------------------------------------------------------------------------
// Device-side bookkeeping of one greedy decode step (CUDA-graph replayable: the position is read from
// device memory).  pos = index of the token that was just processed.
//   greedy_step_kernel : next = argmax(logits[b]); finished/eos/pad handling; cur_ids[b] = next;
//                        seqs[b][pos + 1] = next; hid_buf[b][pos] = final[b]
//   advance_pos_kernel : ++*pos  (last node of the step)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
greedy_step_kernel(const float* __restrict__ logits, int64_t ld, int cols, int64_t* __restrict__ cur_ids,
                   int64_t* __restrict__ seqs, int64_t seqs_ld, const uint16_t* __restrict__ final_h,
                   uint16_t* __restrict__ hid_buf, int64_t hid_bs, int hdim, uint8_t* __restrict__ finished,
                   int eos_id, int pad_id, const int32_t* __restrict__ pos_dev) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const int b = blockIdx.x;
  const int pos = *pos_dev;
  if (hid_buf) {
    const uint4* src = reinterpret_cast<const uint4*>(final_h + static_cast<int64_t>(b) * hdim);
    uint4* dst = reinterpret_cast<uint4*>(hid_buf + b * hid_bs + static_cast<int64_t>(pos) * hdim);
    for (int v = threadIdx.x; v < hdim / 8; v += blockDim.x) dst[v] = src[v];
  }
  const float* row = logits + b * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = row[c];
    if (v > best || (v == best && c < bi)) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sv[w] = best; si[w] = bi; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    best = l < nw ? sv[l] : -INFINITY;
    bi = l < nw ? si[l] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (l == 0) {
      int64_t nxt = (bi == 0x7fffffff) ? 0 : bi;
      if (finished) {
        if (finished[b]) nxt = pad_id;
        else if (eos_id >= 0 && nxt == eos_id) finished[b] = 1;
      }
      cur_ids[b] = nxt;
      if (seqs) seqs[b * seqs_ld + pos + 1] = nxt;
    }
  }
}

__global__ void advance_pos_kernel(int32_t* pos) { *pos += 1; }

int greedy_step_run(Context* ctx, const float* logits, int64_t ld, int rows, int cols, int64_t* cur_ids, int64_t* seqs,
                    int64_t seqs_ld, const void* final_h, void* hid_buf, int64_t hid_bs, int hdim, uint8_t* finished,
                    int eos_id, int pad_id, int32_t* pos_dev, cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_GLUE, 0.0, 4.0 * rows * cols);
  ULLAVA_REQUIRE(logits && cur_ids && pos_dev && cols > 0 && hdim % 8 == 0, "greedy_step: bad arguments");
  if (rows == 0) return OK;
  greedy_step_kernel<<<rows, 1024, 0, stream>>>(logits, ld, cols, cur_ids, seqs, seqs_ld,
                                                static_cast<const uint16_t*>(final_h), static_cast<uint16_t*>(hid_buf),
                                                hid_bs, hdim, finished, eos_id, pad_id, pos_dev);
  advance_pos_kernel<<<1, 1, 0, stream>>>(pos_dev);
  ctx->launches += 2;
  return check_cuda(cudaGetLastError(), "greedy_step launch");
}

int argmax_run(Context* ctx, const float* logits, int64_t ld, int64_t* out, int rows, int cols, cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_GLUE, 0.0, 4.0 * rows * cols);
  ULLAVA_REQUIRE(logits && out && cols > 0, "argmax: bad arguments");
  if (rows == 0) return OK;
  argmax_kernel<<<rows, 1024, 0, stream>>>(logits, ld, out, cols);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "argmax launch");
}

}  // namespace ullava
