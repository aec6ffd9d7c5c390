// LayerNorm / RMSNorm for the u-LLaVA hot path (fp32 statistics, 16-bit storage).
//   LayerNorm : CLIP pre_layrnorm + layer_norm1/2 (hf:models/clip/modeling_clip.py:354-385,677),
//               SAM TwoWayTransformer norms (segment_anything/modeling/transformer.py:151-182).
//   RMSNorm   : LlamaRMSNorm (hf:models/llama/modeling_llama.py:52-69).
// HBM-bound: one CTA per row, 16-byte vector loads, the row stays in registers between the
// statistics pass and the normalise pass (read once, write once).
#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

static constexpr int kMaxVec = 4;  // 16-byte vectors per thread kept in registers

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int nw = (blockDim.x + 31) >> 5;
  __syncthreads();  // protect red[] from the previous use
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.f;
  t = warp_sum(t);
  return t;
}

template <typename T, bool kRms>
__global__ void __launch_bounds__(512)
norm_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ w, const T* __restrict__ b,
            T* __restrict__ y, int64_t ldy, int cols, float eps, int act, const int32_t* __restrict__ dst_rows) {
  __shared__ float red[32];
  pdl_launch_dependents();  // a following weight-streaming GEMM may start prefetching its weights now
  pdl_wait();               // launched with programmatic serialization: the producer of x must have finished
  const int row = blockIdx.x;
  const T* xr = x + static_cast<int64_t>(row) * ldx;
  // dst_rows (optional) scatters the output rows, e.g. SAM's window_partition fused into norm1
  T* yr = y + static_cast<int64_t>(dst_rows ? dst_rows[row] : row) * ldy;
  const int nvec = cols >> 3;
  uint4 regs[kMaxVec];
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      regs[i] = *reinterpret_cast<const uint4*>(xr + vi * 8);
      const uint32_t u[4] = {regs[i].x, regs[i].y, regs[i].z, regs[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack2<T>(u[j]);
        sum += f.x + f.y;
        sq += f.x * f.x + f.y * f.y;
      }
    }
  }
  float mean = 0.f, rstd;
  if constexpr (kRms) {
    const float tot = block_sum(sq, red);
    rstd = rsqrtf(tot / cols + eps);
  } else {
    mean = block_sum(sum, red) / cols;
    // second pass over registers for the variance (same two-pass scheme as torch's LayerNorm)
    float d2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int vi = threadIdx.x + i * blockDim.x;
      if (vi < nvec) {
        const uint32_t u[4] = {regs[i].x, regs[i].y, regs[i].z, regs[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack2<T>(u[j]);
          d2 += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
        }
      }
    }
    const float var = block_sum(d2, red) / cols;
    rstd = rsqrtf(var + eps);
  }
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      const uint32_t u[4] = {regs[i].x, regs[i].y, regs[i].z, regs[i].w};
      const uint4 wv = *reinterpret_cast<const uint4*>(w + vi * 8);
      const uint32_t wu[4] = {wv.x, wv.y, wv.z, wv.w};
      uint32_t bu[4] = {0, 0, 0, 0};
      if constexpr (!kRms) {
        const uint4 bv = *reinterpret_cast<const uint4*>(b + vi * 8);
        bu[0] = bv.x; bu[1] = bv.y; bu[2] = bv.z; bu[3] = bv.w;
      }
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack2<T>(u[j]);
        const float2 g = unpack2<T>(wu[j]);
        if constexpr (kRms) {
          // HF: weight * (x_fp32 * rstd).to(dtype)  -> round once before the weight multiply
          const float2 n = unpack2<T>(pack2<T>(f.x * rstd, f.y * rstd));
          o[j] = pack2<T>(g.x * n.x, g.y * n.y);
        } else {
          const float2 bb = unpack2<T>(bu[j]);
          float y0 = (f.x - mean) * rstd * g.x + bb.x, y1 = (f.y - mean) * rstd * g.y + bb.y;
          if (act == EPI_GELU) {
            y0 = fast_gelu(y0);
            y1 = fast_gelu(y1);
          }
          o[j] = pack2<T>(y0, y1);
        }
      }
      *reinterpret_cast<uint4*>(yr + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}


// Many-row flavour (activations of a whole batch: ViT / SAM LayerNorms, prefill RMSNorm): one WARP per row, 8 rows per
// CTA, VPL 16-byte vectors per lane held in registers, shuffle-only reductions (no shared memory, no CTA barrier).
static constexpr int kRowsPerCta = 8;
template <typename T, bool kRms, int VPL>
__global__ void __launch_bounds__(32 * kRowsPerCta)
norm_rows_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ w, const T* __restrict__ b,
                 T* __restrict__ y, int64_t ldy, int rows, int cols, float eps, int act,
                 const int32_t* __restrict__ dst_rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kRowsPerCta + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + static_cast<int64_t>(row) * ldx;
  T* yr = y + static_cast<int64_t>(dst_rows ? dst_rows[row] : row) * ldy;
  const int nvec = cols >> 3;
  uint4 regs[VPL];
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      regs[i] = *reinterpret_cast<const uint4*>(xr + vi * 8);
      const uint32_t u[4] = {regs[i].x, regs[i].y, regs[i].z, regs[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack2<T>(u[j]);
        sum += f.x + f.y;
        sq += f.x * f.x + f.y * f.y;
      }
    }
  }
  float mean = 0.f, rstd;
  if constexpr (kRms) {
    rstd = rsqrtf(warp_sum(sq) / cols + eps);
  } else {
    mean = warp_sum(sum) / cols;
    float d2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + i * 32 < nvec) {
        const uint32_t u[4] = {regs[i].x, regs[i].y, regs[i].z, regs[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack2<T>(u[j]);
          d2 += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
        }
      }
    }
    rstd = rsqrtf(warp_sum(d2) / cols + eps);
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint32_t u[4] = {regs[i].x, regs[i].y, regs[i].z, regs[i].w};
      const uint4 wv = *reinterpret_cast<const uint4*>(w + vi * 8);
      const uint32_t wu[4] = {wv.x, wv.y, wv.z, wv.w};
      uint32_t bu[4] = {0, 0, 0, 0};
      if constexpr (!kRms) {
        const uint4 bv = *reinterpret_cast<const uint4*>(b + vi * 8);
        bu[0] = bv.x; bu[1] = bv.y; bu[2] = bv.z; bu[3] = bv.w;
      }
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack2<T>(u[j]);
        const float2 g = unpack2<T>(wu[j]);
        if constexpr (kRms) {
          const float2 n = unpack2<T>(pack2<T>(f.x * rstd, f.y * rstd));
          o[j] = pack2<T>(g.x * n.x, g.y * n.y);
        } else {
          const float2 bb = unpack2<T>(bu[j]);
          float y0 = (f.x - mean) * rstd * g.x + bb.x, y1 = (f.y - mean) * rstd * g.y + bb.y;
          if (act == EPI_GELU) {
            y0 = fast_gelu(y0);
            y1 = fast_gelu(y1);
          }
          o[j] = pack2<T>(y0, y1);
        }
      }
      *reinterpret_cast<uint4*>(yr + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

template <typename T, bool kRms>
static cudaError_t norm_rows_launch(Context* ctx, const void* x, int64_t ldx, const void* w, const void* b, void* y,
                                    int64_t ldy, int rows, int cols, float eps, int act, cudaStream_t stream,
                                    const int32_t* dst_rows) {
  const int vpl = (cols / 8 + 31) / 32;
  const dim3 grid((rows + kRowsPerCta - 1) / kRowsPerCta), block(32 * kRowsPerCta);
  const T* xx = static_cast<const T*>(x);
  const T* ww = static_cast<const T*>(w);
  const T* bb = static_cast<const T*>(b);
  T* yy = static_cast<T*>(y);
  const bool pdl = ctx->pdl != 0;
#define ULLAVA_NR(V) \
  return launch_pdl(norm_rows_kernel<T, kRms, V>, grid, block, 0, stream, pdl, xx, ldx, ww, bb, yy, ldy, rows, cols, eps, \
                    act, dst_rows)
  if (vpl <= 1) { ULLAVA_NR(1); }
  if (vpl <= 2) { ULLAVA_NR(2); }
  if (vpl <= 4) { ULLAVA_NR(4); }
  if (vpl <= 5) { ULLAVA_NR(5); }
  if (vpl <= 8) { ULLAVA_NR(8); }
  ULLAVA_NR(16);
#undef ULLAVA_NR
}

template <bool kRms>
static int norm_launch(Context* ctx, const void* x, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy,
                       int rows, int cols, float eps, int act, int dtype, cudaStream_t stream,
                       const int32_t* dst_rows = nullptr) {
  ULLAVA_REQUIRE(x && w && y && (kRms || b), "norm: null pointer");
  ULLAVA_REQUIRE(rows >= 0 && cols > 0 && (cols % 8) == 0, "norm: cols (%d) must be a positive multiple of 8", cols);
  ULLAVA_REQUIRE((ldx % 8) == 0 && (ldy % 8) == 0, "norm: ldx/ldy must be multiples of 8");
  ULLAVA_REQUIRE(cols <= 8 * kMaxVec * 512, "norm: cols (%d) too large", cols);
  ULLAVA_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w) |
                   reinterpret_cast<uintptr_t>(b)) & 15) == 0, "norm: pointers must be 16-byte aligned");
  if (rows == 0) return OK;
  const int nvec = cols / 8;
  if (rows >= 1024 && nvec <= 32 * 16) {
    // whole-batch activations: warp-per-row kernel
    if (dtype == DT_BF16)
      ULLAVA_CHECK_CUDA((norm_rows_launch<__nv_bfloat16, kRms>(ctx, x, ldx, w, b, y, ldy, rows, cols, eps, act, stream, dst_rows)));
    else if (dtype == DT_F16)
      ULLAVA_CHECK_CUDA((norm_rows_launch<__half, kRms>(ctx, x, ldx, w, b, y, ldy, rows, cols, eps, act, stream, dst_rows)));
    else { set_last_error("norm: unsupported dtype %d", dtype); return ERR_UNSUPPORTED; }
    ctx->launches++;
    return OK;
  }
  int threads = ((nvec + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  while (threads * kMaxVec < nvec) threads += 32;  // cannot trigger given the cols bound above
  if (dtype == DT_BF16) {
    ULLAVA_CHECK_CUDA(launch_pdl(norm_kernel<__nv_bfloat16, kRms>, dim3(rows), dim3(threads), 0, stream, ctx->pdl != 0,
                                 static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(w),
                                 static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(y), ldy, cols, eps,
                                 act, dst_rows));
  } else if (dtype == DT_F16) {
    ULLAVA_CHECK_CUDA(launch_pdl(norm_kernel<__half, kRms>, dim3(rows), dim3(threads), 0, stream, ctx->pdl != 0,
                                 static_cast<const __half*>(x), ldx, static_cast<const __half*>(w),
                                 static_cast<const __half*>(b), static_cast<__half*>(y), ldy, cols, eps, act,
                                 dst_rows));
  } else {
    set_last_error("norm: unsupported dtype %d", dtype);
    return ERR_UNSUPPORTED;
  }
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "norm_kernel launch");
}

int layernorm_run(Context* ctx, const void* x, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy,
                  int rows, int cols, float eps, int act, int dtype, cudaStream_t stream, const int32_t* dst_rows) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_NORM, 0.0, 4.0 * rows * cols);
  ULLAVA_REQUIRE(act == EPI_NONE || act == EPI_GELU, "layernorm: act must be NONE or GELU");
  return norm_launch<false>(ctx, x, ldx, w, b, y, ldy, rows, cols, eps, act, dtype, stream, dst_rows);
}
int rmsnorm_run(Context* ctx, const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int rows, int cols,
                float eps, int dtype, cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_NORM, 0.0, 4.0 * rows * cols);
  return norm_launch<true>(ctx, x, ldx, w, nullptr, y, ldy, rows, cols, eps, 0, dtype, stream);
}

}  // namespace ullava
