// Image preprocessing on the GPU: from the decoded uint8 HWC image to the two model inputs, with the
// arithmetic of the host libraries the reference calls, so that the tensors are bit-identical to the reference's.
//
//   resize_u8        PIL.Image.resize (Pillow 8bpc resampler: ImagingResample, src/libImaging/Resample.c): separable,
//                    antialiased, integer arithmetic.  Coefficients are computed on the host in double exactly as
//                    precompute_coeffs / normalize_coeffs_8bpc do (22 fractional bits), the two passes are integer
//                    MACs + clip8 on the device.  Call sites: CLIPImageProcessor.resize (BICUBIC, shortest edge;
//                    dataset/processors/clip_processor.py:93) and ResizeLongestSide.apply_image (BILINEAR;
//                    models/segment_anything/utils/transforms.py:29-37, dataset/tools/mask_toolbox.py:27).
//   clip_preprocess  optional white square padding (clip_processor.py:36-79) is done by the caller; this kernel does
//                    center crop -> x * (1/255) (double product rounded to fp32) -> (x - mean) / std in fp32 ->
//                    16-bit CHW  (CLIPImageProcessor.preprocess of transformers 4.29.1: rescale, normalize;
//                    evaluation/tools.py:55-67 casts to the model dtype).
//   sam_preprocess   (x - mean) / std in fp32, zero pad to S x S, 16-bit CHW (mask_toolbox.py:15-25).
// HBM-bound byte work: one thread per output pixel, channels together, coalesced along x.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

static constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow: PRECISION_BITS

// ---- host: Pillow's precompute_coeffs + normalize_coeffs_8bpc -------------------------------------------------
static inline double pil_bilinear(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}
static inline double pil_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// bounds: [out][2] = (first input index, taps); kk: [out][ksize] fixed-point taps
static int pil_coeffs(int in_size, int out_size, int filter, std::vector<int32_t>& bounds, std::vector<int32_t>& kk) {
  const double fsupport = filter == 1 ? 2.0 : 1.0;
  double (*f)(double) = filter == 1 ? pil_bicubic : pil_bilinear;
  const float in0 = 0.f, in1 = static_cast<float>(in_size);
  double scale = static_cast<double>(in1 - in0) / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = fsupport * filterscale;
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  bounds.assign(static_cast<size_t>(out_size) * 2, 0);
  kk.assign(static_cast<size_t>(out_size) * ksize, 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = in0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = f((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < xmax; ++x) {
      const double v = k[x];
      kk[static_cast<size_t>(xx) * ksize + x] = v < 0 ? static_cast<int32_t>(-0.5 + v * (1 << kPrecisionBits))
                                                      : static_cast<int32_t>(0.5 + v * (1 << kPrecisionBits));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}

__device__ __forceinline__ uint8_t pil_clip8(int v) {
  v >>= kPrecisionBits;  // arithmetic shift, like the C source
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal: out[y][xx][c] = clip8(2^21 + sum_x in[y][xmin + x][c] * k[xx][x])
__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ src, int h, int w, uint8_t* __restrict__ dst, int ow,
                  const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (xx >= ow) return;
  const int xmin = bounds[2 * xx], taps = bounds[2 * xx + 1];
  const int32_t* k = kk + static_cast<int64_t>(xx) * ksize;
  const uint8_t* row = src + (static_cast<int64_t>(y) * w + xmin) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < taps; ++x) {
    const int kv = k[x];
    s0 += row[3 * x] * kv;
    s1 += row[3 * x + 1] * kv;
    s2 += row[3 * x + 2] * kv;
  }
  uint8_t* o = dst + (static_cast<int64_t>(y) * ow + xx) * 3;
  o[0] = pil_clip8(s0); o[1] = pil_clip8(s1); o[2] = pil_clip8(s2);
}

// vertical: out[yy][x][c] = clip8(2^21 + sum_y in[ymin + y][x][c] * k[yy][y]); one thread per output byte
__global__ void __launch_bounds__(256)
resample_v_kernel(const uint8_t* __restrict__ src, int row_bytes, uint8_t* __restrict__ dst,
                  const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y;
  if (i >= row_bytes) return;
  const int ymin = bounds[2 * yy], taps = bounds[2 * yy + 1];
  const int32_t* k = kk + static_cast<int64_t>(yy) * ksize;
  const uint8_t* p = src + static_cast<int64_t>(ymin) * row_bytes + i;
  int s = 1 << (kPrecisionBits - 1);
  for (int y = 0; y < taps; ++y) s += p[static_cast<int64_t>(y) * row_bytes] * k[y];
  dst[static_cast<int64_t>(yy) * row_bytes + i] = pil_clip8(s);
}

// Host-only entry: the tap table the device passes consume (for callers that cache it, and for CPU tests of the
// Pillow restatement).  Returns ksize; bounds [out][2], kk [out][ksize] are written when the pointers are given and
// `capacity` (entries of kk) is large enough, otherwise only ksize is returned.
int resample_coeffs_host(int in_size, int out_size, int filter, int32_t* bounds_out, int32_t* kk_out, size_t capacity) {
  if (in_size <= 0 || out_size <= 0 || (filter != 0 && filter != 1)) return -1;
  std::vector<int32_t> bounds, kk;
  const int ksize = pil_coeffs(in_size, out_size, filter, bounds, kk);
  if (bounds_out && kk_out && capacity >= kk.size()) {
    std::copy(bounds.begin(), bounds.end(), bounds_out);
    std::copy(kk.begin(), kk.end(), kk_out);
  }
  return ksize;
}

static inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

size_t resize_u8_scratch(int h, int w, int oh, int ow) {
  // intermediate [h, ow, 3] + two coefficient tables (ksize <= 2 * ceil(2 * max_scale) + 1)
  auto table = [](int in, int out) {
    double fs = static_cast<double>(in) / out;
    if (fs < 1.0) fs = 1.0;
    const size_t ksize = static_cast<size_t>(std::ceil(2.0 * fs)) * 2 + 1;
    return up256(static_cast<size_t>(out) * 2 * 4) + up256(static_cast<size_t>(out) * ksize * 4);
  };
  return up256(static_cast<size_t>(h) * ow * 3) + table(w, ow) + table(h, oh) + 1024;
}

int resize_u8_run(Context* ctx, const uint8_t* src, int h, int w, uint8_t* dst, int oh, int ow, int filter,
                  void* scratch, size_t scratch_bytes, cudaStream_t s) {
  ProfScope _ps(ctx, s, ULLAVA_PROF_GLUE, 0.0, 3.0 * (static_cast<double>(h) * w + 2.0 * h * ow + oh * ow));
  ULLAVA_REQUIRE(src && dst && scratch, "resize_u8: null pointer");
  ULLAVA_REQUIRE(h > 0 && w > 0 && oh > 0 && ow > 0 && h <= 65535 && oh <= 65535, "resize_u8: bad geometry");
  ULLAVA_REQUIRE(filter == 0 || filter == 1, "resize_u8: filter must be 0 (bilinear) or 1 (bicubic)");
  ULLAVA_REQUIRE(scratch_bytes >= resize_u8_scratch(h, w, oh, ow), "resize_u8: scratch too small");
  uint8_t* sc = static_cast<uint8_t*>(scratch);
  uint8_t* mid = sc;
  size_t off = up256(static_cast<size_t>(h) * ow * 3);
  const bool need_h = ow != w, need_v = oh != h;
  const uint8_t* vin = src;
  if (!need_h && !need_v) {
    ULLAVA_CHECK_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(h) * w * 3, cudaMemcpyDeviceToDevice, s));
    return OK;
  }
  std::vector<int32_t> bounds, kk;
  if (need_h) {
    const int ksize = pil_coeffs(w, ow, filter, bounds, kk);
    int32_t* d_b = reinterpret_cast<int32_t*>(sc + off);
    off += up256(bounds.size() * 4);
    int32_t* d_k = reinterpret_cast<int32_t*>(sc + off);
    off += up256(kk.size() * 4);
    ULLAVA_CHECK_CUDA(cudaMemcpyAsync(d_b, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice, s));
    ULLAVA_CHECK_CUDA(cudaMemcpyAsync(d_k, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice, s));
    uint8_t* hout = need_v ? mid : dst;
    dim3 grid((ow + 255) / 256, h);
    resample_h_kernel<<<grid, 256, 0, s>>>(src, h, w, hout, ow, d_b, d_k, ksize);
    ctx->launches++;
    vin = hout;
  }
  if (need_v) {
    const int ksize = pil_coeffs(h, oh, filter, bounds, kk);
    int32_t* d_b = reinterpret_cast<int32_t*>(sc + off);
    off += up256(bounds.size() * 4);
    int32_t* d_k = reinterpret_cast<int32_t*>(sc + off);
    off += up256(kk.size() * 4);
    ULLAVA_CHECK_CUDA(cudaMemcpyAsync(d_b, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice, s));
    ULLAVA_CHECK_CUDA(cudaMemcpyAsync(d_k, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice, s));
    const int row_bytes = ow * 3;
    dim3 grid((row_bytes + 255) / 256, oh);
    resample_v_kernel<<<grid, 256, 0, s>>>(vin, row_bytes, dst, d_b, d_k, ksize);
    ctx->launches++;
  }
  return check_cuda(cudaGetLastError(), "resize_u8 launch");
}

// ---- normalisation kernels -------------------------------------------------------------------------------------
// output element: 16-bit (the model dtype, one rounding of the fp32 value) or fp32 (what the reference's processors
// return before their callers do .cuda().to(dtype))
template <typename T> struct PxOut { static __device__ __forceinline__ T cvt(float v) { return T16<T>::from_f(v); } };
template <> struct PxOut<float> { static __device__ __forceinline__ float cvt(float v) { return v; } };

template <typename T>
__global__ void __launch_bounds__(256)
clip_preprocess_kernel(const uint8_t* __restrict__ src, int w, int top, int left, int size, float m0, float m1,
                       float m2, float s0, float s1, float s2, double rescale, T* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= size) return;
  const uint8_t* p = src + (static_cast<int64_t>(y + top) * w + (x + left)) * 3;
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __double2float_rn(__dmul_rn(static_cast<double>(p[c]), rescale));
    out[(static_cast<int64_t>(c) * size + y) * size + x] = PxOut<T>::cvt(__fdiv_rn(__fsub_rn(v, mean[c]), sd[c]));
  }
}

int clip_preprocess_run(Context* ctx, const uint8_t* src, int h, int w, int top, int left, int size,
                        const float* mean, const float* stdv, double rescale, void* out, int dtype, cudaStream_t s) {
  ProfScope _ps(ctx, s, ULLAVA_PROF_GLUE, 0.0, 3.0 * size * size * 3.0);
  ULLAVA_REQUIRE(src && out && mean && stdv, "clip_preprocess: null pointer");
  ULLAVA_REQUIRE(size > 0 && top >= 0 && left >= 0 && top + size <= h && left + size <= w,
                 "clip_preprocess: crop %d+%d x %d+%d outside the %d x %d image", top, size, left, size, h, w);
  dim3 grid((size + 255) / 256, size);
  if (dtype == DT_BF16)
    clip_preprocess_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(src, w, top, left, size, mean[0], mean[1], mean[2],
                                                               stdv[0], stdv[1], stdv[2], rescale,
                                                               static_cast<__nv_bfloat16*>(out));
  else if (dtype == DT_F16)
    clip_preprocess_kernel<__half><<<grid, 256, 0, s>>>(src, w, top, left, size, mean[0], mean[1], mean[2], stdv[0],
                                                        stdv[1], stdv[2], rescale, static_cast<__half*>(out));
  else if (dtype == DT_F32)
    clip_preprocess_kernel<float><<<grid, 256, 0, s>>>(src, w, top, left, size, mean[0], mean[1], mean[2], stdv[0],
                                                       stdv[1], stdv[2], rescale, static_cast<float*>(out));
  else { set_last_error("clip_preprocess: unsupported dtype"); return ERR_UNSUPPORTED; }
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "clip_preprocess launch");
}

template <typename T>
__global__ void __launch_bounds__(256)
sam_preprocess_kernel(const uint8_t* __restrict__ src, int h, int w, int S, float m0, float m1, float m2, float s0,
                      float s1, float s2, T* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= S) return;
  const bool inside = x < w && y < h;
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = 0.f;  // F.pad after the normalisation: the border is exactly zero
    if (inside) v = __fdiv_rn(__fsub_rn(static_cast<float>(src[(static_cast<int64_t>(y) * w + x) * 3 + c]), mean[c]), sd[c]);
    out[(static_cast<int64_t>(c) * S + y) * S + x] = PxOut<T>::cvt(v);
  }
}

int sam_preprocess_run(Context* ctx, const uint8_t* src, int h, int w, int S, const float* mean, const float* stdv,
                       void* out, int dtype, cudaStream_t s) {
  ProfScope _ps(ctx, s, ULLAVA_PROF_GLUE, 0.0, 3.0 * (static_cast<double>(h) * w + 2.0 * S * S));
  ULLAVA_REQUIRE(src && out && mean && stdv, "sam_preprocess: null pointer");
  ULLAVA_REQUIRE(h > 0 && w > 0 && h <= S && w <= S, "sam_preprocess: %d x %d does not fit %d x %d", h, w, S, S);
  dim3 grid((S + 255) / 256, S);
  if (dtype == DT_BF16)
    sam_preprocess_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(src, h, w, S, mean[0], mean[1], mean[2], stdv[0], stdv[1],
                                                              stdv[2], static_cast<__nv_bfloat16*>(out));
  else if (dtype == DT_F16)
    sam_preprocess_kernel<__half><<<grid, 256, 0, s>>>(src, h, w, S, mean[0], mean[1], mean[2], stdv[0], stdv[1],
                                                       stdv[2], static_cast<__half*>(out));
  else if (dtype == DT_F32)
    sam_preprocess_kernel<float><<<grid, 256, 0, s>>>(src, h, w, S, mean[0], mean[1], mean[2], stdv[0], stdv[1],
                                                      stdv[2], static_cast<float*>(out));
  else { set_last_error("sam_preprocess: unsupported dtype"); return ERR_UNSUPPORTED; }
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "sam_preprocess launch");
}

}  // namespace ullava
