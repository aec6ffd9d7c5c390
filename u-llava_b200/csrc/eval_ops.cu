// Evaluation-side kernels of the hot path's callers: the per-sentence metrics of
// evaluation/eval_ullava.py:validate computed where the masks already are, instead of one .cpu() per sentence.
//
//   mask_iou_counts   intersectionAndUnionGPU (evaluation/tools.py:29-41) for K = 2 classes, ignore_index 255,
//                     on thresholded mask logits or integer labels: exact integer pixel counts.
//   seg_meter_update  the accumulation of validate() (evaluation/eval_ullava.py:66-86): per-image intersection /
//                     union / acc_iou sums added to running meters, in the reference's order, in fp64.
//   box_iou_diag      bbox_iou (evaluation/tools.py:13-26): torchvision box_iou(pred * 1000, gt * 1000) diagonal.
//
// All HBM-bound integer / elementwise work: coalesced 16-byte loads, warp ballots + popc, integer atomics
// (deterministic), one launch for a whole batch of masks.
#include <algorithm>

#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

// ---------------------------------------------------------------------------------------------------------------
// counts[m] = {I0, I1, U0, U1, T0, T1} for mask m.
//   output = pred label (logits > 0 -> 1 else 0, or the integer label itself), forced to `ignore` where target ==
//   ignore; I_k = #{output == k and target == k}; O_k = #{output == k}; T_k = #{target == k}; U_k = O_k + T_k - I_k
//   (torch.histc with min = 0, max = K - 1 drops every value outside [0, K - 1], i.e. the ignored pixels).
// ---------------------------------------------------------------------------------------------------------------
template <typename P>
__device__ __forceinline__ int pred_label(P v);
template <> __device__ __forceinline__ int pred_label<float>(float v) { return v > 0.f ? 1 : 0; }
template <> __device__ __forceinline__ int pred_label<int32_t>(int32_t v) { return v; }
template <> __device__ __forceinline__ int pred_label<uint8_t>(uint8_t v) { return v; }

template <typename P, typename G>
__global__ void __launch_bounds__(256)
mask_iou_counts_kernel(const P* __restrict__ pred, const G* __restrict__ target, int64_t hw, int ignore,
                       int32_t* __restrict__ counts) {
  const int m = blockIdx.y;
  const P* p = pred + static_cast<int64_t>(m) * hw;
  const G* g = target + static_cast<int64_t>(m) * hw;
  int c[6] = {0, 0, 0, 0, 0, 0};  // I0 I1 O0 O1 T0 T1 (per thread)
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < hw;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(g[i]);
    int o = pred_label<P>(p[i]);
    if (t == ignore) o = ignore;
    c[0] += (o == 0 && t == 0);
    c[1] += (o == 1 && t == 1);
    c[2] += (o == 0);
    c[3] += (o == 1);
    c[4] += (t == 0);
    c[5] += (t == 1);
  }
  __shared__ int red[6][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int v = c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    int v = 0;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    if (v) atomicAdd(counts + static_cast<int64_t>(m) * 6 + threadIdx.x, v);
  }
}

// O_k -> U_k = O_k + T_k - I_k, in place (second tiny launch so that the first can use plain integer atomics)
__global__ void mask_iou_finish_kernel(int32_t* __restrict__ counts, int n) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  int32_t* c = counts + static_cast<int64_t>(m) * 6;
  c[2] = c[2] + c[4] - c[0];
  c[3] = c[3] + c[5] - c[1];
}

template <typename P, typename G>
static int iou_counts_launch(const void* pred, const void* target, int n, int64_t hw, int ignore, int32_t* counts,
                             cudaStream_t s) {
  const int bx = static_cast<int>(std::min<int64_t>((hw + 256 * 8 - 1) / (256 * 8), 148 * 4));
  dim3 grid(bx < 1 ? 1 : bx, n);
  mask_iou_counts_kernel<P, G><<<grid, 256, 0, s>>>(static_cast<const P*>(pred), static_cast<const G*>(target), hw,
                                                    ignore, counts);
  mask_iou_finish_kernel<<<(n + 127) / 128, 128, 0, s>>>(counts, n);
  return check_cuda(cudaGetLastError(), "mask_iou_counts launch");
}

int mask_iou_counts_run(Context* ctx, const void* pred, int pred_kind, const void* target, int target_kind, int n,
                        int64_t hw, int ignore, int32_t* counts, cudaStream_t s) {
  ProfScope _ps(ctx, s, ULLAVA_PROF_GLUE, 0.0, static_cast<double>(n) * hw * 5.0);
  ULLAVA_REQUIRE(pred && target && counts, "mask_iou_counts: null pointer");
  ULLAVA_REQUIRE(n >= 0 && hw > 0 && n <= 65535, "mask_iou_counts: bad shape (n=%d hw=%lld)", n,
                 static_cast<long long>(hw));
  if (n == 0) return OK;
  ULLAVA_CHECK_CUDA(cudaMemsetAsync(counts, 0, static_cast<size_t>(n) * 6 * sizeof(int32_t), s));
  int st;
  const int key = pred_kind * 4 + target_kind;  // kinds: 0 = f32 logits (pred only), 1 = int32 labels, 2 = uint8 labels
  switch (key) {
    case 0 * 4 + 1: st = iou_counts_launch<float, int32_t>(pred, target, n, hw, ignore, counts, s); break;
    case 0 * 4 + 2: st = iou_counts_launch<float, uint8_t>(pred, target, n, hw, ignore, counts, s); break;
    case 1 * 4 + 1: st = iou_counts_launch<int32_t, int32_t>(pred, target, n, hw, ignore, counts, s); break;
    case 1 * 4 + 2: st = iou_counts_launch<int32_t, uint8_t>(pred, target, n, hw, ignore, counts, s); break;
    case 2 * 4 + 1: st = iou_counts_launch<uint8_t, int32_t>(pred, target, n, hw, ignore, counts, s); break;
    case 2 * 4 + 2: st = iou_counts_launch<uint8_t, uint8_t>(pred, target, n, hw, ignore, counts, s); break;
    default:
      set_last_error("mask_iou_counts: unsupported kinds pred=%d target=%d", pred_kind, target_kind);
      return ERR_UNSUPPORTED;
  }
  if (st == OK) ctx->launches += 2;
  return st;
}

// ---------------------------------------------------------------------------------------------------------------
// validate() accumulation, one thread, images in order (evaluation/eval_ullava.py:66-86):
//   per image: intersection = sum_i I_i, union = sum_i U_i, acc = sum_i I_i / (U_i + 1e-5), acc[U_i == 0] += 1;
//              acc /= n_masks;  inter_meter.sum += intersection; union_meter.sum += union;
//              acc_meter.sum += acc * n_masks; acc_meter.count += n_masks; inter_meter.count += 1
//   state (fp64): [0:2] inter sum, [2:4] union sum, [4:6] acc_iou sum, [6] images, [7] masks
// The reference holds these in numpy arrays promoted to float64 (python float 0.0 + integer histc counts).
// ---------------------------------------------------------------------------------------------------------------
__global__ void seg_meter_update_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ offsets,
                                        int n_images, double* __restrict__ state) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int im = 0; im < n_images; ++im) {
    const int lo = offsets[im], hi = offsets[im + 1];
    const int n = hi - lo;
    if (n <= 0) continue;  // masks_list.shape[0] == 0: the reference divides by zero and records nan; we skip
    double inter[2] = {0.0, 0.0}, uni[2] = {0.0, 0.0}, acc[2] = {0.0, 0.0};
    for (int i = lo; i < hi; ++i) {
      const int32_t* c = counts + static_cast<int64_t>(i) * 6;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double I = static_cast<double>(c[k]), U = static_cast<double>(c[2 + k]);
        inter[k] += I;
        uni[k] += U;
        acc[k] += __ddiv_rn(I, __dadd_rn(U, 1e-5));
        if (c[2 + k] == 0) acc[k] += 1.0;
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      acc[k] = __ddiv_rn(acc[k], static_cast<double>(n));
      state[k] += inter[k];
      state[2 + k] += uni[k];
      state[4 + k] += __dmul_rn(acc[k], static_cast<double>(n));
    }
    state[6] += 1.0;
    state[7] += static_cast<double>(n);
  }
}

int seg_meter_update_run(Context* ctx, const int32_t* counts, const int32_t* offsets, int n_images, double* state,
                         cudaStream_t s) {
  ULLAVA_REQUIRE(counts && offsets && state && n_images >= 0, "seg_meter_update: bad arguments");
  if (n_images == 0) return OK;
  seg_meter_update_kernel<<<1, 32, 0, s>>>(counts, offsets, n_images, state);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "seg_meter_update launch");
}

// ---------------------------------------------------------------------------------------------------------------
// bbox_iou (evaluation/tools.py:13-26) = diag(torchvision.ops.box_iou(pred * 1000, gt * 1000)), xyxy boxes.
// Rounding follows torchvision 0.26 on a tensor of dtype T: `* 1000` and `rb - lt` are evaluated in T, areas and the
// quotient in fp32 (_upcast).  iou[i] fp32; meter (optional, fp64): [0] += sum(iou > 0.5), [1] += n.
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct BoxT;
template <> struct BoxT<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ float rnd(float v) { return v; }
};
template <> struct BoxT<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};
template <> struct BoxT<__half> {
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
};

template <typename T>
__global__ void box_iou_diag_kernel(const T* __restrict__ pred, const T* __restrict__ gt, int n,
                                    float* __restrict__ iou, double* __restrict__ meter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int hit = 0;
  if (i < n) {
    float a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a[k] = BoxT<T>::rnd(__fmul_rn(BoxT<T>::ld(pred + 4 * i + k), 1000.f));
      b[k] = BoxT<T>::rnd(__fmul_rn(BoxT<T>::ld(gt + 4 * i + k), 1000.f));
    }
    const float area1 = __fmul_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]));
    const float area2 = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
    const float w = fmaxf(BoxT<T>::rnd(__fsub_rn(fminf(a[2], b[2]), fmaxf(a[0], b[0]))), 0.f);
    const float h = fmaxf(BoxT<T>::rnd(__fsub_rn(fminf(a[3], b[3]), fmaxf(a[1], b[1]))), 0.f);
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(area1, area2), inter);
    const float v = __fdiv_rn(inter, uni);
    iou[i] = v;
    hit = v > 0.5f;
  }
  if (meter) {
    const unsigned b = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(reinterpret_cast<unsigned long long*>(meter + 2), __popc(b));
  }
}

// integer hit counter (deterministic atomics) folded into the fp64 meter afterwards
__global__ void box_meter_fold_kernel(double* meter, int n) {
  const unsigned long long hits = *reinterpret_cast<unsigned long long*>(meter + 2);
  meter[0] += static_cast<double>(hits);
  meter[1] += static_cast<double>(n);
  *reinterpret_cast<unsigned long long*>(meter + 2) = 0ull;
}

int box_iou_diag_run(Context* ctx, const void* pred, const void* gt, int n, int dtype, float* iou, double* meter,
                     cudaStream_t s) {
  ULLAVA_REQUIRE(pred && gt && iou && n >= 0, "box_iou_diag: bad arguments");
  if (n == 0) return OK;
  const int grid = (n + 127) / 128;
  if (dtype == DT_F32)
    box_iou_diag_kernel<float><<<grid, 128, 0, s>>>(static_cast<const float*>(pred), static_cast<const float*>(gt), n,
                                                    iou, meter);
  else if (dtype == DT_BF16)
    box_iou_diag_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(pred),
                                                            static_cast<const __nv_bfloat16*>(gt), n, iou, meter);
  else if (dtype == DT_F16)
    box_iou_diag_kernel<__half><<<grid, 128, 0, s>>>(static_cast<const __half*>(pred), static_cast<const __half*>(gt),
                                                     n, iou, meter);
  else { set_last_error("box_iou_diag: unsupported dtype"); return ERR_UNSUPPORTED; }
  ctx->launches++;
  if (meter) {
    box_meter_fold_kernel<<<1, 1, 0, s>>>(meter, n);
    ctx->launches++;
  }
  return check_cuda(cudaGetLastError(), "box_iou_diag launch");
}

// ---------------------------------------------------------------------------------------------------------------
// Token cross-entropy of UllavaCoreForCausalLM.forward when labels are given (models/ullava_core.py:327-338:
// CrossEntropyLoss() over logits[..., :-1, :] vs labels[..., 1:], mean over the labels != ignore_index).  The
// inference callers discard it, but the reference computes it whenever labels are passed (validate() passes them),
// so the drop-in does too -- on the device, without materialising log-softmax.
//   ce_rows_kernel   : row r = (b, t), t < T - 1: loss = logsumexp(logits[b, t, :]) - logits[b, t, labels[b, t + 1]]
//   ce_reduce_kernel : fixed-order sum of the row losses / number of valid rows (NaN if none, like torch)
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct CeT { static __device__ __forceinline__ float ld(const T* p) { return T16<T>::to_f(*p); } };
template <> struct CeT<float> { static __device__ __forceinline__ float ld(const float* p) { return *p; } };

template <typename T>
__global__ void __launch_bounds__(256)
ce_rows_kernel(const T* __restrict__ logits, int64_t ld_row, int64_t ld_batch, const int64_t* __restrict__ labels,
               int64_t labels_ld, int T_len, int cols, int ignore_index, float* __restrict__ row_loss,
               float* __restrict__ row_valid) {
  const int r = blockIdx.x;
  const int b = r / (T_len - 1), t = r - b * (T_len - 1);
  const int64_t label = labels[b * labels_ld + t + 1];
  if (label == ignore_index || label < 0 || label >= cols) {
    if (threadIdx.x == 0) { row_loss[r] = 0.f; row_valid[r] = 0.f; }
    return;
  }
  const T* row = logits + b * ld_batch + t * ld_row;
  float m = -INFINITY, s = 0.f;   // online logsumexp
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = CeT<T>::ld(row + c);
    if (v > m) { s = s * __expf(m - v) + 1.f; m = v; }
    else s += __expf(v - m);
  }
  __shared__ float sm[8], ss[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    const float nm = fmaxf(m, om);
    s = (m == -INFINITY ? 0.f : s * __expf(m - nm)) + (om == -INFINITY ? 0.f : os * __expf(om - nm));
    m = nm;
  }
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = sm[0], S = ss[0];
    for (int w = 1; w < 8; ++w) {
      const float nm = fmaxf(M, sm[w]);
      S = (M == -INFINITY ? 0.f : S * __expf(M - nm)) + (sm[w] == -INFINITY ? 0.f : ss[w] * __expf(sm[w] - nm));
      M = nm;
    }
    row_loss[r] = M + __logf(S) - CeT<T>::ld(row + label);
    row_valid[r] = 1.f;
  }
}

__global__ void __launch_bounds__(1024)
ce_reduce_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_valid, int rows,
                 float* __restrict__ out) {
  __shared__ double sl[32], sv[32];
  double l = 0.0, v = 0.0;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) { l += row_loss[r]; v += row_valid[r]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { l += __shfl_xor_sync(0xffffffffu, l, o); v += __shfl_xor_sync(0xffffffffu, v, o); }
  if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = l; sv[threadIdx.x >> 5] = v; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double L = 0.0, V = 0.0;
    for (int w = 0; w < 32; ++w) { L += sl[w]; V += sv[w]; }
    out[0] = static_cast<float>(L / V);   // 0 / 0 -> NaN when every label is ignored, as torch does
    out[1] = static_cast<float>(V);
  }
}

size_t cross_entropy_scratch(int batch, int T_len) { return static_cast<size_t>(batch) * (T_len > 1 ? T_len - 1 : 0) * 8 + 256; }

int cross_entropy_run(Context* ctx, const void* logits, int logits_f32, int dtype, int64_t ld_row, int64_t ld_batch,
                      const int64_t* labels, int64_t labels_ld, int batch, int T_len, int cols, int ignore_index,
                      float* out, void* scratch, size_t scratch_bytes, cudaStream_t s) {
  ProfScope _ps(ctx, s, ULLAVA_PROF_GLUE, 0.0, static_cast<double>(batch) * T_len * cols * (logits_f32 ? 4.0 : 2.0));
  ULLAVA_REQUIRE(logits && labels && out && scratch, "cross_entropy: null pointer");
  ULLAVA_REQUIRE(batch > 0 && T_len > 1 && cols > 0, "cross_entropy: needs at least two positions");
  ULLAVA_REQUIRE(scratch_bytes >= cross_entropy_scratch(batch, T_len), "cross_entropy: scratch too small");
  const int rows = batch * (T_len - 1);
  float* row_loss = static_cast<float*>(scratch);
  float* row_valid = row_loss + rows;
  if (logits_f32)
    ce_rows_kernel<float><<<rows, 256, 0, s>>>(static_cast<const float*>(logits), ld_row, ld_batch, labels, labels_ld,
                                               T_len, cols, ignore_index, row_loss, row_valid);
  else if (dtype == DT_BF16)
    ce_rows_kernel<__nv_bfloat16><<<rows, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), ld_row, ld_batch,
                                                       labels, labels_ld, T_len, cols, ignore_index, row_loss, row_valid);
  else if (dtype == DT_F16)
    ce_rows_kernel<__half><<<rows, 256, 0, s>>>(static_cast<const __half*>(logits), ld_row, ld_batch, labels, labels_ld,
                                                T_len, cols, ignore_index, row_loss, row_valid);
  else { set_last_error("cross_entropy: unsupported dtype"); return ERR_UNSUPPORTED; }
  ce_reduce_kernel<<<1, 1024, 0, s>>>(row_loss, row_valid, rows, out);
  ctx->launches += 2;
  return check_cuda(cudaGetLastError(), "cross_entropy launch");
}

}  // namespace ullava
