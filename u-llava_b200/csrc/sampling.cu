// Device-side temperature / top-k / top-p sampling step of the generate loop.
//
// Reference: UllavaForCausalLM.evaluate passes do_sample = temperature > 0 (default 0.2), top_p to
// GenerationMixin.generate (models/ullava.py:343-362); GenerationConfig's default top_k = 50 applies as well, i.e. per step
//   scores = logits / temperature                                  (TemperatureLogitsWarper)
//   drop scores < the k-th largest score (ties kept)               (TopKLogitsWarper, top_k = 50 unless the caller says otherwise)
//   sort ascending, cum = cumsum(softmax(sorted)); drop cum <= 1 - top_p, always keep the last   (TopPLogitsWarper)
//   next ~ Categorical(softmax(filtered scores))                   (torch.multinomial)
// torch.multinomial's random stream cannot be reproduced by another kernel, so parity here is distributional: this
// kernel draws by inverse CDF in vocabulary order from the SAME filtered, renormalised distribution with one
// caller-supplied uniform number per row (oracle: oracle/ullava_oracle.py:sample_inverse_cdf, same arithmetic in
// fp64; tests compare the kept set, the per-token probabilities and the drawn ids).
//
// One CTA per row, the row's exp() values live in shared memory (V = 32 011 -> 125 KB): logits are read once.
//   1. m = max(logits); e_i = exp((l_i - m) / T); Z = sum e_i
//   2. top-k: the k-th largest e is the largest float t with #{e_i >= t} >= k: bisection on the float bit pattern of t
//      (non-negative floats order like their bits), 31 block-wide counts; Z_k = sum of the survivors.
//      top-p over the survivors: the kept set is {i : S(e_i) > (1 - top_p) * Z_k}, S(x) = sum of all surviving
//      e_j <= x (what the ascending cumsum tests, up to ties).  It is an upper set in e, found by the same kind of
//      bisection with masked sums.  top_p = 0 keeps the maximum only (min_tokens_to_keep = 1).
//   3. target = u * Z_kept; block-wide exclusive scan of per-thread chunk sums (contiguous chunks: index order), the
//      owning thread walks its chunk.
// Bookkeeping (finished / eos / pad, sequence buffer, hidden-state copy, position) is that of greedy_step_kernel.
#include "common.cuh"
#include "ullava_internal.h"

namespace ullava {

static constexpr int SMP_THREADS = 1024;

__device__ __forceinline__ float block_sum_f(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // red may still be read by the previous call
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = l < (SMP_THREADS / 32) ? red[l] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;  // every thread holds the block total (same reduction tree everywhere: deterministic)
}

__device__ __forceinline__ float block_max_f(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = l < (SMP_THREADS / 32) ? red[l] : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  return t;
}

__global__ void __launch_bounds__(SMP_THREADS)
sample_step_kernel(const float* __restrict__ logits, int64_t ld, int cols, float inv_temp, float top_p, int top_k,
                   const float* __restrict__ uniforms, int64_t uni_ld, int64_t* __restrict__ cur_ids,
                   int64_t* __restrict__ seqs, int64_t seqs_ld, const uint16_t* __restrict__ final_h,
                   uint16_t* __restrict__ hid_buf, int64_t hid_bs, int hdim, uint8_t* __restrict__ finished,
                   int eos_id, int pad_id, const int32_t* __restrict__ pos_dev, float* __restrict__ probs_out) {
  extern __shared__ float ev[];  // [cols] exp values
  __shared__ float red[32];
  __shared__ float scan[SMP_THREADS / 32];
  __shared__ int pick;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int pos = pos_dev ? *pos_dev : 0;
  if (hid_buf) {
    const uint4* src = reinterpret_cast<const uint4*>(final_h + static_cast<int64_t>(b) * hdim);
    uint4* dst = reinterpret_cast<uint4*>(hid_buf + b * hid_bs + static_cast<int64_t>(pos) * hdim);
    for (int v = tid; v < hdim / 8; v += SMP_THREADS) dst[v] = src[v];
  }
  const float* row = logits + b * ld;
  // contiguous chunk per thread, so that a scan over threads is a scan over the vocabulary in index order
  const int chunk = (cols + SMP_THREADS - 1) / SMP_THREADS;
  const int c0 = min(tid * chunk, cols), c1 = min(c0 + chunk, cols);

  float mx = -INFINITY;
  for (int c = tid; c < cols; c += SMP_THREADS) {
    const float v = row[c];
    ev[c] = v;
    mx = fmaxf(mx, v);
  }
  mx = block_max_f(mx, red);
  float z = 0.f;
  for (int c = tid; c < cols; c += SMP_THREADS) {
    const float e = __expf((ev[c] - mx) * inv_temp);
    ev[c] = e;
    z += e;
  }
  z = block_sum_f(z, red);

  // ---- top-k threshold: largest float tk with #{e >= tk} >= k (the k-th largest value; ties survive like HF's
  //      `scores < kth` test) ----
  uint32_t thr_bits = 0u;  // keep everything
  if (top_k > 0 && top_k < cols) {
    uint32_t lo = 0u, hi = __float_as_uint(1.0f);  // count(e >= 0) = cols >= k, count(e >= 1.0) >= 1 (the maximum)
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo + 1) >> 1);
      const float t = __uint_as_float(mid);
      float n = 0.f;
      for (int c = tid; c < cols; c += SMP_THREADS) n += ev[c] >= t ? 1.f : 0.f;
      n = block_sum_f(n, red);  // exact: counts stay below 2^24
      if (n >= static_cast<float>(top_k)) lo = mid; else hi = mid - 1;
    }
    thr_bits = lo;
    const float tk = __uint_as_float(lo);
    float zk = 0.f;
    for (int c = tid; c < cols; c += SMP_THREADS) zk += ev[c] >= tk ? ev[c] : 0.f;
    z = block_sum_f(zk, red);
  }
  // ---- top-p threshold over the survivors: smallest float thr with S(thr) = sum{tk <= e <= thr} > (1 - top_p) * z ----
  if (top_p >= 0.f && top_p < 1.f) {
    const float tk = __uint_as_float(thr_bits);
    const float cut = (1.f - top_p) * z;
    uint32_t lo = thr_bits, hi = __float_as_uint(1.0f);  // e in [0, 1]; S(1.0) = z > cut unless top_p = 0 (-> thr = 1.0)
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      const float t = __uint_as_float(mid);
      float s = 0.f;
      for (int c = tid; c < cols; c += SMP_THREADS) s += (ev[c] <= t && ev[c] >= tk) ? ev[c] : 0.f;
      s = block_sum_f(s, red);
      if (s > cut) hi = mid; else lo = mid + 1;
    }
    thr_bits = lo;
  }
  const float thr = __uint_as_float(thr_bits);

  // ---- inverse CDF over the kept set, in index order ----
  float local = 0.f;
  for (int c = c0; c < c1; ++c) local += ev[c] >= thr ? ev[c] : 0.f;
  // block exclusive scan of `local`
  float incl = local;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) scan[warp] = incl;
  if (tid == 0) pick = -1;
  __syncthreads();
  float warp_off = 0.f, total = 0.f;
  for (int w = 0; w < SMP_THREADS / 32; ++w) {
    const float s = scan[w];
    if (w < warp) warp_off += s;
    total += s;
  }
  const float excl = warp_off + incl - local;
  const float u = uniforms ? uniforms[static_cast<int64_t>(pos) * uni_ld + b] : 0.5f;
  const float target = u * total;
  if (probs_out) {
    const float inv = 1.f / total;
    for (int c = c0; c < c1; ++c) probs_out[static_cast<int64_t>(b) * cols + c] = ev[c] >= thr ? ev[c] * inv : 0.f;
  }
  if (local > 0.f && target >= excl && target < excl + local) {
    float run = excl;
    int chosen = -1;
    for (int c = c0; c < c1; ++c) {
      if (ev[c] >= thr) {
        run += ev[c];
        chosen = c;  // last kept index of the chunk if rounding leaves target >= run at the end
        if (target < run) break;
      }
    }
    pick = chosen;  // intervals of different threads are disjoint: at most one writer
  }
  __syncthreads();
  if (tid == 0) {
    int chosen = pick;
    if (chosen < 0) {  // target == total after rounding (u -> 1): the last kept token
      for (int c = cols - 1; c >= 0; --c)
        if (ev[c] >= thr) { chosen = c; break; }
      if (chosen < 0) chosen = 0;
    }
    int64_t nxt = chosen;
    if (finished) {
      if (finished[b]) nxt = pad_id;
      else if (eos_id >= 0 && nxt == eos_id) finished[b] = 1;
    }
    cur_ids[b] = nxt;
    if (seqs) seqs[b * seqs_ld + pos + 1] = nxt;
  }
}

__global__ void sample_advance_pos_kernel(int32_t* pos) { *pos += 1; }

int sample_step_run(Context* ctx, const float* logits, int64_t ld, int rows, int cols, float temperature, float top_p,
                    int top_k, const float* uniforms, int64_t uni_ld, int64_t* cur_ids, int64_t* seqs, int64_t seqs_ld,
                    const void* final_h, void* hid_buf, int64_t hid_bs, int hdim, uint8_t* finished, int eos_id,
                    int pad_id, int32_t* pos_dev, float* probs_out, cudaStream_t stream) {
  ProfScope _ps(ctx, stream, ULLAVA_PROF_GLUE, 0.0, 4.0 * rows * cols);
  ULLAVA_REQUIRE(logits && cur_ids && cols > 0 && hdim % 8 == 0, "sample_step: bad arguments");
  ULLAVA_REQUIRE(temperature > 0.f, "sample_step: temperature must be > 0 (greedy decoding is ullava_greedy_step)");
  ULLAVA_REQUIRE(top_p >= 0.f && top_p <= 1.f, "sample_step: top_p must be in [0, 1] (1 = no filtering, 0 = top-1 only)");
  ULLAVA_REQUIRE(top_k >= 0, "sample_step: top_k must be >= 0 (0 = no filtering)");
  const size_t smem = static_cast<size_t>(cols) * sizeof(float);
  ULLAVA_REQUIRE(smem <= 200 * 1024, "sample_step: vocabulary of %d does not fit the shared-memory row buffer", cols);
  if (rows == 0) return OK;
  // per device attribute, cheap: set on every call (a process may drive several devices)
  ULLAVA_CHECK_CUDA(cudaFuncSetAttribute(sample_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  sample_step_kernel<<<rows, SMP_THREADS, smem, stream>>>(
      logits, ld, cols, 1.f / temperature, top_p, top_k, uniforms, uni_ld, cur_ids, seqs, seqs_ld,
      static_cast<const uint16_t*>(final_h), static_cast<uint16_t*>(hid_buf), hid_bs, hdim, finished, eos_id, pad_id,
      pos_dev, probs_out);
  ctx->launches++;
  if (pos_dev) {
    sample_advance_pos_kernel<<<1, 1, 0, stream>>>(pos_dev);
    ctx->launches++;
  }
  return check_cuda(cudaGetLastError(), "sample_step launch");
}

}  // namespace ullava
