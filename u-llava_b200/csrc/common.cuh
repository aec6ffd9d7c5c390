// Shared device/host helpers for the u-LLaVA sm_100a hot path.
//
// Everything here is hand-written PTX for Blackwell (sm_100a): mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (TMEM alloc / mma / commit / ld) plus a few
// conversion helpers.  No CUTLASS / CuTe is included; the bit layouts of the
// shared-memory matrix descriptor and the tcgen05 instruction descriptor follow
// the PTX ISA "tcgen05" chapter.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace ullava {

// ----------------------------------------------------------------------------
// Error plumbing: no exception ever crosses the C ABI.  Every entry point
// returns an int status and stores a message retrievable by ullava_last_error().
// ----------------------------------------------------------------------------
enum Status : int {
  OK = 0,
  ERR_BAD_ARG = -1,     // shape / alignment / null pointer
  ERR_UNSUPPORTED = -2, // dtype / head_dim / tile not compiled
  ERR_CUDA = -3,        // CUDA runtime or driver error
  ERR_ARCH = -4,        // device is not sm_100
  ERR_WORKSPACE = -5,   // workspace too small
};

void set_last_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define ULLAVA_CHECK_CUDA(expr)                                   \
  do {                                                            \
    int _st = ::ullava::check_cuda((expr), #expr);                \
    if (_st != 0) return _st;                                     \
  } while (0)

#define ULLAVA_REQUIRE(cond, ...)                                 \
  do {                                                            \
    if (!(cond)) {                                                \
      ::ullava::set_last_error(__VA_ARGS__);                      \
      return ::ullava::ERR_BAD_ARG;                               \
    }                                                             \
  } while (0)

enum DType : int { DT_BF16 = 0, DT_F16 = 1, DT_F32 = 2 };

// ----------------------------------------------------------------------------
// 16-bit type traits (every kernel is templated on __nv_bfloat16 / __half).
// ----------------------------------------------------------------------------
template <typename T> struct T16;
template <> struct T16<__nv_bfloat16> {
  using t2 = __nv_bfloat162;
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float2 to_f2(__nv_bfloat162 v) { return __bfloat1622float2(v); }
  static __device__ __forceinline__ __nv_bfloat162 from_f2(float a, float b) { return __floats2bfloat162_rn(a, b); }
  static constexpr int kUmmaFormat = 1;  // tcgen05 instruction descriptor a/b format: BF16
};
template <> struct T16<__half> {
  using t2 = __half2;
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ float2 to_f2(__half2 v) { return __half22float2(v); }
  static __device__ __forceinline__ __half2 from_f2(float a, float b) { return __floats2half2_rn(a, b); }
  static constexpr int kUmmaFormat = 0;  // F16
};

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  typename T16<T>::t2 v = T16<T>::from_f2(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  typename T16<T>::t2 v = *reinterpret_cast<typename T16<T>::t2*>(&u);
  return T16<T>::to_f2(v);
}

#ifdef __CUDACC__
// ----------------------------------------------------------------------------
// Small PTX wrappers
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  // make barrier inits visible to the async proxy (TMA / tcgen05.commit) and to the cluster
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA ----------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tile load global -> shared::cta; c0 = innermost coordinate (elements), c1 = row.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 4D tile store shared::cta -> global (bulk async-group of the issuing thread); out-of-range rows / columns are clipped.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread's stores, all but the newest N groups, have finished READING shared memory
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// Same with an L2 cache-policy hint (createpolicy-style 64-bit constant).
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "l"(hint)
      : "memory");
}
// Prefetch of a 2D tile into L2 only (no shared-memory destination, no barrier): a hint, never a fault.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
// 2-CTA (cta_group::2) flavour: executed by both CTAs of a pair, the transaction
// bytes are signalled on the LEADER CTA's barrier (peer bit of the address cleared).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                     uint64_t hint) {
  uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
// 4D tile load (innermost coordinate first).  Out-of-bounds elements are zero-filled.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;

// ---- programmatic dependent launch ------------------------------------------------------------
// launch_dependents: the next kernel of the stream, IF it was launched with the programmatic-serialization attribute,
// may start as soon as every CTA of this grid has executed this (or exited).  Only gemm_stream_kernel is launched that
// way; everything it does before its own griddepcontrol.wait is independent of this grid (weight prefetch).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- cluster ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------
template <int kCtaGroup>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  if constexpr (kCtaGroup == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int kCtaGroup>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (kCtaGroup == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  } else {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], 16-bit inputs, fp32 accumulate.
template <int kCtaGroup>
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  if constexpr (kCtaGroup == 1) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// D[tmem] (+)= A[tmem] * B[smem]: A is a 128-lane x K tile of packed 16-bit pairs in tensor memory
// (lane = row, 32-bit column c holds elements 2c, 2c+1), always K-major.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Make `bar` observe completion of all previously issued tcgen05.mma of this thread.
// (implicitly performs tcgen05.fence::before_thread_sync)
template <int kCtaGroup>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  if constexpr (kCtaGroup == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
  } else {
    // arrive on the barrier at this offset in BOTH CTAs of the pair
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(0x3))
        : "memory");
  }
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane
// (32*(warp%4)+i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// store 16 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major tile whose rows are exactly 128
// bytes (64 x 16-bit) and which was written by TMA with SWIZZLE_128B:
//   bits [0,14)  start address >> 4
//   bits [16,30) leading byte offset >> 4   (unused for swizzled K-major; 1)
//   bits [32,46) stride byte offset >> 4    (8 rows * 128 B = 1024 B between 8-row groups)
//   bits [46,48) descriptor version = 1 on sm_100
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// General form: layout 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B, 0 = none.
//   K-major  (swizzled): SBO = bytes between 8-row groups, LBO unused.
//   MN-major (swizzled): LBO = bytes between swizzle-width atoms along M/N, SBO = bytes between 8-element K groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// tcgen05 instruction descriptor, kind::f16, fp32 accumulate, A and B K-major:
//   [4,6) D format (1 = F32)  [7,10) A format  [10,13) B format (0 = F16, 1 = BF16)
//   [15] A major (0 = K)  [16] B major (0 = K)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16(int ab_format, int umma_m, int umma_n) {
  return (1u << 4) | (static_cast<uint32_t>(ab_format) << 7) | (static_cast<uint32_t>(ab_format) << 10) |
         (static_cast<uint32_t>(umma_n >> 3) << 17) | (static_cast<uint32_t>(umma_m >> 4) << 24);
}

// ---- packed fp32 x2 arithmetic (sm_100: one issue slot for two lanes of work) -----------------------------------------
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 2^x for two values on the FMA / ALU pipes instead of the MUFU (16 ex2 / clk / SM on B200 -- the pipe that bounds an
// attention tile: 128 x 128 exponentials = 1024 cycles against 512 cycles of tcgen05 work at head_dim 64).
// x = n + f, n = floor(x) from a round-down add of 1.5 * 2^23 (n lands in the low mantissa bits), 2^f from a minimax
// polynomial with p(0) = 1 (degree 3: relative error 8.6e-5, below the rounding of a bf16 P; degree 4: 3.0e-6, below
// fp16's), times the exact power of two (n + 127) << 23.  x <= -127 (masked keys are -inf) gives exactly 0; x up to
// +127 is fine (the lazy rescale keeps x <= 8).
template <int DEG>
__device__ __forceinline__ uint64_t ex2_fma2(float x0, float x1) {
  static_assert(DEG == 3 || DEG == 4, "ex2_fma2: degree 3 or 4");
  const uint64_t x = pk2(fmaxf(x0, -127.f), fmaxf(x1, -127.f));
  const uint64_t magic = pk2(12582912.f, 12582912.f);
  uint64_t r, fl, f;
  asm("add.rm.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(magic));
  asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(fl) : "l"(r), "l"(magic));
  asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(f) : "l"(x), "l"(fl));
  uint64_t p;
  if constexpr (DEG == 3) {
    p = ffma2(pk2(0.07706641405820847f, 0.07706641405820847f), f, pk2(0.22764568030834198f, 0.22764568030834198f));
    p = ffma2(p, f, pk2(0.6951166391372681f, 0.6951166391372681f));
  } else {
    p = ffma2(pk2(0.013426647521555424f, 0.013426647521555424f), f, pk2(0.05224253237247467f, 0.05224253237247467f));
    p = ffma2(p, f, pk2(0.2412801831960678f, 0.2412801831960678f));
    p = ffma2(p, f, pk2(0.6930448412895203f, 0.6930448412895203f));
  }
  p = ffma2(p, f, pk2(1.f, 1.f));
  uint32_t n0, n1;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(n0), "=r"(n1) : "l"(r));
  const uint64_t scale = pk2(__uint_as_float((n0 << 23) + 0x3F800000u), __uint_as_float((n1 << 23) + 0x3F800000u));
  return fmul2(p, scale);
}

// ---- fast activations (MUFU based) ------------------------------------------------------------
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x * sigmoid(k x) = x / (1 + 2^(-k log2(e) x)): one MUFU.EX2 + one MUFU.RCP.  x -> -inf gives x * 0.
__device__ __forceinline__ float fast_x_sigmoid(float x, float k_log2e) {
  return x * rcp_approx(1.f + ex2_approx(-k_log2e * x));
}
// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) (nn.GELU, approximate='none'), erf by Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7):  with a = |x|, s = a / sqrt 2, t = 1 / (1 + p s), E = poly(t) exp(-s^2) = 1 - erf(s):
//   GELU(x) = max(x, 0) - 0.5 a E.          Max abs error 3.3e-7 over [-12, 12] (checked against math.erf).
__device__ __forceinline__ float fast_gelu(float x) {
  const float a = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678f, a, 1.f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = ex2_approx(a * a * -0.72134752f);  // exp(-x^2 / 2)
  return fmaf(-0.5f * a, poly * t * e, fmaxf(x, 0.f));
}

// ---- misc math ----------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif  // __CUDACC__

// ----------------------------------------------------------------------------
// Host-side: TMA descriptor encode through the driver entry point (the library
// does not link libcuda, so it loads on a machine without a driver).
// ----------------------------------------------------------------------------
int encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                   uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, bool swizzle128);
// 4D 16-bit tensor (dims innermost first, strides in bytes for dims 1..3), box = box0 x box1 x 1 x 1,
// swizzle_bytes in {0, 32, 64, 128}.
int encode_tmap_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                   uint32_t box0, uint32_t box1, int swizzle_bytes);

int device_sm_count();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: remember what was configured per device (a
// process may drive several GPUs through several contexts), keyed by the device current at launch time.
struct SmemOptIn {
  int bytes[64] = {0};
};
template <typename Kern>
inline int ensure_dynamic_smem(Kern kern, int bytes, SmemOptIn& state) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = -1;
  if (dev >= 0 && bytes <= state.bytes[dev]) return OK;
  const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  if (dev >= 0) state.bytes[dev] = bytes;
  return OK;
}


#ifdef __CUDACC__
// Launch with the programmatic-stream-serialization attribute: the kernel may start while its predecessor in the
// stream is still running, once every CTA of the predecessor has executed griddepcontrol.launch_dependents (or
// exited).  The kernel MUST execute griddepcontrol.wait before it touches global memory that the predecessor (or
// anything before it) writes or reads-then-expects-unchanged.  With `pdl == false` this is a plain launch.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace ullava
