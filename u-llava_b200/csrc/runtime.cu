// Context, error reporting and TMA-descriptor encoding for libullava_sm100.so.
#include <cstdlib>
#include "common.cuh"
#include "ullava_internal.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace ullava {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return OK;
  set_last_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return ERR_CUDA;
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_encode = nullptr;
static std::once_flag g_encode_once;

static void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<EncodeFn>(fn);
}

int encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                   uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, bool swizzle128) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) {
    set_last_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return ERR_CUDA;
  }
  if (elem_bytes != 2) {
    set_last_error("encode_tmap_2d: only 16-bit elements supported");
    return ERR_UNSUPPORTED;
  }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  // BF16 and F16 are both opaque 16-bit payloads for a tiled copy; use the 16-bit uint type so one
  // descriptor serves both.
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu stride=%llu box=%ux%u",
                   static_cast<int>(r), base, (unsigned long long)inner, (unsigned long long)outer,
                   (unsigned long long)row_stride_bytes, box_inner, box_outer);
    return ERR_CUDA;
  }
  return OK;
}

int encode_tmap_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                   uint32_t box0, uint32_t box1, int swizzle_bytes) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) {
    set_last_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return ERR_CUDA;
  }
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstride[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t box[4] = {box0, box1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(4d) failed (%d): base=%p dims=%llu,%llu,%llu,%llu strides=%llu,%llu,%llu box=%ux%u sw=%d",
                   static_cast<int>(r), base, (unsigned long long)dims[0], (unsigned long long)dims[1],
                   (unsigned long long)dims[2], (unsigned long long)dims[3], (unsigned long long)strides_bytes[0],
                   (unsigned long long)strides_bytes[1], (unsigned long long)strides_bytes[2], box0, box1, swizzle_bytes);
    return ERR_CUDA;
  }
  return OK;
}

ProfScope::ProfScope(Context* c, cudaStream_t s, int cls, double flops, double bytes)
    : ctx(c), stream(s), idx(-1), launches0(c ? c->launches : 0) {
  if (!c || !c->prof_on) return;
  ullava_prof_rec r{};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  r.cls = cls; r.flops = flops; r.bytes = bytes; r.launches = 0;
  cudaEventRecord(r.a, s);
  c->prof.push_back(r);
  idx = static_cast<int>(c->prof.size()) - 1;
}

ProfScope::~ProfScope() {
  if (idx < 0) return;
  ullava_prof_rec& r = ctx->prof[idx];
  r.launches = static_cast<int>(ctx->launches - launches0);
  cudaEventRecord(r.b, stream);
}

}  // namespace ullava

using namespace ullava;

extern "C" {

int ullava_abi_version(void) { return ULLAVA_ABI_VERSION; }

const char* ullava_last_error(void) { return ullava::last_error(); }

int ullava_create(int device, ullava_ctx** out) {
  if (!out) {
    set_last_error("ullava_create: out is NULL");
    return ERR_BAD_ARG;
  }
  *out = nullptr;
  int n = 0;
  ULLAVA_CHECK_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) {
    set_last_error("ullava_create: device %d out of range (%d devices)", device, n);
    return ERR_BAD_ARG;
  }
  cudaDeviceProp prop;
  ULLAVA_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_last_error("ullava_create: device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major,
                   prop.minor);
    return ERR_ARCH;
  }
  ullava_ctx* c = new ullava_ctx();
  c->device = device;
  c->sm_count = c->device_sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("ULLAVA_PREFETCH_UNITS")) {
    const int v = atoi(e);
    if (v >= 0 && v <= 256) c->prefetch_units = v;
  }
  if (const char* e = getenv("ULLAVA_GEMM_PAIR")) c->gemm_pair = atoi(e) != 0 ? 1 : 0;
  if (const char* e = getenv("ULLAVA_GEMM_HINTS")) c->gemm_hints = atoi(e) != 0 ? 1 : 0;
  if (const char* e = getenv("ULLAVA_GEMM_TMA_STORE")) c->gemm_tma_store = atoi(e) != 0 ? 1 : 0;
  if (const char* e = getenv("ULLAVA_GROUP_M")) {
    const int v = atoi(e);
    if (v > 0 && v <= 1024) c->group_m = v;
  }
  *out = c;
  return OK;
}

int ullava_destroy(ullava_ctx* ctx) {
  delete ctx;
  return OK;
}

int ullava_set_workspace(ullava_ctx* ctx, void* ptr, size_t bytes) {
  if (!ctx) {
    set_last_error("ullava_set_workspace: ctx is NULL");
    return ERR_BAD_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 255) != 0) {
    set_last_error("ullava_set_workspace: pointer must be 256-byte aligned");
    return ERR_BAD_ARG;
  }
  if (ptr != nullptr && bytes < kStreamCounterBytes) {
    set_last_error("ullava_set_workspace: at least %zu bytes are needed", kStreamCounterBytes);
    return ERR_BAD_ARG;
  }
  ctx->workspace = ptr;
  ctx->workspace_bytes = bytes;
  // tile-arrival counters of the weight-streaming GEMM live at the front and must start at zero
  if (ptr != nullptr) ULLAVA_CHECK_CUDA(cudaMemset(ptr, 0, kStreamCounterBytes));
  return OK;
}

int64_t ullava_launch_count(ullava_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ullava_profile_begin(ullava_ctx* ctx) {
  if (!ctx) { set_last_error("ullava_profile_begin: ctx is NULL"); return ERR_BAD_ARG; }
  for (auto& r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  ctx->prof.clear();
  ctx->prof_on = true;
  return OK;
}

int ullava_profile_end(ullava_ctx* ctx, double* out) {
  if (!ctx || !out) { set_last_error("ullava_profile_end: NULL argument"); return ERR_BAD_ARG; }
  ctx->prof_on = false;
  for (int i = 0; i < ULLAVA_PROF_CLASSES * 4; ++i) out[i] = 0.0;
  int st = OK;
  for (auto& r : ctx->prof) {
    if (st == OK) {
      cudaError_t e = cudaEventSynchronize(r.b);
      float ms = 0.f;
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.a, r.b);
      if (e != cudaSuccess) st = check_cuda(e, "ullava_profile_end");
      else if (r.cls >= 0 && r.cls < ULLAVA_PROF_CLASSES) {
        double* o = out + 4 * r.cls;
        o[0] += ms; o[1] += r.flops; o[2] += r.bytes; o[3] += r.launches;
      }
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  ctx->prof.clear();
  return st;
}

int ullava_gemm(ullava_ctx* ctx, const ullava_gemm_args* args, void* stream) {
  if (!ctx || !args) {
    set_last_error("ullava_gemm: NULL ctx/args");
    return ERR_BAD_ARG;
  }
  return gemm_run(ctx, *args, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
