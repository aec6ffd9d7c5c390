// SM partitions: two streams whose kernels run on DISJOINT sets of SMs of one B200 (CUDA green contexts).
//
// Why: evaluate() (models/ullava.py:335-434) runs two stages with complementary bottlenecks back to back -- the 64
// decode steps stream 13.5 GB of weights + the KV cache per step (HBM-bound, tensor pipe ~6 % busy) and the SAM
// ViT-H image encoder is dense tcgen05 GEMM / attention work (tensor-bound, ~1 TB/s of HBM) that depends only on
// images_sam (the reference merely happens to call get_visual_embs after generate, :399).  Run concurrently on a
// spatial split of the 148 SMs they overlap instead of queueing; plain stream concurrency does not do it, because
// every persistent kernel here sizes its grid to the whole machine and a 1 ms GEMM would hold all SMs against the
// 30 us decode kernels.  A green context confines every kernel launched into its streams -- persistent or not -- to
// its SM set; the per-lane ullava_ctx (ullava_set_sm_limit) sizes persistent grids / stream-K splits to that set.
//
// The driver API is reached through cudaGetDriverEntryPoint (the library does not link libcuda).
#include "common.cuh"
#include "ullava_internal.h"

#include <mutex>

struct ullava_partition {
  int device = 0;
  CUgreenCtx green[2] = {nullptr, nullptr};
  CUstream stream[2] = {nullptr, nullptr};
  int sms[2] = {0, 0};
};

namespace ullava {

namespace {
struct DriverFns {
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*,
                                        unsigned int, unsigned int) = nullptr;
  CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
  CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  CUresult (*StreamDestroy)(CUstream) = nullptr;
  bool ok = false;
};
DriverFns g_drv;
std::once_flag g_drv_once;

template <typename F>
bool load(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
    return false;
  fn = reinterpret_cast<F>(p);
  return true;
}

void load_driver() {
  DriverFns& d = g_drv;
  d.ok = load("cuDeviceGet", d.DeviceGet) && load("cuDeviceGetDevResource", d.DeviceGetDevResource) &&
         load("cuDevSmResourceSplitByCount", d.DevSmResourceSplitByCount) &&
         load("cuDevResourceGenerateDesc", d.DevResourceGenerateDesc) && load("cuGreenCtxCreate", d.GreenCtxCreate) &&
         load("cuGreenCtxDestroy", d.GreenCtxDestroy) && load("cuGreenCtxStreamCreate", d.GreenCtxStreamCreate) &&
         load("cuStreamDestroy", d.StreamDestroy);
}

int drv_fail(CUresult r, const char* what) {
  set_last_error("SM partition: %s failed (CUresult %d)", what, static_cast<int>(r));
  return ERR_CUDA;
}
}  // namespace

}  // namespace ullava

using namespace ullava;

extern "C" {

int ullava_partition_create(int device, int sms_a, int priority_a, int priority_b, ullava_partition** out) {
  if (!out) { set_last_error("ullava_partition_create: out is NULL"); return ERR_BAD_ARG; }
  *out = nullptr;
  std::call_once(g_drv_once, load_driver);
  if (!g_drv.ok) {
    set_last_error("SM partition: the driver does not export the green-context API (needs CUDA >= 12.4)");
    return ERR_UNSUPPORTED;
  }
  int prev_device = -1;
  ULLAVA_CHECK_CUDA(cudaGetDevice(&prev_device));
  struct Restore {   // the caller's current device is not ours to change
    int dev;
    ~Restore() { if (dev >= 0) cudaSetDevice(dev); }
  } restore{prev_device != device ? prev_device : -1};
  ULLAVA_CHECK_CUDA(cudaSetDevice(device));
  ULLAVA_CHECK_CUDA(cudaFree(nullptr));  // primary context up before the green contexts retain it
  CUdevice dev;
  CUresult r = g_drv.DeviceGet(&dev, device);
  if (r != CUDA_SUCCESS) return drv_fail(r, "cuDeviceGet");
  CUdevResource all{};
  r = g_drv.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM);
  if (r != CUDA_SUCCESS) return drv_fail(r, "cuDeviceGetDevResource");
  const int total = static_cast<int>(all.sm.smCount);
  if (sms_a < 8 || sms_a > total - 8) {
    set_last_error("SM partition: %d SMs asked for lane A, the device has %d (both lanes need >= 8)", sms_a, total);
    return ERR_BAD_ARG;
  }
  // group A = sms_a SMs (rounded up to the architecture's granularity, 8 on sm_90+), lane B = everything else
  CUdevResource part[2]{};
  unsigned int n_groups = 1;
  r = g_drv.DevSmResourceSplitByCount(&part[0], &n_groups, &all, &part[1], 0, static_cast<unsigned int>(sms_a));
  if (r != CUDA_SUCCESS || n_groups != 1) return drv_fail(r, "cuDevSmResourceSplitByCount");
  if (part[1].sm.smCount < 8) {
    set_last_error("SM partition: only %u SMs left for lane B", part[1].sm.smCount);
    return ERR_BAD_ARG;
  }
  ullava_partition* p = new ullava_partition();
  p->device = device;
  const int prio[2] = {priority_a, priority_b};
  for (int i = 0; i < 2; ++i) {
    CUdevResourceDesc desc = nullptr;
    r = g_drv.DevResourceGenerateDesc(&desc, &part[i], 1);
    if (r == CUDA_SUCCESS) r = g_drv.GreenCtxCreate(&p->green[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM);
    if (r == CUDA_SUCCESS) r = g_drv.GreenCtxStreamCreate(&p->stream[i], p->green[i], CU_STREAM_NON_BLOCKING, prio[i]);
    if (r != CUDA_SUCCESS) {
      const int st = drv_fail(r, "cuGreenCtxCreate / cuGreenCtxStreamCreate");
      ullava_partition_destroy(p);
      return st;
    }
    p->sms[i] = static_cast<int>(part[i].sm.smCount);
  }
  *out = p;
  return OK;
}

int ullava_partition_info(const ullava_partition* p, int32_t* sms_a, int32_t* sms_b, void** stream_a, void** stream_b) {
  if (!p) { set_last_error("ullava_partition_info: NULL partition"); return ERR_BAD_ARG; }
  if (sms_a) *sms_a = p->sms[0];
  if (sms_b) *sms_b = p->sms[1];
  if (stream_a) *stream_a = p->stream[0];
  if (stream_b) *stream_b = p->stream[1];
  return OK;
}

int ullava_partition_destroy(ullava_partition* p) {
  if (!p) return OK;
  for (int i = 0; i < 2; ++i) {
    if (p->stream[i]) g_drv.StreamDestroy(p->stream[i]);
    if (p->green[i]) g_drv.GreenCtxDestroy(p->green[i]);
  }
  delete p;
  return OK;
}

int ullava_set_sm_limit(ullava_ctx* ctx, int sms) {
  if (!ctx) { set_last_error("ullava_set_sm_limit: ctx is NULL"); return ERR_BAD_ARG; }
  if (sms < 0 || sms > ctx->device_sm_count) {
    set_last_error("ullava_set_sm_limit: %d not in [0, %d]", sms, ctx->device_sm_count);
    return ERR_BAD_ARG;
  }
  ctx->sm_count = sms == 0 ? ctx->device_sm_count : sms;
  return OK;
}

int ullava_sm_count(ullava_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

}  // extern "C"
