// Host-side plumbing of libcf_b200: error string, version, device properties.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void cf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cf_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;  // B200
  }
  return sms;
}

extern "C" const char* cf_last_error(void) { return g_err; }
extern "C" int cf_abi_version(void) { return CF_ABI_VERSION; }
extern "C" const char* cf_build_arch(void) { return "sm_100a"; }

// ---- CUDA IPC plumbing of the peer-pull exchange (include/cf_b200.h); the reference is single-device
#include <cuda.h>

extern "C" int cf_ipc_export(const void* devptr, void* handle64, int64_t* offset_bytes) {
  CF_CHECK_ARG(devptr && handle64 && offset_bytes, "cf_ipc_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  typedef CUresult (*range_fn_t)(CUdeviceptr*, size_t*, CUdeviceptr);
  static range_fn_t range_fn = nullptr;
  if (!range_fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      range_fn = reinterpret_cast<range_fn_t>(p);
  }
  CF_CHECK_ARG(range_fn != nullptr, "cf_ipc_export: cuMemGetAddressRange is not available from the driver");
  CUdeviceptr base = 0;
  size_t size = 0;
  CF_CHECK_ARG(range_fn(&base, &size, (CUdeviceptr)(uintptr_t)devptr) == CUDA_SUCCESS, "cf_ipc_export: not a device allocation");
  cudaIpcMemHandle_t h;
  CF_CUDA_OK(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
  memcpy(handle64, &h, 64);
  *offset_bytes = (int64_t)((uintptr_t)devptr - (uintptr_t)base);
  return 0;
}

extern "C" int cf_ipc_open(const void* handle64, void** base) {
  CF_CHECK_ARG(handle64 && base, "cf_ipc_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CF_CUDA_OK(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int cf_ipc_close(void* base) {
  CF_CHECK_ARG(base != nullptr, "cf_ipc_close: NULL argument");
  CF_CUDA_OK(cudaIpcCloseMemHandle(base));
  return 0;
}
