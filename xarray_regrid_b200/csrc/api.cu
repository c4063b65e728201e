// Library-level entry points of the regrid_b200 C ABI: version, error text, device info.
#include "common.cuh"

namespace rg {

int device_info(DeviceInfo* out) {
  // Per-device cache of two immutable attributes (benign race: every writer stores the same values).
  static DeviceInfo cache[64];
  static bool have[64];
  int dev = 0;
  RG_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && have[dev]) {
    *out = cache[dev];
    return RG_OK;
  }
  DeviceInfo info;
  RG_CUDA(cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev));
  RG_CUDA(cudaDeviceGetAttribute(&info.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  RG_CUDA(cudaDeviceGetAttribute(&info.max_smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
  if (dev >= 0 && dev < 64) {
    cache[dev] = info;
    have[dev] = true;
  }
  *out = info;
  return RG_OK;
}

}  // namespace rg

extern "C" {

int rg_abi_version(void) { return RG_ABI_VERSION; }

const char* rg_error_string(int code) {
  switch (code) {
    case RG_OK: return "ok";
    case RG_ERR_NULL: return "regrid_b200: a required pointer is NULL";
    case RG_ERR_SHAPE: return "regrid_b200: invalid extent, stride or workspace size";
    case RG_ERR_SMEM: return "regrid_b200: problem does not fit the kernel's shared memory";
    case RG_ERR_UNSUPPORTED: return "regrid_b200: unsupported configuration";
    case RG_ERR_ALIGN: return "regrid_b200: pointer is not sufficiently aligned";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "regrid_b200: unknown error code";
}

int rg_device_sm_count(void) {
  rg::DeviceInfo info;
  if (rg::device_info(&info) != RG_OK) return -1;
  return info.sm_count;
}

}  // extern "C"
