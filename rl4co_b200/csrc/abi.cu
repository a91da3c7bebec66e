// Library-wide state of libcorollout: error string, device attribute cache, version.
#include "co_common.cuh"

namespace co {
thread_local char g_err[512] = "";

const DeviceInfo& device_info() {
  static thread_local DeviceInfo info{0, 0};
  static thread_local int cached_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&info.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cached_dev = dev;
  }
  return info;
}
}  // namespace co

extern "C" int co_version(void) { return CO_VERSION; }
extern "C" const char* co_last_error_string(void) { return co::g_err; }
extern "C" int co_device_sm_count(void) { return co::device_info().sm_count; }
