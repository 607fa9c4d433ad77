// Library-level entry points of libtlod_b200.so.
#include <mutex>

#include "common.cuh"

namespace tlod {

unsigned long long g_launches = 0ULL;

const DeviceInfo& device_info() {
  static DeviceInfo info[64];
  static bool ready[64] = {false};
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!ready[dev]) {
    std::lock_guard<std::mutex> lock(mu);
    if (!ready[dev]) {
      DeviceInfo d;
      d.sm_count = 148;
      d.max_smem_optin = 48 * 1024;
      cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
      cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      info[dev] = d;
      __atomic_store_n(&ready[dev], true, __ATOMIC_RELEASE);
    }
  }
  return info[dev];
}

}  // namespace tlod

extern "C" int tlod_version(void) { return TLOD_B200_VERSION; }

extern "C" unsigned long long tlod_launch_count(void) {
  return __atomic_load_n(&tlod::g_launches, __ATOMIC_RELAXED);
}

extern "C" const char* tlod_error_string(int code) {
  switch (code) {
    case TLOD_OK: return "ok";
    case TLOD_ERR_NULL_POINTER: return "tlod: null pointer argument";
    case TLOD_ERR_BAD_SHAPE: return "tlod: non-positive or inconsistent shape argument";
    case TLOD_ERR_UNSUPPORTED: return "tlod: request outside the implemented range";
    case TLOD_ERR_WORKSPACE: return "tlod: workspace missing, misaligned or too small";
    case TLOD_ERR_INT32_OVERFLOW: return "tlod: tensor has 2^31 or more elements";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "tlod: unknown error code";
}
