// Library-level entry points of libtlod_b200.so.
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace tlod {

unsigned long long g_launches = 0ULL;
int g_profile_on = 0;

namespace {
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};
struct ProfTotal {
  double ms = 0.0;
  long long launches = 0;
};
std::mutex g_prof_mu;
std::vector<ProfRec*> g_prof_pending;
std::map<std::string, ProfTotal> g_prof_totals;
std::vector<std::string> g_prof_names;  // stable index order for tlod_profile_get
}  // namespace

void profile_begin(const char* name, cudaStream_t st, void** token) {
  ProfRec* r = new ProfRec;
  r->name = name;
  if (cudaEventCreate(&r->a) != cudaSuccess || cudaEventCreate(&r->b) != cudaSuccess) {
    delete r;
    return;
  }
  cudaEventRecord(r->a, st);
  *token = r;
}

void profile_end(void* token, cudaStream_t st) {
  ProfRec* r = (ProfRec*)token;
  cudaEventRecord(r->b, st);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof_pending.push_back(r);
}

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

// ---------------------------------------------------------------------------
// per-kernel timing for bench.py
// ---------------------------------------------------------------------------
extern "C" void tlod_profile_enable(int on) { tlod::g_profile_on = on ? 1 : 0; }

extern "C" void tlod_profile_reset(void) {
  std::lock_guard<std::mutex> lock(tlod::g_prof_mu);
  for (auto* r : tlod::g_prof_pending) {
    cudaEventDestroy(r->a);
    cudaEventDestroy(r->b);
    delete r;
  }
  tlod::g_prof_pending.clear();
  tlod::g_prof_totals.clear();
  tlod::g_prof_names.clear();
}

// Waits for the recorded events (this one call does synchronise the host) and folds them
// into the per-kernel totals.  Returns the number of distinct kernel names.
extern "C" int tlod_profile_collect(void) {
  std::lock_guard<std::mutex> lock(tlod::g_prof_mu);
  for (auto* r : tlod::g_prof_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(r->b) == cudaSuccess && cudaEventElapsedTime(&ms, r->a, r->b) == cudaSuccess) {
      auto it = tlod::g_prof_totals.find(r->name);
      if (it == tlod::g_prof_totals.end()) {
        tlod::g_prof_names.push_back(r->name);
        it = tlod::g_prof_totals.emplace(r->name, tlod::ProfTotal()).first;
      }
      it->second.ms += ms;
      it->second.launches += 1;
    }
    cudaEventDestroy(r->a);
    cudaEventDestroy(r->b);
    delete r;
  }
  tlod::g_prof_pending.clear();
  return (int)tlod::g_prof_names.size();
}

extern "C" int tlod_profile_get(int index, const char** name, double* total_ms, long long* launches) {
  std::lock_guard<std::mutex> lock(tlod::g_prof_mu);
  if (index < 0 || index >= (int)tlod::g_prof_names.size()) return TLOD_ERR_BAD_SHAPE;
  const std::string& n = tlod::g_prof_names[index];
  const tlod::ProfTotal& t = tlod::g_prof_totals[n];
  if (name) *name = n.c_str();
  if (total_ms) *total_ms = t.ms;
  if (launches) *launches = t.launches;
  return TLOD_OK;
}
