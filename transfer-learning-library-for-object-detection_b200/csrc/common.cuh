// Shared helpers for libtlod_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tlod_b200.h"

namespace tlod {

// Count of kernel launches enqueued by the library (tlod_launch_count()).
extern unsigned long long g_launches;
inline void count_launch(int n = 1) {
  __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED);
}

struct DeviceInfo {
  int sm_count;
  int max_smem_optin;  // bytes of dynamic shared memory one CTA may opt in to
};
// Cached per device; cheap after the first call on a device.
const DeviceInfo& device_info();

inline int last_launch_status() { return (int)cudaGetLastError(); }

// Optional per-kernel timing (tlod_profile_*): when enabled, every launch is bracketed by
// two CUDA events on its own stream; bench.py reads the accumulated durations.
extern int g_profile_on;
void profile_begin(const char* name, cudaStream_t st, void** token);
void profile_end(void* token, cudaStream_t st);

// RAII around one kernel launch: counts it and, if profiling is on, times it.
struct LaunchScope {
  cudaStream_t st;
  void* token;
  LaunchScope(const char* name, cudaStream_t s) : st(s), token(nullptr) {
    count_launch();
    if (g_profile_on) profile_begin(name, st, &token);
  }
  ~LaunchScope() {
    if (token) profile_end(token, st);
  }
};

// tlod_roi_pool_forward; batch_known = false (compat_launchers.cu): `batch` is an upper bound only
int roi_pool_forward_launch(const float* features, const float* rois, float* output, int* argmax, int batch,
                            int channels, int height, int width, int num_rois, int pooled_h, int pooled_w,
                            float spatial_scale, bool batch_known, cudaStream_t stream);

constexpr int kWarp = 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// ---------------------------------------------------------------------------
// RoIAlign sampling geometry along one axis, op-for-op the reference's
// arithmetic (lib/model/roi_align/src/roi_align_kernel.cu:31-49,57-58):
//   lo = px * scale; extent = max(hi - lo + 1., 0); bin = extent / (aligned - 1.)
//   pos = p * bin + lo; start = min(floor(pos), size - 2); ratio = pos - start
// The `1.` literals are doubles in the reference, so those two steps are done in
// fp64 here as well; the fp32 steps use the _rn intrinsics so that nvcc cannot
// contract them into FMAs (the contract is the source-level formula).
// ---------------------------------------------------------------------------
struct AlignAxis {
  int start;    // first of the two cells, already clamped to size-2
  float ratio;  // pos - start, may exceed 1 on the last cell (extrapolation)
  bool valid;   // !(pos < 0 || pos >= size)
};

__device__ __forceinline__ AlignAxis align_axis(float lo_px, float hi_px, float scale, int aligned,
                                                int size, int p) {
  float lo = __fmul_rn(lo_px, scale);
  float hi = __fmul_rn(hi_px, scale);
  float extent = fmaxf((float)((double)__fsub_rn(hi, lo) + 1.0), 0.f);
  float bin = (float)((double)extent / ((double)aligned - 1.0));
  float pos = __fadd_rn(__fmul_rn((float)p, bin), lo);
  AlignAxis a;
  a.start = (int)fminf(floorf(pos), (float)(size - 2));
  a.valid = !(pos < 0.f || pos >= (float)size);
  a.ratio = __fsub_rn(pos, (float)a.start);
  return a;
}

// Entry of the per-RoI sampling tables kept in shared memory.
//   off : start (columns) or start * row_stride (rows); -1 marks an out-of-range sample
//   w0  : weight of cell `start`     = (float)(1. - ratio)
//   w1  : weight of cell `start + 1` = ratio
struct AxisTab {
  int off;
  float w0;
  float w1;
};

__device__ __forceinline__ AxisTab make_tab(const AlignAxis& a, int stride) {
  AxisTab t;
  t.off = a.valid ? a.start * stride : -1;
  t.w0 = (float)(1.0 - (double)a.ratio);
  t.w1 = a.ratio;
  return t;
}

}  // namespace tlod
