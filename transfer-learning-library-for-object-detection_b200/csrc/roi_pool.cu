// RoIPool (Caffe max pooling over RoI bins) forward / backward for sm_100a.
//
// Reference semantics: lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93
// (forward, argmax saved) and :128-203 (backward, a gather over all RoIs).
// The backward here is a scatter by argmax that applies exactly the
// reference's admission tests (same image, same channel, cell inside the
// rounded RoI, bin inside the cell's feasible set), so it differs from the
// reference only in fp32 summation order.
#include <float.h>

#include "common.cuh"

namespace tlod {

struct PoolRoi {
  int batch, sw, sh, ew, eh;
  float bin_h, bin_w;
};

// roi_pooling_kernel.cu:44-55
__device__ __forceinline__ PoolRoi pool_roi(const float* __restrict__ r, float scale, int PH,
                                            int PW) {
  PoolRoi g;
  g.batch = (int)__ldg(r);
  g.sw = (int)roundf(__fmul_rn(__ldg(r + 1), scale));
  g.sh = (int)roundf(__fmul_rn(__ldg(r + 2), scale));
  g.ew = (int)roundf(__fmul_rn(__ldg(r + 3), scale));
  g.eh = (int)roundf(__fmul_rn(__ldg(r + 4), scale));
  const int rw = max(g.ew - g.sw + 1, 1);
  const int rh = max(g.eh - g.sh + 1, 1);
  g.bin_h = __fdiv_rn((float)rh, (float)PH);
  g.bin_w = __fdiv_rn((float)rw, (float)PW);
  return g;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// PHT, PWT > 0: compile-time pooled size (7 x 7: the index divisions become multiplications).
template <int PHT, int PWT>
__global__ void __launch_bounds__(256)
    roi_pool_fwd_kernel(const float* __restrict__ features, const float* __restrict__ rois,
                        float* __restrict__ output, int* __restrict__ argmax, int B, int C, int H,
                        int W, int PH_rt, int PW_rt, float scale, int chans_per_block) {
  const int PH = PHT > 0 ? PHT : PH_rt, PW = PWT > 0 ? PWT : PW_rt;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * chans_per_block;
  const PoolRoi g = pool_roi(rois + (size_t)n * 5, scale, PH, PW);
  const int S = PH * PW;
  const int cb = min(chans_per_block, C - c0);
  const bool image_ok = g.batch >= 0 && g.batch < B;
  const size_t roi_base = ((size_t)n * C + c0) * S;
  for (int o = threadIdx.x; o < cb * S; o += blockDim.x) {
    const int c = o / S;
    const int i = o - c * S;
    const int ph = i / PW;
    const int pw = i - ph * PW;
    int hs = (int)floorf(__fmul_rn((float)ph, g.bin_h));
    int he = (int)ceilf(__fmul_rn((float)(ph + 1), g.bin_h));
    int ws = (int)floorf(__fmul_rn((float)pw, g.bin_w));
    int we = (int)ceilf(__fmul_rn((float)(pw + 1), g.bin_w));
    hs = clampi(hs + g.sh, 0, H);
    he = clampi(he + g.sh, 0, H);
    ws = clampi(ws + g.sw, 0, W);
    we = clampi(we + g.sw, 0, W);
    const bool empty = (he <= hs) || (we <= ws) || !image_ok;
    float best = empty ? 0.f : -FLT_MAX;
    int besti = -1;
    if (!empty) {
      const int plane = (g.batch * C + c0 + c) * H * W;
      const float* p = features + plane;
      for (int h = hs; h < he; ++h)
        for (int w = ws; w < we; ++w) {
          const float v = __ldg(p + h * W + w);
          if (v > best) {
            best = v;
            besti = plane + h * W + w;
          }
        }
    }
    output[roi_base + o] = best;
    if (argmax) argmax[roi_base + o] = besti;
  }
}

template <int PHT, int PWT>
__global__ void __launch_bounds__(256)
    roi_pool_bwd_kernel(const float* __restrict__ top_grad, const int* __restrict__ argmax,
                        const float* __restrict__ rois, float* __restrict__ bottom_grad, int B, int C,
                        int H, int W, int PH_rt, int PW_rt, float scale, int chans_per_block) {
  const int PH = PHT > 0 ? PHT : PH_rt, PW = PWT > 0 ? PWT : PW_rt;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * chans_per_block;
  const PoolRoi g = pool_roi(rois + (size_t)n * 5, scale, PH, PW);
  if (g.batch < 0 || g.batch >= B) return;
  const int S = PH * PW;
  const int cb = min(chans_per_block, C - c0);
  const size_t roi_base = ((size_t)n * C + c0) * S;
  const int HW = H * W;
  const int total = B * C * HW;
  for (int o = threadIdx.x; o < cb * S; o += blockDim.x) {
    const int idx = __ldg(argmax + roi_base + o);
    if (idx < 0 || idx >= total) continue;
    const int c = c0 + o / S;
    const int i = o % S;
    const int ph = i / PW, pw = i - ph * PW;
    // decode the cell: the reference's thread for cell `idx` only looks at RoIs of its own image
    // (:150) and at argmax entries of its own channel (:189) -- i.e. idx must lie in the plane of
    // (this RoI's image, this channel); then one division gives the cell
    const int rem = idx - (g.batch * C + c) * HW;
    if (rem < 0 || rem >= HW) continue;
    const int h = rem / W;
    const int w = rem - h * W;
    if (!(w >= g.sw && w <= g.ew && h >= g.sh && h <= g.eh)) continue;  // :157-159
    int phs = (int)floorf(__fdiv_rn((float)(h - g.sh), g.bin_h));       // :178-186
    int phe = (int)ceilf(__fdiv_rn((float)(h - g.sh + 1), g.bin_h));
    int pws = (int)floorf(__fdiv_rn((float)(w - g.sw), g.bin_w));
    int pwe = (int)ceilf(__fdiv_rn((float)(w - g.sw + 1), g.bin_w));
    phs = clampi(phs, 0, PH);
    phe = clampi(phe, 0, PH);
    pws = clampi(pws, 0, PW);
    pwe = clampi(pwe, 0, PW);
    if (ph < phs || ph >= phe || pw < pws || pw >= pwe) continue;
    atomicAdd(bottom_grad + idx, __ldg(top_grad + roi_base + o));
  }
}

static int pool_check(const void* a, const void* b, const void* c, int batch, int channels,
                      int height, int width, int num_rois, int ph, int pw) {
  if (!a || !b || !c) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || num_rois < 0 || ph <= 0 || pw <= 0)
    return TLOD_ERR_BAD_SHAPE;
  if ((long long)batch * channels * height * width >= (1LL << 31)) return TLOD_ERR_INT32_OVERFLOW;
  return TLOD_OK;
}

static dim3 pool_grid(int num_rois, int channels, int S, int* cpb_out) {
  int cpb = (2048 + S - 1) / S;
  if (cpb > channels) cpb = channels;
  while ((channels + cpb - 1) / cpb > 65535) ++cpb;
  *cpb_out = cpb;
  return dim3(num_rois, (channels + cpb - 1) / cpb);
}

}  // namespace tlod

using namespace tlod;

extern "C" int tlod_roi_pool_forward(const float* features, const float* rois, float* output,
                                     int* argmax, int batch, int channels, int height, int width,
                                     int num_rois, int pooled_h, int pooled_w, float spatial_scale,
                                     void* stream) {
  int rc = pool_check(features, rois, output, batch, channels, height, width, num_rois, pooled_h,
                      pooled_w);
  if (rc != TLOD_OK) return rc;
  if (num_rois == 0) return TLOD_OK;
  int cpb;
  dim3 grid = pool_grid(num_rois, channels, pooled_h * pooled_w, &cpb);
  {
    LaunchScope scope("roi_pool_fwd_kernel", (cudaStream_t)stream);
    auto kern = (pooled_h == 7 && pooled_w == 7) ? roi_pool_fwd_kernel<7, 7> : roi_pool_fwd_kernel<0, 0>;
    kern<<<grid, 256, 0, (cudaStream_t)stream>>>(features, rois, output, argmax, batch, channels, height, width,
                                                 pooled_h, pooled_w, spatial_scale, cpb);
  }
  return last_launch_status();
}

extern "C" int tlod_roi_pool_backward(const float* top_grad, const int* argmax, const float* rois,
                                      float* bottom_grad, int batch, int channels, int height,
                                      int width, int num_rois, int pooled_h, int pooled_w,
                                      float spatial_scale, void* stream) {
  int rc = pool_check(top_grad, rois, bottom_grad, batch, channels, height, width, num_rois,
                      pooled_h, pooled_w);
  if (rc != TLOD_OK) return rc;
  if (!argmax) return TLOD_ERR_NULL_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(bottom_grad, 0,
                                  (size_t)batch * channels * height * width * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  if (num_rois == 0) return TLOD_OK;
  int cpb;
  dim3 grid = pool_grid(num_rois, channels, pooled_h * pooled_w, &cpb);
  {
    LaunchScope scope("roi_pool_bwd_kernel", st);
    auto kern = (pooled_h == 7 && pooled_w == 7) ? roi_pool_bwd_kernel<7, 7> : roi_pool_bwd_kernel<0, 0>;
    kern<<<grid, 256, 0, st>>>(top_grad, argmax, rois, bottom_grad, batch, channels, height, width, pooled_h,
                               pooled_w, spatial_scale, cpb);
  }
  return last_launch_status();
}
