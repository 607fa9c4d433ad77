// RoICrop: bilinear sampling of NCHW feature maps on a per-RoI (y, x) grid in [-1, 1]
// (BASELINE cfg3's "crop 14 -> maxpool 7" path).
//
// Reference semantics: lib/model/roi_crop/src/roi_crop_cuda_kernel.cu:12-23 (getTopLeft),
// :45-108 (forward), :111-190 (backward; the reference kernel accumulates the gradient of
// the images only and never writes the gradient of the grid), glue roi_crop_cuda.c.
//
// One CTA per (RoI, channel block): the GH*GW sampling sites (cell offset, the four corner
// weights with out-of-range corners zeroed) are computed once into shared memory and reused by
// every channel; outputs are written coalesced.  The reference recomputes the site geometry,
// including two divisions and four range tests, for every channel of every output element.
#include "common.cuh"

namespace tlod {

struct CropSite {
  int off;                  // yi * W + xi of the top-left corner (may be outside the map)
  float w00, w01, w10, w11; // corner weights, 0 where the corner is outside the map
  int m;                    // bit k set: corner k is inside
};

__device__ __forceinline__ void crop_top_left(float x, int size, int& point, float& weight) {
  const float coord = __fdiv_rn(__fmul_rn(__fadd_rn(x, 1.0f), (float)(size - 1)), 2.0f);
  const float fl = floorf(coord);
  point = (int)fl;
  weight = __fsub_rn(1.0f, __fsub_rn(coord, fl));
}

__device__ __forceinline__ CropSite crop_site(const float* __restrict__ g, int H, int W) {
  const float yf = __ldg(g), xf = __ldg(g + 1);
  int xi, yi;
  float xw, yw;
  crop_top_left(xf, W, xi, xw);
  crop_top_left(yf, H, yi, yw);
  const bool x0 = xi >= 0 && xi <= W - 1, x1 = xi + 1 >= 0 && xi + 1 <= W - 1;
  const bool y0 = yi >= 0 && yi <= H - 1, y1 = yi + 1 >= 0 && yi + 1 <= H - 1;
  CropSite s;
  s.off = yi * W + xi;
  s.m = (x0 && y0 ? 1 : 0) | (x1 && y0 ? 2 : 0) | (x0 && y1 ? 4 : 0) | (x1 && y1 ? 8 : 0);
  const float xw1 = __fsub_rn(1.0f, xw), yw1 = __fsub_rn(1.0f, yw);
  s.w00 = (s.m & 1) ? __fmul_rn(xw, yw) : 0.f;
  s.w01 = (s.m & 2) ? __fmul_rn(xw1, yw) : 0.f;
  s.w10 = (s.m & 4) ? __fmul_rn(xw, yw1) : 0.f;
  s.w11 = (s.m & 8) ? __fmul_rn(xw1, yw1) : 0.f;
  return s;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(256)
    roi_crop_kernel(const float* __restrict__ src, const float* __restrict__ grid, float* __restrict__ dst,
                    int C, int H, int W, int GH, int GW, int per_image, int chans_per_block) {
  extern __shared__ CropSite sites[];  // [GH * GW]
  const int b = blockIdx.x;
  const int c0 = blockIdx.y * chans_per_block;
  const int S = GH * GW;
  for (int i = threadIdx.x; i < S; i += blockDim.x) sites[i] = crop_site(grid + ((size_t)b * S + i) * 2, H, W);
  __syncthreads();
  const int bi = b / per_image;
  const int cb = min(chans_per_block, C - c0);
  const size_t tile = ((size_t)b * C + c0) * S;
  const size_t img = ((size_t)bi * C + c0) * H * W;
  for (int o = threadIdx.x; o < cb * S; o += blockDim.x) {
    const int c = o / S, i = o - c * S;
    const CropSite s = sites[i];
    const size_t cell = img + (size_t)c * H * W + s.off;  // only dereferenced where s.m allows
    if (!BACKWARD) {
      float v = 0.f;
      if (s.m & 1) v = __ldg(src + cell) * s.w00;
      if (s.m & 2) v = fmaf(__ldg(src + cell + 1), s.w01, v);
      if (s.m & 4) v = fmaf(__ldg(src + cell + W), s.w10, v);
      if (s.m & 8) v = fmaf(__ldg(src + cell + W + 1), s.w11, v);
      dst[tile + o] = v;
    } else {
      const float g = __ldg(src + tile + o);
      if (s.m & 1) atomicAdd(dst + cell, s.w00 * g);
      if (s.m & 2) atomicAdd(dst + cell + 1, s.w01 * g);
      if (s.m & 4) atomicAdd(dst + cell + W, s.w10 * g);
      if (s.m & 8) atomicAdd(dst + cell + W + 1, s.w11 * g);
    }
  }
}

static int launch_crop(bool backward, const float* src, const float* grid, float* dst, int in_batch,
                       int channels, int height, int width, int out_batch, int grid_h, int grid_w,
                       cudaStream_t st) {
  if (!src || !grid || !dst) return TLOD_ERR_NULL_POINTER;
  if (in_batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || out_batch < 0 || grid_h <= 0 || grid_w <= 0)
    return TLOD_ERR_BAD_SHAPE;
  if (out_batch % in_batch != 0) return TLOD_ERR_BAD_SHAPE;  // the reference divides silently
  if ((long long)in_batch * channels * height * width >= (1LL << 31)) return TLOD_ERR_INT32_OVERFLOW;
  const int S = grid_h * grid_w;
  const size_t smem = (size_t)S * sizeof(CropSite);
  if (smem > 48 * 1024) return TLOD_ERR_UNSUPPORTED;
  if (backward) {
    cudaError_t e = cudaMemsetAsync(dst, 0, (size_t)in_batch * channels * height * width * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
  }
  if (out_batch == 0) return TLOD_OK;
  int cpb = (4096 + S - 1) / S;
  if (cpb > channels) cpb = channels;
  while ((channels + cpb - 1) / cpb > 65535) ++cpb;
  dim3 g(out_batch, (channels + cpb - 1) / cpb);
  {
    LaunchScope scope(backward ? "roi_crop_bwd_kernel" : "roi_crop_fwd_kernel", st);
    if (backward)
      roi_crop_kernel<true><<<g, 256, smem, st>>>(src, grid, dst, channels, height, width, grid_h, grid_w,
                                                  out_batch / in_batch, cpb);
    else
      roi_crop_kernel<false><<<g, 256, smem, st>>>(src, grid, dst, channels, height, width, grid_h, grid_w,
                                                   out_batch / in_batch, cpb);
  }
  return last_launch_status();
}

}  // namespace tlod

using namespace tlod;

extern "C" int tlod_roi_crop_forward(const float* features, const float* grid_yx, float* output, int in_batch,
                                     int channels, int height, int width, int out_batch, int grid_h,
                                     int grid_w, void* stream) {
  return launch_crop(false, features, grid_yx, output, in_batch, channels, height, width, out_batch, grid_h,
                     grid_w, (cudaStream_t)stream);
}

extern "C" int tlod_roi_crop_backward(const float* grad_output, const float* grid_yx, float* grad_features,
                                      int in_batch, int channels, int height, int width, int out_batch,
                                      int grid_h, int grid_w, void* stream) {
  return launch_crop(true, grad_output, grid_yx, grad_features, in_batch, channels, height, width, out_batch,
                     grid_h, grid_w, (cudaStream_t)stream);
}
