// 2x2 / stride-1 average pooling over (tiles, AH, AW) -> (tiles, AH-1, AW-1) and its adjoint:
// the second half of RoIAlignAvg (lib/model/roi_align/modules/roi_align.py:26-29,
// `avg_pool2d(x, kernel_size=2, stride=1)`), which the reference leaves to a generic pooling
// kernel.  Both directions are pure streaming: a CTA stages a run of tiles in shared memory with
// coalesced 128-bit loads and writes the result with coalesced stores.
#include "common.cuh"

namespace tlod {

constexpr int AP_THREADS = 256;
constexpr int AP_SMEM_FLOATS = 8192;  // 32 KB of tiles per CTA pass

// grid-stride over groups of T tiles.  FWD: in tile (AH, AW) -> out tile (AH-1, AW-1).
// !FWD: in tile (AH-1, AW-1) (gradient of the pooled map) -> out tile (AH, AW).
// AHT / AWT > 0: compile-time tile size (8 x 8 for RoIAlignAvg(7, 7): every division below
// becomes a shift or a multiply); 0: run-time.
template <bool FWD, int AHT, int AWT>
__global__ void __launch_bounds__(AP_THREADS)
    avgpool2x2_kernel(const float* __restrict__ in, float* __restrict__ out, long long tiles, int AH_rt,
                      int AW_rt, int T) {
  extern __shared__ __align__(16) float sm[];
  const int AH = AHT > 0 ? AHT : AH_rt, AW = AWT > 0 ? AWT : AW_rt;
  const int PH = AH - 1, PW = AW - 1;
  const int s_in = FWD ? AH * AW : PH * PW;
  const int s_out = FWD ? PH * PW : AH * AW;
  const int wi = FWD ? AW : PW;  // row length of the staged tiles
  for (long long t0 = (long long)blockIdx.x * T; t0 < tiles; t0 += (long long)gridDim.x * T) {
    const int nt = (int)min((long long)T, tiles - t0);
    const float* src = in + t0 * s_in;
    const int n_in = nt * s_in;
    __syncthreads();  // previous pass' readers
    if ((((uintptr_t)src) & 15) == 0) {
      const float4* src4 = reinterpret_cast<const float4*>(src);
      for (int i = threadIdx.x; i < n_in / 4; i += AP_THREADS) reinterpret_cast<float4*>(sm)[i] = __ldg(src4 + i);
      for (int i = (n_in / 4) * 4 + threadIdx.x; i < n_in; i += AP_THREADS) sm[i] = __ldg(src + i);
    } else {
      for (int i = threadIdx.x; i < n_in; i += AP_THREADS) sm[i] = __ldg(src + i);
    }
    __syncthreads();
    float* dst = out + t0 * s_out;
    const int n_out = nt * s_out;
    for (int k = threadIdx.x; k < n_out; k += AP_THREADS) {
      const int tile = k / s_out, r = k - tile * s_out;
      const float* tp = sm + tile * s_in;
      float v;
      if (FWD) {
        const int i = r / PW, j = r - i * PW;
        const float* p = tp + i * wi + j;
        v = 0.25f * ((p[0] + p[1]) + (p[wi] + p[wi + 1]));
      } else {
        const int i = r / AW, j = r - i * AW;  // output cell (i, j) of the AH x AW tile
        float acc = 0.f;
        if (i > 0 && j > 0) acc += tp[(i - 1) * wi + (j - 1)];
        if (i > 0 && j < PW) acc += tp[(i - 1) * wi + j];
        if (i < PH && j > 0) acc += tp[i * wi + (j - 1)];
        if (i < PH && j < PW) acc += tp[i * wi + j];
        v = 0.25f * acc;
      }
      __stcs(dst + k, v);
    }
  }
}

static int launch_avgpool(bool fwd, const float* in, float* out, long long tiles, int ah, int aw,
                          cudaStream_t st) {
  if (!in || !out) return TLOD_ERR_NULL_POINTER;
  if (tiles < 0 || ah < 2 || aw < 2) return TLOD_ERR_BAD_SHAPE;
  if (tiles == 0) return TLOD_OK;
  const int s_big = ah * aw;
  if (s_big > AP_SMEM_FLOATS) return TLOD_ERR_UNSUPPORTED;
  int T = AP_SMEM_FLOATS / s_big;
  T = T / 4 * 4;  // keeps every pass' source 16-byte aligned when the tensor is
  if (T < 1) T = 1;
  long long groups = (tiles + T - 1) / T;
  long long grid = (long long)device_info().sm_count * 6;
  if (grid > groups) grid = groups;
  const size_t smem = (size_t)T * s_big * sizeof(float);
  {
    LaunchScope scope(fwd ? "avgpool2x2_fwd_kernel" : "avgpool2x2_bwd_kernel", st);
    if (ah == 8 && aw == 8) {
      if (fwd)
        avgpool2x2_kernel<true, 8, 8><<<(unsigned)grid, AP_THREADS, smem, st>>>(in, out, tiles, ah, aw, T);
      else
        avgpool2x2_kernel<false, 8, 8><<<(unsigned)grid, AP_THREADS, smem, st>>>(in, out, tiles, ah, aw, T);
    } else if (fwd) {
      avgpool2x2_kernel<true, 0, 0><<<(unsigned)grid, AP_THREADS, smem, st>>>(in, out, tiles, ah, aw, T);
    } else {
      avgpool2x2_kernel<false, 0, 0><<<(unsigned)grid, AP_THREADS, smem, st>>>(in, out, tiles, ah, aw, T);
    }
  }
  return last_launch_status();
}

}  // namespace tlod

using namespace tlod;

extern "C" int tlod_avgpool2x2_forward(const float* in, float* out, long long tiles, int height, int width,
                                       void* stream) {
  return launch_avgpool(true, in, out, tiles, height, width, (cudaStream_t)stream);
}

extern "C" int tlod_avgpool2x2_backward(const float* grad_out, float* grad_in, long long tiles, int height,
                                        int width, void* stream) {
  return launch_avgpool(false, grad_out, grad_in, tiles, height, width, (cudaStream_t)stream);
}
