// MAF / PT-MAF scale-reduce rearrangement: the chunk / reshape / cat loops of
// DRM.forward (lib/MAF/drm.py:21-42) are a space-to-depth (pixel-unshuffle):
//   out[b, c*s*s + dy*s + dx, i, j] = in[b, c, i*s + dy, j*s + dx]   for i < H/s, j < W/s
// (rows / columns beyond s*floor(H/s), s*floor(W/s) are dropped).  The reference builds it from
// (H/s)*(W/s) Python-level chunk + reshape + cat calls -- ~2800 tiny launches for the 150x300
// conv3 map at s = 4; here it is one streaming launch each way.  The label-resize layers
// (lib/DAF/LabelResizeLayer.py:25-58: two .cpu() copies + cv2.resize per call) become device
// fills: the label map of a domain is a constant.
#include "common.cuh"

namespace tlod {

// one thread per OUTPUT element of the unshuffled tensor (coalesced writes); BWD: the same index
// map, reading the gradient of the output and writing (with zeros in the dropped border) the input.
template <bool BWD>
__global__ void __launch_bounds__(256)
    space_to_depth_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W, int s,
                          long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int h2 = H / s, w2 = W / s;
  if (!BWD) {
    const int j = (int)(e % w2);
    long long r = e / w2;
    const int i = (int)(r % h2); r /= h2;
    const int k = (int)(r % (s * s)); r /= (s * s);
    const int c = (int)(r % C);
    const long long b = r / C;
    const int dy = k / s, dx = k - dy * s;
    dst[e] = __ldg(src + ((b * C + c) * H + (i * s + dy)) * (long long)W + (j * s + dx));
  } else {
    // e indexes the INPUT gradient (B, C, H, W): coalesced writes, zeros in the dropped border
    const int x = (int)(e % W);
    long long r = e / W;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r % C);
    const long long b = r / C;
    float v = 0.f;
    if (y < h2 * s && x < w2 * s) {
      const int i = y / s, dy = y - i * s, j = x / s, dx = x - j * s;
      v = __ldg(src + (((b * C + c) * (s * s) + dy * s + dx) * h2 + i) * (long long)w2 + j);
    }
    dst[e] = v;
  }
}

__global__ void __launch_bounds__(256)
    instance_labels_kernel(const float* __restrict__ domain, float* __restrict__ out, int rows, int images,
                           int minibatch, float fill) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int i = r / minibatch;
  out[r] = i < images ? __ldg(domain + i) : fill;
}

}  // namespace tlod

using namespace tlod;

extern "C" int tlod_space_to_depth_forward(const float* in, float* out, int batch, int channels, int height,
                                           int width, int scale, void* stream) {
  if (!in || !out) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || scale <= 0 || height < scale || width < scale)
    return TLOD_ERR_BAD_SHAPE;
  const long long total = (long long)batch * channels * scale * scale * (height / scale) * (width / scale);
  cudaStream_t st = (cudaStream_t)stream;
  {
    LaunchScope scope("space_to_depth_fwd_kernel", st);
    space_to_depth_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, channels, height, width,
                                                                              scale, total);
  }
  return last_launch_status();
}

extern "C" int tlod_space_to_depth_backward(const float* grad_out, float* grad_in, int batch, int channels,
                                            int height, int width, int scale, void* stream) {
  if (!grad_out || !grad_in) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || scale <= 0 || height < scale || width < scale)
    return TLOD_ERR_BAD_SHAPE;
  const long long total = (long long)batch * channels * height * width;
  cudaStream_t st = (cudaStream_t)stream;
  {
    LaunchScope scope("space_to_depth_bwd_kernel", st);
    space_to_depth_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(grad_out, grad_in, channels, height,
                                                                             width, scale, total);
  }
  return last_launch_status();
}

extern "C" int tlod_instance_labels(const float* domain_labels, float* out, int rows, int images, int minibatch,
                                    float fill, void* stream) {
  if (!domain_labels || !out) return TLOD_ERR_NULL_POINTER;
  if (rows < 0 || images <= 0 || minibatch <= 0) return TLOD_ERR_BAD_SHAPE;
  if (rows == 0) return TLOD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  {
    LaunchScope scope("instance_labels_kernel", st);
    instance_labels_kernel<<<(rows + 255) / 256, 256, 0, st>>>(domain_labels, out, rows, images, minibatch, fill);
  }
  return last_launch_status();
}
