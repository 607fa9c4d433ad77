// RPN head losses: classification cross-entropy over the non-ignored anchors
// and the smooth-L1 box loss, forward and backward, two launches each way.
//
// Reference semantics: lib/model/rpn/rpn.py:90-108 and _smooth_l1_loss,
// lib/model/utils/net_utils.py:72-86 (sigma = 3, dim = [1, 2, 3]):
//   score (B, 2A, H, W) viewed (B, 2, A*H, W): anchor cell i = (a*H + h)*W + w has the logits
//   (score[b, a, h, w], score[b, A + a, h, w]); label (B, 1, A*H, W) in {-1, 0, 1};
//   loss_cls = mean over the cells with label != -1 of (logsumexp - logit[label])
//   x = inside_w * (pred - target);  l = |x| < 1/9 ? 4.5 x^2 : |x| - 1/18
//   loss_box = sum(outside_w * l) / B
// The reference builds this from nonzero() + two index_select + cross_entropy + ten elementwise
// launches; here the forward is one launch (CTA partial sums in fp64, the last CTA to finish adds
// them in a fixed order, so the result is deterministic) and the backward one elementwise launch.
#include "common.cuh"

namespace tlod {

constexpr int RL_THREADS = 256;
constexpr int RL_MAXCTAS = 1024;

struct RpnLossWs {
  double part[RL_MAXCTAS][3];  // per CTA: cls sum, kept count, box sum
  unsigned int done;           // CTAs finished (zeroed by the entry point)
  int pad[3];
};

__device__ __forceinline__ double rl_block_sum(double v, double* scratch) {
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < RL_THREADS / 32; ++w) t += scratch[w];
  return t;
}

__device__ __forceinline__ float smooth_l1(float x, float sigma2) {
  const float ax = fabsf(x);
  return ax < 1.f / sigma2 ? x * x * (sigma2 * 0.5f) : ax - 0.5f / sigma2;
}

// out[0] = loss_cls, out[1] = loss_box, out[2] = kept anchors, out[3] = foreground anchors
__global__ void __launch_bounds__(RL_THREADS)
    rpn_loss_fwd_kernel(const float* __restrict__ score, const float* __restrict__ label,
                        const float* __restrict__ pred, const float* __restrict__ target,
                        const float* __restrict__ inside, const float* __restrict__ outside,
                        float* __restrict__ out, RpnLossWs* __restrict__ ws, int B, int A, int HW, float sigma2) {
  __shared__ double scratch[RL_THREADS / 32];
  __shared__ bool last;
  const long long cells = (long long)B * A * HW;  // anchors
  const long long stride = (long long)gridDim.x * RL_THREADS;
  double cls = 0.0, kept = 0.0, fg = 0.0;
  for (long long e = (long long)blockIdx.x * RL_THREADS + threadIdx.x; e < cells; e += stride) {
    const float lab = __ldg(label + e);
    if (lab != -1.f) {
      const long long b = e / ((long long)A * HW), i = e - b * (long long)A * HW;
      const float s0 = __ldg(score + b * 2 * A * HW + i), s1 = __ldg(score + b * 2 * A * HW + (long long)A * HW + i);
      const float m = fmaxf(s0, s1);
      const float lse = m + logf(expf(s0 - m) + expf(s1 - m));
      cls += (double)(lse - (lab != 0.f ? s1 : s0));
      kept += 1.0;
      if (lab != 0.f) fg += 1.0;
    }
  }
  double box = 0.0;
  const long long nbox = cells * 4;
  for (long long e = (long long)blockIdx.x * RL_THREADS + threadIdx.x; e < nbox; e += stride) {
    const float ow = __ldg(outside + e);
    if (ow != 0.f) box += (double)(ow * smooth_l1(__ldg(inside + e) * (__ldg(pred + e) - __ldg(target + e)), sigma2));
  }
  cls = rl_block_sum(cls, scratch);
  kept = rl_block_sum(kept, scratch);
  fg = rl_block_sum(fg, scratch);
  box = rl_block_sum(box, scratch);
  if (threadIdx.x == 0) {
    ws->part[blockIdx.x][0] = cls;
    ws->part[blockIdx.x][1] = kept + fg * 4294967296.0;  // two exact integer counts in one double
    ws->part[blockIdx.x][2] = box;
    __threadfence();
    last = atomicAdd(&ws->done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double c = 0.0, k = 0.0, f = 0.0, x = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += RL_THREADS) {  // fixed order per thread, then fixed tree
    c += ws->part[i][0];
    const double kf = ws->part[i][1];
    const double fi = floor(kf / 4294967296.0);
    f += fi;
    k += kf - fi * 4294967296.0;
    x += ws->part[i][2];
  }
  c = rl_block_sum(c, scratch);
  k = rl_block_sum(k, scratch);
  f = rl_block_sum(f, scratch);
  x = rl_block_sum(x, scratch);
  if (threadIdx.x == 0) {
    out[0] = k > 0.0 ? (float)(c / k) : 0.f;
    out[1] = (float)(x / (double)B);
    out[2] = (float)k;
    out[3] = (float)f;
  }
}

// g_score (B, 2A, H, W), g_pred (B, 4A, H, W); up (2 floats on the device, may be NULL):
// autograd's incoming gradients of (loss_cls, loss_box)
__global__ void __launch_bounds__(RL_THREADS)
    rpn_loss_bwd_kernel(const float* __restrict__ score, const float* __restrict__ label,
                        const float* __restrict__ pred, const float* __restrict__ target,
                        const float* __restrict__ inside, const float* __restrict__ outside,
                        const float* __restrict__ fwd_out, const float* __restrict__ up,
                        float* __restrict__ g_score, float* __restrict__ g_pred, int B, int A, int HW,
                        float sigma2) {
  const float u_cls = up ? __ldg(up) : 1.f, u_box = up ? __ldg(up + 1) : 1.f;
  const long long cells = (long long)B * A * HW;
  const long long e = (long long)blockIdx.x * RL_THREADS + threadIdx.x;
  if (e < cells) {
    const long long b = e / ((long long)A * HW), i = e - b * (long long)A * HW;
    const long long o0 = b * 2 * A * HW + i, o1 = o0 + (long long)A * HW;
    const float lab = __ldg(label + e);
    float g0 = 0.f, g1 = 0.f;
    if (lab != -1.f) {
      const float s0 = __ldg(score + o0), s1 = __ldg(score + o1);
      const float m = fmaxf(s0, s1);
      const float e0 = expf(s0 - m), e1 = expf(s1 - m);
      const float inv = 1.f / (e0 + e1);
      const float k = u_cls / __ldg(fwd_out + 2);
      g0 = k * (e0 * inv - (lab == 0.f ? 1.f : 0.f));
      g1 = k * (e1 * inv - (lab != 0.f ? 1.f : 0.f));
    }
    g_score[o0] = g0;
    g_score[o1] = g1;
  }
  if (e < cells * 4) {
    const float iw = __ldg(inside + e), ow = __ldg(outside + e);
    const float x = iw * (__ldg(pred + e) - __ldg(target + e));
    const float d = fabsf(x) < 1.f / sigma2 ? sigma2 * x : (x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f));
    g_pred[e] = u_box * ow * iw * d / (float)B;
  }
}

}  // namespace tlod

using namespace tlod;

extern "C" size_t tlod_rpn_loss_workspace_bytes(void) { return sizeof(RpnLossWs) + 256; }

extern "C" int tlod_rpn_loss_forward(const float* cls_score, const float* labels, const float* bbox_pred,
                                     const float* bbox_targets, const float* inside_w, const float* outside_w,
                                     float* losses_out, int batch, int num_anchors, int height, int width,
                                     float sigma, void* workspace, size_t workspace_bytes, void* stream) {
  if (!cls_score || !labels || !bbox_pred || !bbox_targets || !inside_w || !outside_w || !losses_out)
    return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || num_anchors <= 0 || height <= 0 || width <= 0 || !(sigma > 0.f)) return TLOD_ERR_BAD_SHAPE;
  if (!workspace || workspace_bytes < tlod_rpn_loss_workspace_bytes() || ((uintptr_t)workspace & 15))
    return TLOD_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  RpnLossWs* ws = (RpnLossWs*)workspace;
  cudaError_t e = cudaMemsetAsync(&ws->done, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return (int)e;
  const long long work = (long long)batch * num_anchors * height * width * 4;
  long long grid = (work + RL_THREADS * 8 - 1) / (RL_THREADS * 8);
  if (grid > RL_MAXCTAS) grid = RL_MAXCTAS;
  if (grid < 1) grid = 1;
  {
    LaunchScope scope("rpn_loss_fwd_kernel", st);
    rpn_loss_fwd_kernel<<<(unsigned)grid, RL_THREADS, 0, st>>>(cls_score, labels, bbox_pred, bbox_targets, inside_w,
                                                              outside_w, losses_out, ws, batch, num_anchors,
                                                              height * width, sigma * sigma);
  }
  return last_launch_status();
}

extern "C" int tlod_rpn_loss_backward(const float* cls_score, const float* labels, const float* bbox_pred,
                                      const float* bbox_targets, const float* inside_w, const float* outside_w,
                                      const float* losses_out, const float* upstream, float* grad_cls_score,
                                      float* grad_bbox_pred, int batch, int num_anchors, int height, int width,
                                      float sigma, void* stream) {
  if (!cls_score || !labels || !bbox_pred || !bbox_targets || !inside_w || !outside_w || !losses_out ||
      !grad_cls_score || !grad_bbox_pred)
    return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || num_anchors <= 0 || height <= 0 || width <= 0 || !(sigma > 0.f)) return TLOD_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)batch * num_anchors * height * width * 4;
  {
    LaunchScope scope("rpn_loss_bwd_kernel", st);
    rpn_loss_bwd_kernel<<<(unsigned)((n + RL_THREADS - 1) / RL_THREADS), RL_THREADS, 0, st>>>(
        cls_score, labels, bbox_pred, bbox_targets, inside_w, outside_w, losses_out, upstream, grad_cls_score,
        grad_bbox_pred, batch, num_anchors, height * width, sigma * sigma);
  }
  return last_launch_status();
}
